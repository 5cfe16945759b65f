"""ncu target: one ICP3D::Run on the bunny Go-ICP clouds (NUMERICS env: 0 strict, 2 fast)."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("cuda-go-icp_b200")
G = os.path.join(ROOT, "tests", "golden")
ld = lambda n: np.fromfile(os.path.join(G, n), np.float32).reshape(-1, 3)
g = pkg.GoICP(1e-3); g.pModel, g.pData = ld("bunny_model_s0.1_seed1234.f32"), ld("bunny_data_s0.1_seed1235.f32"); g.numerics = int(os.environ.get("NUMERICS", "0"))
print(g.ICP(np.eye(3), np.zeros(3), 10000, 1e-7)[3])
g.close()
