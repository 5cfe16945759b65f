import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("cuda-go-icp_b200")
G = os.path.join(ROOT, "tests", "golden")
model = np.fromfile(os.path.join(G, "bunny_model_s0.1_seed1234.f32"), np.float32).reshape(-1, 3)
g = pkg.GoICP(1e-3); g.pModel, g.pData = model, model[:100]; g.dt.SIZE = int(sys.argv[1]) if len(sys.argv) > 1 else 128
g.BuildDT(); g.close()
