import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("cuda-go-icp_b200")
G = os.path.join(ROOT, "tests", "golden")
model = np.fromfile(os.path.join(G, "bunny_model_s0.1_seed1234.f32"), np.float32).reshape(-1, 3)
data = np.fromfile(os.path.join(G, "bunny_data_s0.1_seed1235.f32"), np.float32).reshape(-1, 3)
mse = float(sys.argv[1]) if len(sys.argv) > 1 else 1e-3
g = pkg.GoICP(mse); g.pModel, g.pData = model, data; g.dt_mode = int(os.environ.get("DT_MODE", "0")); g.spec_cubes = int(os.environ.get("SPEC", "0")); g.cluster_size = int(os.environ.get("CLUSTER", "0"))
g.BuildDT()
for rep in range(int(os.environ.get("REPS", "2"))):
    g.Register()
    print({k: (round(v, 5) if isinstance(v, float) else v) for k, v in g.result.items() if k not in ("R", "t")}, flush=True)
g.close()
