import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("cuda-go-icp_b200")
G = os.path.join(ROOT, "tests", "golden")
model = np.fromfile(os.path.join(G, "bunny_model_s0.1_seed1234.f32"), np.float32).reshape(-1, 3)
data = np.fromfile(os.path.join(G, "bunny_data_s0.1_seed1235.f32"), np.float32).reshape(-1, 3)
for S in [int(a) for a in sys.argv[1:]] or [100, 300]:
    g = pkg.GoICP(1e-3); g.pModel, g.pData = model, data; g.dt.SIZE = S
    ts = []
    for _ in range(5):
        t = time.time(); g.BuildDT(); ts.append(round(time.time() - t, 4))
    print("S", S, "reference-mode DT build s", ts, flush=True)
    g.close()
