import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("cuda-go-icp_b200")
G = os.path.join(ROOT, "tests", "golden")
model = np.fromfile(os.path.join(G, "bunny_model_s0.1_seed1234.f32"), np.float32).reshape(-1, 3)
data = np.fromfile(os.path.join(G, "bunny_data_s0.1_seed1235.f32"), np.float32).reshape(-1, 3)
for rep in range(int(os.environ.get('REPS', '3'))):
    t0 = time.perf_counter()
    g = pkg.GoICP(1e-3); g.pModel, g.pData = model, data
    g._handle(); t1 = time.perf_counter()
    g.BuildDT(); t2 = time.perf_counter()
    g.Register(); t3 = time.perf_counter()
    r = g.result
    g.close(); t4 = time.perf_counter()
    print("rep %d: create %.4f  build_dt %.4f  register %.4f (internal %.4f, bnb %.4f, icp %.4f)  destroy %.4f" % (rep, t1 - t0, t2 - t1, t3 - t2, r["seconds_total"], r["seconds_bnb_kernels"], r["seconds_icp"], t4 - t3), flush=True)
