import importlib, os, sys, time
import numpy as np
sys.path.insert(0, '/root/repo')
import torch
pkg = importlib.import_module("cuda-go-icp_b200")
G = '/root/repo/tests/golden'
model = np.fromfile(os.path.join(G, "bunny_model_s0.1_seed1234.f32"), np.float32).reshape(-1, 3)
data = np.fromfile(os.path.join(G, "bunny_data_s0.1_seed1235.f32"), np.float32).reshape(-1, 3)
def one(tag, pre=None):
    for rep in range(4):
        if pre: pre()
        t0 = time.perf_counter()
        g = pkg.GoICP(1e-3); g.pModel, g.pData = model, data
        g.BuildDT(); t2 = time.perf_counter()
        g.Register(); t3 = time.perf_counter()
        g.close(); t4 = time.perf_counter()
        print("%s rep %d: build_dt %.4f register %.4f destroy %.4f" % (tag, rep, t2 - t0, t3 - t2, t4 - t3), flush=True)
one("plain")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def fl():
    flush.zero_(); torch.cuda.synchronize()
one("torch-flush", fl)
eng = pkg.GoICP(1e-3); eng.pModel, eng.pData = model, data; eng.BuildDT(); eng.Register()
one("resident-engine-alive", fl)
for _ in range(5): eng.Register()
one("after-5-registers", fl)
