#!/usr/bin/env python
"""Bunny config (mse 1e-3) and its certified variant (5e-4): Register time, inner-BnB kernel time and the committed
counters, median of 5 after 3 warm-ups.  GOICP_ROUND_STATS=1 adds the per-round / per-phase cycle lines on stderr.

    [BNB_TIME_CLUSTER=n] python scripts/bnb_time.py [mse ...]
"""
import importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("cuda-go-icp_b200")
G = os.path.join(ROOT, "tests", "golden")
ld = lambda n: np.fromfile(os.path.join(G, n), np.float32).reshape(-1, 3)
model, data = ld("bunny_model_s0.1_seed1234.f32"), ld("bunny_data_s0.1_seed1235.f32")
for mse in [float(a) for a in sys.argv[1:]] or [1e-3, 5e-4]:
    g = pkg.GoICP(mse); g.pModel, g.pData = model, data; g.cluster_size = int(os.environ.get("BNB_TIME_CLUSTER", "0")); g.BuildDT()
    for _ in range(3): g.Register()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); g.Register(); ts.append(time.perf_counter() - t0)
    r = g.result
    print(json.dumps({"mse": mse, "register_ms": 1e3 * float(np.median(ts)), "bnb_ms": 1e3 * r["seconds_bnb_kernels"], "icp_ms": 1e3 * r["seconds_icp"], "rounds": int(r["rounds"]),
                      "rot_pops": int(r["rot_pops"]), "trans_pops": int(r["trans_pops"]), "bound_evals": int(r["bound_evals"]), "sse": r["sse"], "exit": r["exit_path"]}), flush=True)
    g.close()
