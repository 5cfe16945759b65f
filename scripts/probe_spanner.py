#!/usr/bin/env python
"""How long is BASELINE config 4 as its TOML writes it (test/spanner_goicp.toml: mse 1e-4)?  Runs the committed spanner
fixtures at mse 1e-4 with trim 0 / 0.1 under a time box and prints the counters (sizes the CPU reference run)."""
import importlib, json, os, sys, threading, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("cuda-go-icp_b200")
G = os.path.join(ROOT, "tests", "golden")
ld = lambda n: np.fromfile(os.path.join(G, n), np.float32).reshape(-1, 3)
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
for mse in (3e-4, 1e-4):
    for trim in (0.0, 0.1):
        g = pkg.GoICP(mse)
        g.pModel, g.pData = ld("spanner_model_noisy_flipped_s0.02_seed1234.f32"), ld("spanner_data_rotated_s0.02_seed1235.f32")
        g.trimFraction = trim
        g.initNodeTrans = [-1.0, -1.0, -1.0, 2.0]
        g.BuildDT()
        t = threading.Timer(budget, g.Cancel); t.start()
        t0 = time.perf_counter()
        try:
            g.Register()
        except pkg.GoicpError as e:
            pass
        t.cancel()
        r = g.result
        print(json.dumps({"mse": mse, "trim": trim, "seconds": time.perf_counter() - t0, "exit": r["exit_path"], "sse": r["sse"], "sse_thresh": r["sse_thresh"], "rot_pops": int(r["rot_pops"]),
                          "trans_pops": int(r["trans_pops"]), "bound_evals": int(r["bound_evals"]), "executed": int(r["bound_evals_executed"]), "icp_calls": int(r["icp_calls"])}), flush=True)
        g.close()
