#!/usr/bin/env python
"""Gather warps of the leader CTA x cluster size (incl. 3, 5, 6), bunny config mse 1e-3 and 5e-4.  profiles/r2i_leader_warps_cluster_sweep.jsonl
was taken with an experimental kernel parameter (GOICP_BNB_LEADER_WARPS: 0 = owner-only leader CTA, its L1TEX idle) that was NOT
kept: taking gathers away from the owner's SM never paid.  The committed library ignores the variable."""
import importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("cuda-go-icp_b200")
G = os.path.join(ROOT, "tests", "golden")
ld = lambda n: np.fromfile(os.path.join(G, n), np.float32).reshape(-1, 3)
model, data = ld("bunny_model_s0.1_seed1234.f32"), ld("bunny_data_s0.1_seed1235.f32")
base = pkg.GoICP(1e-3); base.pModel, base.pData = model, data; base.BuildDT(); dt = base.GetDT(); base.close()
for mse in (1e-3, 5e-4):
    for lw in ("15", "8", "4", "0"):
        os.environ["GOICP_BNB_LEADER_WARPS"] = lw
        for cl in (2, 3, 4, 5, 6, 8):
            g = pkg.GoICP(mse); g.pModel, g.pData = model, data; g.cluster_size = cl; g.SetDT(*dt)
            try:
                for _ in range(2):
                    g.Register()
                ts = []
                for _ in range(4):
                    t0 = time.perf_counter(); g.Register(); ts.append(time.perf_counter() - t0)
                r = g.result
                print(json.dumps({"mse": mse, "leader_warps": int(lw), "cluster": cl, "register_ms": 1e3 * float(np.median(ts)), "bnb_ms": 1e3 * r["seconds_bnb_kernels"], "rounds": int(r["rounds"]),
                                  "executed": int(r["bound_evals_executed"]), "rot_pops": int(r["rot_pops"]), "trans_pops": int(r["trans_pops"]), "sse": r["sse"]}), flush=True)
            except Exception as e:
                print(json.dumps({"mse": mse, "leader_warps": int(lw), "cluster": cl, "error": str(e)[:200]}), flush=True)
            g.close()
