#!/usr/bin/env python
"""Speculation width of the outer (rotation) BnB rounds and cluster size of the inner BnB, measured: Register on the bunny
config (mse 1e-3 and the certified 5e-4) for spec_cubes / cluster_size in a range.  profiles/r2c_spec_solo_cluster_sweep.jsonl
was taken with two experimental switches that were NOT kept because they did not pay: `solo` (warps
4/8/12 of the leader CTA leave the owner warp's issue port alone) and `descend` (underfilled rounds also evaluate the
children of their cubes, breadth first).  Both environment variables are ignored by the committed library."""
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
  for solo in ("0", "1"):
    os.environ["GOICP_BNB_SOLO"] = solo
    for descend in ("0", "1"):
        os.environ["GOICP_SPEC_DESCEND"] = descend
        for spec in ((37,) if solo == "0" else (18, 37, 74, 148, 296)):
            g = pkg.GoICP(mse); g.pModel, g.pData = model, data; g.spec_cubes = spec; g.SetDT(*dt)
            for _ in range(3):
                g.Register()
            ts = []
            for _ in range(5):
                t0 = time.perf_counter(); g.Register(); ts.append(time.perf_counter() - t0)
            r = g.result
            print(json.dumps({"mse": mse, "solo": int(solo), "descend": int(descend), "spec": spec, "register_ms": 1e3 * float(np.median(ts)), "bnb_ms": 1e3 * r["seconds_bnb_kernels"], "icp_ms": 1e3 * r["seconds_icp"],
                              "rounds": int(r["rounds"]), "executed": int(r["bound_evals_executed"]), "committed": int(r["bound_evals"]), "rot_pops": int(r["rot_pops"]), "trans_pops": int(r["trans_pops"]), "sse": r["sse"]}), flush=True)
            g.close()

# cluster size x owner-solo on the default width
os.environ["GOICP_SPEC_DESCEND"] = "1"
for mse in (1e-3, 5e-4):
    for solo in ("0", "1"):
        os.environ["GOICP_BNB_SOLO"] = solo
        for cl in (2, 4, 8, 16):
            g = pkg.GoICP(mse); g.pModel, g.pData = model, data; g.cluster_size = cl; g.SetDT(*dt)
            for _ in range(3):
                g.Register()
            ts = []
            for _ in range(5):
                t0 = time.perf_counter(); g.Register(); ts.append(time.perf_counter() - t0)
            r = g.result
            print(json.dumps({"mse": mse, "solo": int(solo), "cluster": cl, "register_ms": 1e3 * float(np.median(ts)), "bnb_ms": 1e3 * r["seconds_bnb_kernels"], "rounds": int(r["rounds"]),
                              "executed": int(r["bound_evals_executed"]), "rot_pops": int(r["rot_pops"]), "trans_pops": int(r["trans_pops"]), "sse": r["sse"]}), flush=True)
            g.close()
# ICP grid size
os.environ["GOICP_BNB_SOLO"] = "1"
for blocks in ("95", "148"):
    os.environ["GOICP_ICP_BLOCKS"] = blocks
    os.environ["GOICP_ICP_STATS"] = "1"
    g = pkg.GoICP(1e-3); g.pModel, g.pData = model, data; g.SetDT(*dt)
    for _ in range(3):
        g.Register()
    r = g.result
    print(json.dumps({"icp_blocks": int(blocks), "icp_ms": 1e3 * r["seconds_icp"]}), flush=True)
    g.close()
