#!/usr/bin/env python
"""Strict vs tolerance numerics, measured: every golden registration of tests/golden/goicp_runs.json (the reference's
own runs, oracle/_ref) plus BASELINE config 2 (ICP only) through the C ABI in each numerics mode, tabulated against
the reference: rotation error (rad), translation error, relative SSE error, exit path, search counters, seconds.

    python scripts/parity_modes.py [--modes strict fastdt ...] [--out gpurun_out/parity_modes]

Modes = (dt_mode, numerics bitmask) pairs of include/goicp_b200.h:
    strict      reference-order DT propagation, reference-order sums wherever a decision hinges on them, reference ICP arithmetic
    fastdt      exact EDT (+ the reference binary's corner seed), everything else strict
    fastdt_icp  + ICP moments by parallel reduction and Jacobi SVD
    fast        + no reference-order re-evaluation of near-tied upper bounds / DT scores (tree sums decide)
    fasticp     reference DT, fast ICP only          fastsums   reference DT, fast sums only
    jacobi_svd  exact EDT, reference-order sums everywhere, but the ICP rotation from the engine's Jacobi solver instead of Matrix::svd

North-star tolerances: R 1e-4 rad, t 1e-4, SSE 1e-5 relative, same exit path.  Voxel and NN indices are bit-exact in
every mode (the index arithmetic and the tree search are the same code).
"""
import argparse, importlib, json, os, sys, time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

MODES = {  # name: (dt_mode, numerics)
    "strict": (0, 0), "fastdt": (2, 0), "fasticp": (0, 2), "fastsums": (0, 1), "fastdt_icp": (2, 2), "fast": (2, 3), "edt_pure": (1, 0), "jacobi_svd": (2, 4),
}
TOL_R, TOL_T, TOL_SSE = 1e-4, 1e-4, 1e-5


def rot_angle(Ra, Rb):
    d = np.linalg.norm(np.asarray(Ra, np.float64) - np.asarray(Rb, np.float64))
    return float(2 * np.arcsin(min(1.0, d / (2 * np.sqrt(2)))))


def load(name):
    return np.fromfile(os.path.join(GOLDEN, name), np.float32).reshape(-1, 3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--modes", nargs="+", default=["strict", "fastdt", "fasticp", "fastsums", "fastdt_icp", "fast"])
    ap.add_argument("--runs", nargs="*", default=None)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "parity_modes"))
    args = ap.parse_args()
    pkg = importlib.import_module("cuda-go-icp_b200")
    runs = json.load(open(os.path.join(GOLDEN, "goicp_runs.json")))
    names = args.runs or list(runs)
    rows = []
    dt_cache = {}
    for mode in args.modes:
        dt_mode, numerics = MODES[mode]
        for name in names:
            gold = runs[name]
            model, data = load(gold["model"]), load(gold["data"])
            g = pkg.GoICP(gold["mse"])
            g.pModel, g.pData = model, data
            g.trimFraction = gold["trim"]
            g.dt_mode = dt_mode
            g.numerics = numerics
            if "trans_cube" in gold:
                g.initNodeTrans = gold["trans_cube"]
            key = (gold["model"], dt_mode)
            t0 = time.perf_counter()
            if key in dt_cache:
                g.SetDT(*dt_cache[key]); dt_s = None
            else:
                os.environ["GOICP_NO_PREFETCH"] = "1"
                g.BuildDT(); dt_s = time.perf_counter() - t0
                dt_cache[key] = g.GetDT()
            t1 = time.perf_counter()
            g.Register()
            reg_s = time.perf_counter() - t1
            r = g.result
            g.close()
            dR = rot_angle(r["R"], np.array(gold["R"]).reshape(3, 3))
            dT = float(np.abs(r["t"] - np.array(gold["t"])).max())
            dS = abs(r["sse"] - gold["sse"]) / gold["sse"]
            row = {"mode": mode, "run": name, "dR_rad": dR, "dt": dT, "dSSE_rel": dS, "exit": r["exit_path"], "exit_ref": gold["exit_path"],
                   "rot_pops": int(r["rot_pops"]), "rot_pops_ref": gold["rot_pops"], "trans_pops": int(r["trans_pops"]), "trans_pops_ref": gold["trans_pops"],
                   "icp_calls": int(r["icp_calls"]), "sse": r["sse"], "sse_ref": gold["sse"], "dt_build_s": dt_s, "register_s": reg_s,
                   "seconds_icp": r["seconds_icp"], "seconds_bnb": r["seconds_bnb_kernels"],
                   "within_tol": bool(dR < TOL_R and dT < TOL_T and dS < TOL_SSE and r["exit_path"] == gold["exit_path"])}
            rows.append(row)
            print(json.dumps(row), flush=True)
        # DT values against the reference-order grid of the bunny model (same frame): how many voxels differ, by how much
        if dt_mode != 0 and (runs[names[0]]["model"], 0) in dt_cache:
            a, _ = dt_cache[(runs[names[0]]["model"], 0)]
            b, meta = dt_cache[(runs[names[0]]["model"], dt_mode)]
            diff = a != b
            row = {"mode": mode, "run": "dt_grid_vs_reference:" + runs[names[0]]["model"], "voxels_differing": int(diff.sum()), "fraction": float(diff.mean()),
                   "max_abs_diff_voxels": float(np.abs(a - b).max() * meta[3])}
            rows.append(row); print(json.dumps(row), flush=True)
        # BASELINE config 2: ICP only, full size
        c2 = dict(np.load(os.path.join(GOLDEN, "bun_icp_config2.npz")))
        g = pkg.GoICP(1e-5)
        g.pModel, g.pData = c2["model"], c2["data"]
        g.numerics = numerics
        idx, d2 = g.NN(c2["data"])
        t1 = time.perf_counter()
        err, R, t, iters = g.ICP(np.eye(3), np.zeros(3), 10000, 1e-9)
        icp_s = time.perf_counter() - t1
        g.close()
        row = {"mode": mode, "run": "config2_icp_only", "dR_rad": rot_angle(R, c2["icp_R"]), "dt": float(np.abs(t - c2["icp_t"]).max()),
               "dSSE_rel": abs(err - float(c2["icp_err"])) / float(c2["icp_err"]), "nn_idx_equal": bool(np.array_equal(idx, c2["nn_idx"])),
               "iterations": int(iters), "icp_s": icp_s}
        row["within_tol"] = bool(row["dR_rad"] < TOL_R and row["dt"] < TOL_T and row["dSSE_rel"] < TOL_SSE and row["nn_idx_equal"])
        rows.append(row); print(json.dumps(row), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out + ".jsonl", "w") as f:
        for r in rows:
            f.write(json.dumps(r) + "\n")
    with open(args.out + ".md", "w") as f:
        f.write("| mode | run | dR (rad) | dt | dSSE/SSE | exit (ref) | rot pops (ref) | trans pops (ref) | ICP calls | register s | within 1e-4/1e-4/1e-5 |\n|---|---|---|---|---|---|---|---|---|---|---|\n")
        for r in rows:
            if "exit" in r:
                f.write(f"| {r['mode']} | {r['run']} | {r['dR_rad']:.2e} | {r['dt']:.2e} | {r['dSSE_rel']:.2e} | {r['exit']} ({r['exit_ref']}) | {r['rot_pops']} ({r['rot_pops_ref']}) | "
                        f"{r['trans_pops']} ({r['trans_pops_ref']}) | {r['icp_calls']} | {r['register_s']:.4f} | {'yes' if r['within_tol'] else 'NO'} |\n")
            elif r["run"] == "config2_icp_only":
                f.write(f"| {r['mode']} | config 2 ICP only (NN idx equal: {r['nn_idx_equal']}, {r['iterations']} it) | {r['dR_rad']:.2e} | {r['dt']:.2e} | {r['dSSE_rel']:.2e} | - | - | - | - | {r['icp_s']:.4f} | {'yes' if r['within_tol'] else 'NO'} |\n")
            else:
                f.write(f"| {r['mode']} | {r['run']} | {r['voxels_differing']} voxels differ ({r['fraction']:.2e}), max {r['max_abs_diff_voxels']:.3f} voxel | | | | | | | | |\n")
    bad = [r for r in rows if r.get("within_tol") is False]
    print(f"{len(rows)} rows, {len(bad)} outside tolerance", file=sys.stderr)


if __name__ == "__main__":
    main()
