#!/usr/bin/env python
"""Synthetic scaling sweep (BASELINE.json configs[4]): data 1k-1M points, model up to 1M points,
300^3-512^3 distance transform, random SE(3) ground truth.  Each case is time-boxed (goicp_cancel
after --budget seconds); prints one JSON line per case with executed bound evaluations per second,
DT look-ups per second and the L1TEX-sector roofline fraction.

    python scripts/sweep.py --cases 1000:100000:300 10000:100000:300 100000:1000000:512 --budget 10
"""
import argparse, importlib, json, os, sys, threading, time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def synth(nm, nd, seed_model=1, seed_data=2, seed_pose=3, sigma=1e-3):
    """closed star-shaped surface r(u) = 0.33 + low-order bumps, scaled into [-0.5,0.5]^3; data = noisy
    subset moved by the inverse of a random rigid motion (rotation uniform on SO(3), |t|_inf <= 0.3)."""
    rng = np.random.default_rng(seed_model)
    u = rng.normal(size=(nm, 3)); u /= np.linalg.norm(u, axis=1, keepdims=True)
    k = rng.normal(size=(6, 3)); ph = rng.uniform(0, 2 * np.pi, 6); amp = rng.uniform(0.02, 0.06, 6)
    r = 0.33 + sum(a * np.sin(3 * (u @ kk) + p) for a, kk, p in zip(amp, k, ph))
    model = (u * r[:, None])
    model *= 0.5 / np.abs(model).max()
    rd = np.random.default_rng(seed_data)
    pick = rd.choice(nm, nd, replace=nd > nm)
    pts = model[pick] + rd.normal(scale=sigma, size=(nd, 3))
    rp = np.random.default_rng(seed_pose)
    q = rp.normal(size=4); q /= np.linalg.norm(q)
    w, x, y, z = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    t = rp.uniform(-0.3, 0.3, 3)
    data = (pts - t) @ R            # data = R^T (p - t)  =>  R data + t = p
    return model.astype(np.float32), data.astype(np.float32), R.astype(np.float32), t.astype(np.float32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", nargs="+", default=["1000:100000:300", "10000:100000:300"])
    ap.add_argument("--budget", type=float, default=10.0)
    ap.add_argument("--mse", type=float, default=1e-4)
    ap.add_argument("--dt-mode", type=int, default=1, help="0 reference-exact propagation, 1 exact EDT (default for the sweep)")
    args = ap.parse_args()
    pkg = importlib.import_module("cuda-go-icp_b200")
    for case in args.cases:
        nd, nm, S = (int(v) for v in case.split(":"))
        model, data, Rgt, tgt = synth(nm, nd)
        g = pkg.GoICP(args.mse)
        g.pModel, g.pData = model, data
        g.dt.SIZE = S
        g.dt_mode = args.dt_mode
        t0 = time.perf_counter(); g.BuildDT(); dt_s = time.perf_counter() - t0
        timer = threading.Timer(args.budget, g.Cancel)
        timer.start()
        t0 = time.perf_counter()
        try:
            g.Register()
            res = g.result
        except pkg.GoicpError as e:
            if e.code != 6:
                raise
            res = g.last_result()
        el = time.perf_counter() - t0
        timer.cancel()
        ang = float(np.linalg.norm(res["R"] - Rgt) / np.sqrt(2))
        lookups = res["bound_evals_executed"] * nd
        out = {"case": case, "Nd": nd, "Nm": nm, "S": S, "dt_mode": args.dt_mode, "dt_build_s": dt_s, "seconds": el, "exit_path": res["exit_path"],
               "sse": res["sse"], "sse_thresh": res["sse_thresh"], "rot_err_rad_vs_truth": ang, "t_err_vs_truth": float(np.abs(res["t"] - tgt).max()),
               "bound_evals_executed": res["bound_evals_executed"], "bound_evals_per_s": res["bound_evals_executed"] / el,
               "dt_lookups_per_s": lookups / el, "bnb_kernel_s": res["seconds_bnb_kernels"], "icp_s": res["seconds_icp"], "rounds": res["rounds"],
               "lookups_per_s_in_bnb_kernels": lookups / max(res["seconds_bnb_kernels"], 1e-9),
               "l1tex_sector_roofline_frac": lookups / max(res["seconds_bnb_kernels"], 1e-9) / (148 * 1.965e9)}
        print(json.dumps(out), flush=True)
        g.close()


if __name__ == "__main__":
    main()
