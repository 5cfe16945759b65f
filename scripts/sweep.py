#!/usr/bin/env python
"""Synthetic scaling sweep (BASELINE.json configs[4]): data 1k-1M points, model up to 1M points,
300^3-512^3 distance transform, random SE(3) ground truth.  Each case is time-boxed (goicp_cancel
after --budget seconds); prints one JSON line per case with executed bound evaluations per second,
DT look-ups per second and the L1TEX-sector roofline fraction.

    python scripts/sweep.py --cases 1000:100000:300 10000:100000:300 100000:1000000:512 --budget 10

Multi-GPU: launch under torchrun (one rank per GPU); every rank builds the same clouds and DT, the
rotation frontier of each registration is sharded across the ranks (ncclAllGather per round, see
DESIGN.md section 7) and rank 0 prints.  The time box is off in that mode (a cancel that reaches the
ranks in different rounds would leave a collective unmatched) -- wrap the run in `timeout` instead.
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
    # goicp_build_dt would run Register's first ICP next to the DT build (include/goicp_b200.h); the sweep reports the
    # registration by itself, so that overlap is off here and `seconds` holds every ICP call
    os.environ["GOICP_NO_PREFETCH"] = "1"
    pkg = importlib.import_module("cuda-go-icp_b200")
    rank, world, local_rank = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    nccl_id = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl")
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(pkg.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)                      # plumbing only: the 128-byte ncclUniqueId
        nccl_id = bytes(idt.cpu().numpy().tobytes())
    if world > 1:                                   # first collective of a communicator sets up its connections (~1 s): not a case's time
        gold = os.path.join(ROOT, "tests", "golden")    # a registration that does go through BnB rounds (926 rotation pops)
        g = pkg.GoICP(2e-3, device=local_rank)
        g.pModel = np.fromfile(os.path.join(gold, "bunny_model_s0.033_seed1234.f32"), np.float32).reshape(-1, 3)
        g.pData = np.fromfile(os.path.join(gold, "bunny_data_s0.033_seed1235.f32"), np.float32).reshape(-1, 3)[::4].copy()
        g.dt.SIZE = 50
        g.dt_mode = 1
        g.init_nccl(nccl_id, rank, world)
        g.BuildDT(); g.Register(); g.close()
    for case in args.cases:
        nd, nm, S = (int(v) for v in case.split(":"))
        model, data, Rgt, tgt = synth(nm, nd)
        g = pkg.GoICP(args.mse, device=local_rank)
        g.pModel, g.pData = model, data
        g.dt.SIZE = S
        g.dt_mode = args.dt_mode
        if world > 1:
            g.init_nccl(nccl_id, rank, world)
        t0 = time.perf_counter(); g.BuildDT(); dt_s = time.perf_counter() - t0
        timer = threading.Timer(args.budget, g.Cancel)
        if world > 1:
            dist.barrier()
        else:
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
        if world > 1:
            tt = torch.tensor([el], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            el = float(tt.item())
        out = {"case": case, "n_gpus": world, "Nd": nd, "Nm": nm, "S": S, "dt_mode": args.dt_mode, "bound_evals_committed": res["bound_evals"],
               "rot_pops": res["rot_pops"], "icp_calls": res.get("icp_calls"), "dt_build_s": dt_s, "seconds": el, "exit_path": res["exit_path"],
               "sse": res["sse"], "sse_thresh": res["sse_thresh"], "rot_err_rad_vs_truth": ang, "t_err_vs_truth": float(np.abs(res["t"] - tgt).max()),
               "bound_evals_executed": res["bound_evals_executed"], "bound_evals_per_s": res["bound_evals_executed"] / el,
               "dt_lookups_per_s": lookups / el, "bnb_kernel_s": res["seconds_bnb_kernels"], "icp_s": res["seconds_icp"], "rounds": res["rounds"],
               "lookups_per_s_in_bnb_kernels": lookups / max(res["seconds_bnb_kernels"], 1e-9),
               "l1tex_sector_roofline_frac": lookups / max(res["seconds_bnb_kernels"], 1e-9) / (148 * 1.965e9 * world)}
        if rank == 0:
            print(json.dumps(out), flush=True)
        g.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
