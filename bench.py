#!/usr/bin/env python
"""bench.py -- Go-ICP hot-path benchmark (BASELINE.json metric: bound-evals/s + time-to-optimum).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" = one full Go-ICP registration (GoICP::Register: nested rotation/translation branch-and-bound + ICP
refinements) of the workload's data cloud against its model cloud.  Every key means what it says:

  value     committed bound evaluations per second of goicp_register on a warm handle -- model, data, kd-tree and DT
            resident in HBM when the timed region starts.  "Committed" = the evaluations the sequential reference
            performs too; speculative extra work is not credited.
  e2e       the same metric through the C ABI from HOST buffers, in the library's default configuration: goicp_create +
            set_model + set_data (H2D) + build_dt + register + result read-back, all inside the timed region; the byte
            counts are the library's own (goicp_transfer_bytes).  `e2e_reference_dt` = the same with the reference-order
            DT propagation (GOICP_DT_REFERENCE) instead of the default exact EDT.
  certified the same clouds at mse 5e-4, where the search ends through the global-optimality certificate
            (time-to-CERTIFIED-optimum; the TOML's own 1e-3 ends through `optError < SSEThresh`).
  roofline  dominant kernel = the persistent translation BnB.  achieved = DT look-ups this GPU executed * 32 B (one
            sector per scattered 4-byte gather) / the kernel time on the launching stream (CUDA events).
            peak, for a grid that fits L2 = the larger of two gather rates MEASURED in the same run -- uniformly random
            4-byte loads over a buffer of the grid's size and nothing else (goicp_measure_gather), and the bound
            evaluation's own access pattern at full occupancy (expand_bounds_kernel) -- so no fraction exceeds 1; for a
            grid beyond L2 = MEASURED_PEAKS.json's HBM copy rate.  `traffic` comes from the committed ncu capture named
            in `traffic_source`.
  cpu_baseline  the unmodified reference (oracle/_ref) or its C restatement on ONE host core (the reference is
            single-threaded) for a bounded sample of the same workload.

--impl reference runs only that CPU arm: `value` = its Register-only rate (the span of our `value`), `e2e` = DT build +
Register (the span of our `e2e`).  Multi-GPU (torchrun, N>1): the rotation frontier of ONE registration is sharded across
the ranks, results all-gathered by NCCL on the engine stream -- strong scaling; roofline numbers are rank 0's own GPU.
"""
from __future__ import annotations

import argparse
import ctypes
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

WORKLOADS = {
    # BASELINE.json configs[0]: test/bunny_goicp.toml (subsample 0.1, mse 1e-3, S=300, trim 0),
    # clouds = the committed deterministic subsamples (seeds 1234/1235) of the reference's bunny.
    "bunny_goicp_toml": dict(model="bunny_model_s0.1_seed1234.f32", data="bunny_data_s0.1_seed1235.f32", mse=1e-3, S=300,
                             ref_register_s=47.5, ref_evals=235552, certified_mse=5e-4),
    # same clouds, tighter threshold: exits through the global-optimality certificate
    "bunny_goicp_certified": dict(model="bunny_model_s0.1_seed1234.f32", data="bunny_data_s0.1_seed1235.f32", mse=5e-4, S=300,
                                  ref_register_s=338.6, ref_evals=1696656),
    # BASELINE.json configs[4]: synthetic sweep point -- throughput-bound regime (one bound evaluation = 1e5 look-ups into a
    # 537 MB grid that does not fit L2).  Seeds fixed: model 1, data 2, pose 3.
    # BASELINE.json configs[3]: test/spanner_goicp.toml as written (mse 1e-4, translation cube [-1,1]^3; noisy + flipped model
    # against a rotated copy): a throughput-bound search -- 5 305 rotation pops, 19.3 M translation pops, 154 M bound
    # evaluations to the certificate; 17 034 s (4.7 h) of the unmodified reference on one core of the build container.
    "spanner_goicp_toml": dict(model="spanner_model_noisy_flipped_s0.02_seed1234.f32", data="spanner_data_rotated_s0.02_seed1235.f32", mse=1e-4, S=300,
                               trans_cube=[-1.0, -1.0, -1.0, 2.0], ref_register_s=17034.4, ref_evals=153867445),
    "sweep_100k": dict(synth=dict(nm=1_000_000, nd=100_000), mse=1e-4, S=512, ref_register_s=None, ref_evals=None),
    "sweep_10k": dict(synth=dict(nm=100_000, nd=10_000), mse=1e-4, S=300, ref_register_s=None, ref_evals=None),
}
KERNEL_NAMES = {1: "inner_bnb_pipelined_kernel<1,1,512,1>", 2: "inner_bnb_pipelined_kernel<1,0,512,2>", 4: "inner_bnb_pipelined_kernel<0,1,512,1>",
                8: "inner_bnb_pipelined_kernel<0,0,512,2>", 16: "inner_bnb_kernel", 32: "inner_bnb_pipelined_kernel<1,0,192,5>", 64: "inner_bnb_pipelined_kernel<0,0,192,5>"}
L2_BYTES = 126 * 1000 * 1000


def synth(nm, nd, seed_model=1, seed_data=2, seed_pose=3, sigma=1e-3):
    """closed star-shaped surface r(u) = 0.33 + low-order bumps, scaled into [-0.5,0.5]^3; data = noisy subset moved by
    the inverse of a random rigid motion (rotation uniform on SO(3), |t|_inf <= 0.3)  (SURVEY.md 8d, config 5)."""
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


def load(name):
    return np.fromfile(os.path.join(GOLDEN, name), np.float32).reshape(-1, 3)


def clouds_of(wl):
    if "synth" in wl:
        m, d, _, _ = synth(wl["synth"]["nm"], wl["synth"]["nd"])
        return m, d
    return load(wl["model"]), load(wl["data"])


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.samples.append([x.strip() for x in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            # nvidia-smi polling holds the driver lock now and then: wait for it to be gone before the e2e passes
            self.proc.terminate()
            try:
                self.proc.wait(timeout=10)
            except Exception:
                self.proc.kill()
        self.join(5)
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        reasons = set()
        for s in self.samples:
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], s[3:7]):
                if v == "Active":
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


class _Quiet:
    """the reference narrates on stdout: keep bench.py's single JSON line clean"""

    def __enter__(self):
        self.devnull = os.open(os.devnull, os.O_WRONLY)
        self.saved = os.dup(1)
        sys.stdout.flush()
        os.dup2(self.devnull, 1)

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.devnull)


def cpu_reference_sample(wl, seconds, model=None, data=None):
    """Bound evaluations per second of the reference CPU Go-ICP (1 thread) over ~`seconds` of its Register() on this
    workload.  Uses oracle/_ref (unmodified reference) when built, else the C restatement.  The DT is built first (timed
    separately, not part of the rate)."""
    from oracle import oracle as orc
    if model is None:
        model, data = clouds_of(wl)
    if orc.Reference.available():
        rf = orc.Reference()
        g = rf.create(model, data, wl["mse"], 0.0, wl["S"], trans_cube=wl.get("trans_cube"))
        with _Quiet():
            dt_s = rf.build_dt(g)
            counter = ctypes.c_longlong.in_dll(rf.L, "ref_select_calls")
            stop = ctypes.c_bool.in_dll(rf.L, "goicp_finished")
            stop.value = False
            th = threading.Thread(target=rf.register, args=(g,), daemon=True)
            th.start()
            time.sleep(min(1.0, seconds / 10))           # past Initialize (kd-tree build) and the first ICP
            c0, t0 = counter.value, time.perf_counter()
            th.join(seconds)
            c1, t1 = counter.value, time.perf_counter()
            stop.value = True                   # the reference's own cooperative exit (jly_goicp.cpp:400)
            th.join(600)
            stop.value = False
        return {"value": (c1 - c0) / max(t1 - t0, 1e-9), "unit": "bound-evals/s", "cores": 1, "kind": "reference",
                "sample": f"{t1 - t0:.1f} s of GoICP::Register on Nd={len(data)}, Nm={len(model)}, S={wl['S']} (after its first second); DT build excluded ({dt_s:.1f} s)",
                "dt_build_s": dt_s}
    rs = orc.Restated()
    g = rs.create(model, data, wl["mse"], 0.0, wl["S"], trans_cube=wl.get("trans_cube"))
    t0 = time.perf_counter(); rs.L.go_build_dt(g); dt_s = time.perf_counter() - t0
    rs.L.go_set_budget(g, float(seconds))
    r = rs.register(g)
    return {"value": r["bound_evals"] / r["register_s"], "unit": "bound-evals/s", "cores": 1, "kind": "port",
            "sample": f"first {r['register_s']:.1f} s of the restated Register on Nd={len(data)}, Nm={len(model)}, S={wl['S']}; DT build excluded ({dt_s:.1f} s)",
            "dt_build_s": dt_s}


def cpu_reference_job(wl):
    """One whole job of the reference CPU Go-ICP on this workload: DT build + full Register, timed separately."""
    from oracle import oracle as orc
    model, data = clouds_of(wl)
    if orc.Reference.available():
        rf = orc.Reference()
        g = rf.create(model, data, wl["mse"], 0.0, wl["S"], trans_cube=wl.get("trans_cube"))
        with _Quiet():
            dt_s = rf.build_dt(g)
            counter = ctypes.c_longlong.in_dll(rf.L, "ref_select_calls")
            c0, t0 = counter.value, time.perf_counter()
            rf.register(g)
            c1, t1 = counter.value, time.perf_counter()
        return {"kind": "reference", "dt_build_s": dt_s, "register_s": t1 - t0, "job_s": dt_s + (t1 - t0),
                "register_rate": wl["ref_evals"] / (t1 - t0), "select_calls": c1 - c0}
    rs = orc.Restated()
    g = rs.create(model, data, wl["mse"], 0.0, wl["S"], trans_cube=wl.get("trans_cube"))
    t0 = time.perf_counter(); rs.L.go_build_dt(g); dt_s = time.perf_counter() - t0
    r = rs.register(g)
    return {"kind": "port", "dt_build_s": dt_s, "register_s": r["register_s"], "job_s": dt_s + r["register_s"],
            "register_rate": r["bound_evals"] / r["register_s"], "select_calls": r["bound_evals"]}


def ncu_traffic(workload, kernel_prefix):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed summary of an
    `ncu --set full` capture of this workload (profiles/ncu_traffic.json, written by scripts/summarise_profiles.py)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        e = t[workload]
        if e["kernel"].startswith(kernel_prefix.split("<")[0]):
            return e["dram_bytes_per_launch"], e["source"]
    except Exception:
        pass
    return None, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="bunny_goicp_toml", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the certified / reference-DT / gather side measurements")
    ap.add_argument("--e2e-samples", type=int, default=0, help="timed end-to-end passes (default: 3..5 by --steps); the passes are whole jobs, so long workloads may want fewer")
    ap.add_argument("--numerics", type=int, default=0, help="goicp_numerics flags of the engine (0 = strict, the library default; 3 = tree sums + parallel ICP moments)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": args.workload, "model": wl.get("model", "synthetic closed surface, seed 1"), "data": wl.get("data", "noisy subset under a random SE(3), seeds 2/3"),
              "Nd": None, "Nm": None, "dt_size": wl["S"], "mse_threshold": wl["mse"], "trim": 0.0,
              "dt_mode": "exact EDT + reference corner seed (library default)",
              "numerics": "strict (reference-order sums, reference ICP arithmetic)" if args.numerics == 0 else f"goicp_numerics flags {args.numerics} (tolerance mode: see profiles/r2_parity_modes.md)",
              "l2": "flushed between timed steps (256 MiB write)"}

    # ------------------------------------------------------------------ reference arm ---------
    if args.impl == "reference":
        if rank != 0:
            return
        # One step = the whole job the GPU arm's e2e times: DT build + Register of the workload, by the reference's own CPU
        # code (1 thread: src/goicp has no threading).  A full bunny job is ~30 s on the B200 host, so steps are capped to a
        # ~200 s budget (at least one); a workload whose full job does not fit is sampled for --cpu-seconds of Register.
        model, data = clouds_of(wl)
        config.update(Nd=len(data), Nm=len(model), dt_mode="reference (DT3D::Build)", numerics="reference")
        budget_s, steps, t_begin = 200.0, [], time.perf_counter()
        full = wl["ref_register_s"] is not None and wl["ref_register_s"] < 120.0
        while len(steps) < max(1, args.steps):
            st = cpu_reference_job(wl) if full else cpu_reference_sample(wl, args.cpu_seconds, model, data)
            steps.append(st)
            spent = time.perf_counter() - t_begin
            if spent + spent / len(steps) > budget_s:
                break
        if full:
            job_s = float(np.mean([s["job_s"] for s in steps]))
            reg_s = float(np.mean([s["register_s"] for s in steps]))
            reg_rate = float(np.mean([s["register_rate"] for s in steps]))
            e2e_v = wl["ref_evals"] / job_s
            sample = (f"{len(steps)} full job(s): DT build {np.mean([s['dt_build_s'] for s in steps]):.1f} s + GoICP::Register "
                      f"{reg_s:.1f} s, {wl['ref_evals']} bound evaluations each")
            span = "measured"
        else:
            reg_rate = float(np.mean([s["value"] for s in steps]))
            dt_s = float(np.mean([s["dt_build_s"] for s in steps]))
            if wl["ref_evals"]:
                reg_s = wl["ref_evals"] / reg_rate
                job_s = dt_s + reg_s
                e2e_v = wl["ref_evals"] / job_s
                sample = steps[-1]["sample"] + f"; whole job derived: DT {dt_s:.1f} s + {wl['ref_evals']} evals / measured rate"
                span = "derived from a bounded sample"
            else:
                reg_s, job_s, e2e_v = None, None, reg_rate
                sample = steps[-1]["sample"] + "; the reference cannot finish this workload: rate only"
                span = "rate of a bounded sample (no whole job)"
        base = {"value": reg_rate, "unit": "bound-evals/s", "cores": 1, "kind": steps[-1]["kind"], "sample": sample,
                "span": "GoICP::Register only (the span of the GPU arm's `value`); e2e = DT build + Register, " + span,
                "register_seconds": reg_s, "job_seconds": job_s}
        print(json.dumps({"impl": "reference", "metric": "goicp_bound_evals_per_sec", "value": reg_rate, "unit": "bound-evals/s",
                          "n_gpus": 0, "steps": len(steps), "warmup": 0, "ms_per_step": None if reg_s is None else 1e3 * reg_s,
                          "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                          "data": "the reference's own clouds, deterministic subsample (committed fixtures)" if "synth" not in wl else "synthetic (seeded)",
                          "config": config, "time_to_optimum_s": reg_s, "cpu_baseline": base,
                          "e2e": {"value": e2e_v, "unit": "bound-evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                                  "seconds_per_step": job_s}}))
        return

    # ------------------------------------------------------------------ our arm ---------------
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl")
    pkg = importlib.import_module("cuda-go-icp_b200")
    model, data = clouds_of(wl)
    config.update(Nd=len(data), Nm=len(model))

    nccl_id = None
    if world > 1:                                       # rank 0's ncclUniqueId travels over torch.distributed (plumbing only)
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(pkg.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        nccl_id = bytes(idt.cpu().numpy().tobytes())

    def make_engine(mse=None, dt_mode=None, wl=wl, clouds=None):
        g = pkg.GoICP(wl["mse"] if mse is None else mse, device=local_rank)
        g.pModel, g.pData = clouds if clouds is not None else (model, data)
        g.dt.SIZE = wl["S"]
        g.numerics = args.numerics
        if "trans_cube" in wl:
            g.initNodeTrans = wl["trans_cube"]
        if dt_mode is not None:
            g.dt_mode = dt_mode
        if world > 1 and os.environ.get("GOICP_EXCHANGE", "nccl") == "nccl":
            g.init_nccl(nccl_id, rank, world)          # native: ncclAllGather on the engine's stream (include/goicp_b200.h)
        elif world > 1:
            send_t = {}

            def allgather(send):
                n = send.size
                if n not in send_t:
                    send_t[n] = (torch.empty(n, dtype=torch.uint8, device="cuda"), torch.empty(n * world, dtype=torch.uint8, device="cuda"))
                s, r = send_t[n]
                s.copy_(torch.from_numpy(send.copy()))
                dist.all_gather_into_tensor(r, s)
                return r.cpu().numpy()
            g.set_exchange(allgather, rank, world)
        return g

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()

    def timed_registers(eng, steps, warm):
        """`steps` timed goicp_register calls on a warm handle, L2 flushed before each; returns (max-over-ranks seconds, results)"""
        for _ in range(warm):
            eng.Register()
        barrier(); torch.cuda.synchronize()
        secs, results = 0.0, []
        for _ in range(steps):
            flush.zero_(); torch.cuda.synchronize()
            barrier()
            t0 = time.perf_counter()
            eng.Register()
            torch.cuda.synchronize()
            secs += time.perf_counter() - t0
            results.append(eng.result)
        t = torch.tensor([secs], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), results

    def timed_e2e(n, warm, mse=None, dt_mode=None):
        """create + H2D + build_dt + register + read-back through the C ABI from host buffers; median seconds, parts, bytes, evals"""
        samples, parts, nbytes, evals, exit_path = [], [], (0, 0), 0, None
        for it in range(n + warm):
            flush.zero_(); torch.cuda.synchronize()
            barrier()
            t0 = time.perf_counter()
            g = make_engine(mse, dt_mode)
            g.BuildDT()
            t1 = time.perf_counter()
            g.Register()
            _ = (g.optR.copy(), g.optT.copy(), g.optError)
            t2 = time.perf_counter()
            if it >= warm:
                samples.append(t2 - t0); parts.append((t1 - t0, t2 - t1))
            nbytes, evals, exit_path = g.TransferBytes(), g.result["bound_evals"], g.result["exit_path"]
            g.close()
        med = float(np.median(samples))
        return {"value": evals / med, "unit": "bound-evals/s", "h2d_bytes_per_step": int(nbytes[0]), "d2h_bytes_per_step": int(nbytes[1]),
                "bytes_source": "goicp_transfer_bytes (every copy the library issued for the step)",
                "seconds_per_step": med, "seconds_per_step_samples": [round(x, 5) for x in samples], "statistic": "median",
                "seconds_create_h2d_dt_build": float(np.median([p[0] for p in parts])),
                "seconds_register_and_readback": float(np.median([p[1] for p in parts])), "exit_path": exit_path, "bound_evals": int(evals)}

    # ---- resident engine (library defaults), timed Register() steps
    eng = make_engine()
    t0 = time.perf_counter(); eng.BuildDT(); torch.cuda.synchronize(); dt_build_s = time.perf_counter() - t0
    warm = max(3, args.warmup)
    for _ in range(warm):
        eng.Register()
    sampler = ClockSampler(local_rank); sampler.start()
    total_s, results = timed_registers(eng, args.steps, 0)
    clocks = sampler.finish()
    res = results[-1]
    evals = sum(r["bound_evals"] for r in results)
    executed = sum(r["bound_evals_executed"] for r in results)
    executed_local = sum(r["bound_evals_executed_local"] for r in results)
    kern_s = sum(r["seconds_bnb_kernels"] for r in results)
    launches = sum(r["kernel_launches"] for r in results)
    rounds = int(sum(r["rounds"] for r in results))
    value = evals / total_s

    # ---- e2e through the C ABI from host buffers, default configuration
    n_e2e = args.e2e_samples if args.e2e_samples > 0 else max(3, min(args.steps, 5))
    e2e = timed_e2e(n_e2e, 2 if args.e2e_samples == 0 else 1)
    e2e["includes"] = "create + H2D clouds + GPU DT build (library default: exact EDT with the reference's corner seed) + Register + result D2H"
    out_extra = {}
    S = wl["S"]
    if not args.no_extras:
        if S <= 640:
            # the reference-order DT propagation (bit-exact DT values; one CTA, 4*S^2 dependent row steps) instead of the exact EDT
            ed = timed_e2e(3, 2, dt_mode=0)            # untimed passes: the 1-CTA kernel only reaches steady speed after ~2 s of activity
            ed["includes"] = "as e2e, with dt_mode = GOICP_DT_REFERENCE (Register's first ICP runs next to the 1-CTA propagation)"
            dt_bytes = (4 * 2 + 2) * 8.0 * S ** 3          # 4 sweeps read+write every 8-byte working voxel, two of them read the adjacent slice once more
            ed["dt_kernel"] = {"kernel": "dt_propagate_split_kernel", "algorithmic_bytes": dt_bytes,
                               "achieved_gbs": dt_bytes / ed["seconds_create_h2d_dt_build"] / 1e9,
                               "bound": "latency: 4*S^2 row steps, each after the previous one, on ONE CTA (DESIGN.md section 5); not a bandwidth kernel"}
            out_extra["e2e_reference_dt"] = ed
        if "certified_mse" in wl:
            ce = make_engine(wl["certified_mse"])
            ce.SetDT(*eng.GetDT())
            c_total, c_results = timed_registers(ce, max(2, min(args.steps, 5)), 2)
            ce.close()
            cr = c_results[-1]
            c_e2e = timed_e2e(3, 1, mse=wl["certified_mse"])
            out_extra["certified"] = {"mse_threshold": wl["certified_mse"], "exit_path": cr["exit_path"], "time_to_certified_optimum_s": c_total / len(c_results),
                                      "ms_per_step": 1e3 * c_total / len(c_results), "bound_evals": int(cr["bound_evals"]), "bound_evals_executed": int(cr["bound_evals_executed"]),
                                      "value": sum(r["bound_evals"] for r in c_results) / c_total, "unit": "bound-evals/s", "rot_pops": int(cr["rot_pops"]),
                                      "trans_pops": int(cr["trans_pops"]), "sse": cr["sse"], "best_lb": cr["best_lb"], "sse_thresh": cr["sse_thresh"],
                                      "seconds_icp": float(np.mean([r["seconds_icp"] for r in c_results])),
                                      "seconds_bnb_kernels": float(np.mean([r["seconds_bnb_kernels"] for r in c_results])),
                                      "e2e_seconds_per_step": c_e2e["seconds_per_step"], "e2e_value": c_e2e["value"],
                                      "reference_cpu_here": {"register_s": 338.6, "bound_evals": 1696656, "note": "oracle/_ref in the build container (tests/golden/goicp_runs.json)"}}

        if "certified_mse" in wl:
            # the reference's OTHER search strategy (its GPU path, src/fgoicp): quaternion cube + relaxed ICP trigger, include/goicp_b200.h
            fe = make_engine()
            fe.search_mode = 1
            fe.SetDT(*eng.GetDT())
            f_total, f_results = timed_registers(fe, 3, 2)
            fe.close()
            fr = f_results[-1]
            dR = float(2 * np.arcsin(min(1.0, np.linalg.norm(fr["R"].astype(np.float64) - res["R"].astype(np.float64)) / (2 * np.sqrt(2)))))
            out_extra["fgoicp_style_search"] = {"seconds_per_step": f_total / 3, "exit_path": fr["exit_path"], "rot_pops": int(fr["rot_pops"]), "trans_pops": int(fr["trans_pops"]),
                                                "bound_evals": int(fr["bound_evals"]), "icp_calls": int(fr["icp_calls"]), "nn_sse": fr["sse"],
                                                "pose_vs_default_mode": {"dR_rad": dR, "dt": float(np.abs(fr["t"] - res["t"]).max())},
                                                "note": "strategy of icp::FastGoICP (quaternion cube, span cut-offs 0.1 / 0.12, ub < 2 best => ICP) on this engine's kernels; no executable oracle"}

        if args.workload == "bunny_goicp_toml":
            # BASELINE config 4 next to the default line: test/spanner_goicp.toml as written (mse 1e-4, [-1,1]^3 translation cube) -- the
            # throughput-bound certified search (5 305 rotation pops, 154 M bound evaluations; 4.7 h of the reference on one core)
            sp = WORKLOADS["spanner_goicp_toml"]
            sm_, sd_ = clouds_of(sp)
            se = make_engine(wl=sp, clouds=(sm_, sd_))
            se.BuildDT()
            s_total, s_results = timed_registers(se, 2, 1)
            sr = s_results[-1]
            s_kern = sum(r["seconds_bnb_kernels"] for r in s_results)
            s_lookups = sum(r["bound_evals_executed_local"] for r in s_results) * len(sd_)
            rng2 = np.random.default_rng(7)
            npar2 = 148 * 16
            Rs2 = np.stack([np.linalg.qr(rng2.normal(size=(3, 3)))[0] for _ in range(npar2)]).astype(np.float32)
            tc2 = np.concatenate([rng2.uniform(-0.5, 0.25, (npar2, 3)), np.full((npar2, 1), 0.25)], 1).astype(np.float32)
            _, _, ms2 = se.ExpandBounds(Rs2.reshape(npar2, 9), np.full(npar2, -1, np.int32), tc2, repeats=20)
            pat2 = npar2 * 8 * len(sd_) / (ms2 * 1e-3)
            smask = 0
            for r in s_results:
                smask |= int(r["bnb_kernel_variants"])
            se.close()
            out_extra["spanner_goicp_toml"] = {
                "config": "test/spanner_goicp.toml as written: Nd %d, Nm %d, S 300, mse 1e-4, translation cube [-1,1]^3, trim 0" % (len(sd_), len(sm_)),
                "exit_path": sr["exit_path"], "time_to_certified_optimum_s": s_total / len(s_results), "value": sum(r["bound_evals"] for r in s_results) / s_total,
                "unit": "bound-evals/s", "bound_evals": int(sr["bound_evals"]), "bound_evals_executed": int(sr["bound_evals_executed"]), "rot_pops": int(sr["rot_pops"]),
                "trans_pops": int(sr["trans_pops"]), "rounds": int(sr["rounds"]), "sse": sr["sse"], "seconds_bnb_kernels": s_kern / len(s_results),
                "seconds_icp": float(np.mean([r["seconds_icp"] for r in s_results])),
                "roofline": {"kernel": " + ".join(v for k, v in KERNEL_NAMES.items() if smask & k), "lookups_per_s": s_lookups / s_kern,
                             "achieved": s_lookups * 32 / s_kern / 1e9, "peak": pat2 * 32 / 1e9, "unit": "GB/s", "frac": s_lookups / s_kern / pat2,
                             "peak_source": "pattern gather of this cloud in this run (expand_bounds_kernel, %.1f G look-ups/s) x 32 B sector" % (pat2 / 1e9)},
                "reference_cpu_published_here": {"register_s": sp["ref_register_s"], "bound_evals": sp["ref_evals"],
                                                 "note": "oracle/_ref (unmodified reference) in the build container, 1 core (tests/golden/goicp_runs.json)"}}

    # ---- roofline of the dominant kernel (this rank's GPU only)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    grid_bytes = 4 * S ** 3
    l2_resident = grid_bytes <= L2_BYTES
    # Two measured ceilings of a scattered 4-byte gather into a grid of this size (32 B sector per look-up):
    #  * uniform: random addresses over the whole buffer and nothing else (goicp_measure_gather) -- what L2 (or HBM) delivers
    #    to loads with no locality at all;
    #  * pattern: the bound evaluation's own access pattern at full occupancy (expand_bounds_kernel: every CTA evaluates the 8
    #    children of one translation cube for all points; the 8 look-ups of a point are neighbours and share sectors in L1).
    #    It runs at 1.0 look-up / clk / SM, the L1TEX sector rate -- the highest rate any kernel with this pattern has shown.
    # peak = the larger of the two for an L2-resident grid (so no fraction can exceed 1), the HBM copy rate otherwise.
    uniform_lps = eng.MeasureGather(grid_bytes, 5)
    rng = np.random.default_rng(7)
    npar = 148 * 16 if len(data) <= 20000 else 148 * 2
    Rs = np.stack([np.linalg.qr(rng.normal(size=(3, 3)))[0] for _ in range(npar)]).astype(np.float32)
    tc = np.concatenate([rng.uniform(-0.5, 0.25, (npar, 3)), np.full((npar, 1), 0.25)], 1).astype(np.float32)
    _, _, ms = eng.ExpandBounds(Rs.reshape(npar, 9), np.full(npar, -1, np.int32), tc, repeats=20)
    gl = npar * 8 * len(data)
    pattern_lps = gl / (ms * 1e-3)
    if l2_resident:
        best = max(uniform_lps, pattern_lps)
        peak = best * 32 / 1e9
        peak_src = (f"measured in this run: max(pattern gather {pattern_lps / 1e9:.1f} G look-ups/s [expand_bounds_kernel, the L1TEX sector rate], "
                    f"uniform random gather {uniform_lps / 1e9:.1f} G look-ups/s [goicp_measure_gather over {grid_bytes / 1e6:.0f} MB, L2-resident]) x 32 B sector")
    else:
        peak, peak_src = hbm, hbm_src + f"; the {grid_bytes / 1e6:.0f} MB grid does not fit L2"
    mask = 0
    for r in results:
        mask |= int(r["bnb_kernel_variants"])
    kname = " + ".join(v for k, v in KERNEL_NAMES.items() if mask & k) or "inner_bnb_pipelined_kernel"
    lookups_local = executed_local * len(data)
    achieved = lookups_local * 32 / kern_s / 1e9
    traffic, traffic_src = ncu_traffic(args.workload, kname)
    roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes_per_launch": lookups_local * 32 / max(1, rounds),
                "peak_source": peak_src, "memory_level": "L2 / L1TEX (grid resident in L2)" if l2_resident else "HBM",
                "basis": "32 B sector per DT look-up (SURVEY 8d); look-ups = bound evals this GPU executed * Nd; time = CUDA events around the kernel on the engine stream (incl. the result exchange at N>1)",
                "launches": rounds, "avg_launch_ms": 1e3 * kern_s / max(1, rounds), "lookups_per_s": lookups_local / kern_s,
                "frac_of_uniform_random_gather": lookups_local / kern_s / uniform_lps,
                "useful_bytes_frac_of_hbm": lookups_local * 4 / kern_s / 1e9 / hbm, "hbm_peak": hbm,
                "gather": {"kernel": "expand_bounds_kernel", "lookups_per_launch": gl, "ms_per_launch": ms, "lookups_per_s": pattern_lps,
                           "achieved": pattern_lps * 32 / 1e9, "peak": peak, "unit": "GB/s", "frac": pattern_lps * 32 / 1e9 / peak,
                           "lookups_per_clk_per_sm": pattern_lps / (148 * 1e6 * (clocks.get("sm_mhz") or 1965.0))},
                "uniform_random_gather": {"kernel": "gather_peak_kernel", "buffer_mb": grid_bytes / 1e6, "lookups_per_s": uniform_lps, "gbs_of_sectors": uniform_lps * 32 / 1e9}}
    if not args.no_extras:
        other = 4 * (512 ** 3 if S != 512 else 300 ** 3)
        roofline["uniform_random_gather_other_size"] = {"buffer_mb": other / 1e6, "lookups_per_s": eng.MeasureGather(other, 5)}

    out = {"metric": "goicp_bound_evals_per_sec", "value": value, "unit": "bound-evals/s", "n_gpus": world, "steps": args.steps,
           "warmup": warm, "ms_per_step": 1e3 * total_s / args.steps, "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "f32",
           "data": "the reference's own clouds, deterministic subsample (committed fixtures tests/golden/*.f32)" if "synth" not in wl else "synthetic (seeded closed surface + noisy moved subset)",
           "config": config, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
           "time_to_optimum_s": total_s / args.steps, "exit_path": res["exit_path"], "sse": res["sse"],
           "bound_evals_per_step": evals // args.steps, "bound_evals_executed_per_step": executed // args.steps,
           "bound_evals_executed_this_gpu_per_step": executed_local // args.steps,
           "rot_pops": res["rot_pops"], "trans_pops": res["trans_pops"], "rounds_per_step": res["rounds"], "icp_calls": res["icp_calls"],
           "dt_build_s_first_call": dt_build_s, "seconds_icp_per_step": float(np.mean([r["seconds_icp"] for r in results])),
           "seconds_bnb_kernels_per_step": kern_s / args.steps, "seconds_dt_score_per_step": float(np.mean([r["seconds_dt_score"] for r in results])),
           "seconds_strict_resolves_per_step": float(np.mean([r["seconds_strict"] for r in results])), "strict_resolves_per_step": int(res["strict_resolves"]),
           "contender_overflows": int(sum(r["contender_overflows"] for r in results)),
           "seconds_host_frontier_per_step": float(np.mean([r["seconds_host"] for r in results])),
           "host_frontier_share_of_register": float(np.mean([r["seconds_host"] / r["seconds_total"] for r in results])),
           "exchange": (os.environ.get("GOICP_EXCHANGE", "nccl") if world > 1 else None)}
    out.update(out_extra)
    if wl.get("ref_register_s"):
        out["reference_cpu_published_here"] = {"register_s": wl["ref_register_s"], "bound_evals": wl["ref_evals"],
                                               "note": "oracle/_ref in the build container, 1 core (tests/golden/goicp_runs.json)"}
    eng.close()
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            try:
                out["cpu_baseline"] = cpu_reference_sample(wl, args.cpu_seconds, model, data)
            except Exception as e:  # the checker missing must not hide the GPU numbers
                out["cpu_baseline"] = {"error": repr(e)}
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
