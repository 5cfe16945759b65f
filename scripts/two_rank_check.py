#!/usr/bin/env python
"""Registrations sharded over the ranks of a torchrun launch (native NCCL exchange, include/goicp_b200.h: goicp_nccl_init);
every rank's result goes to rank 0, which writes {run: [result of rank 0, rank 1, ...]} as JSON.  Used by
tests/test_gpu_gaps.py::test_two_rank_nccl_registration_vs_golden and runnable by hand:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 scripts/two_rank_check.py out.json
"""
import importlib, json, os, sys
import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
NAMES = ["bunny_s0.1_mse1e-3", "bunny_s0.1_mse7e-4", "bunny_s0.1_mse1e-3_trim0.1", "skull_s0.03_mse1e-3"]


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl")
    pkg = importlib.import_module("cuda-go-icp_b200")
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(pkg.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    nccl_id = bytes(idt.cpu().numpy().tobytes())
    runs = json.load(open(os.path.join(GOLDEN, "goicp_runs.json")))
    out = {}
    for name in NAMES:
        gold = runs[name]
        ld = lambda n: np.fromfile(os.path.join(GOLDEN, n), np.float32).reshape(-1, 3)
        g = pkg.GoICP(gold["mse"], device=local)
        g.pModel, g.pData = ld(gold["model"]), ld(gold["data"])
        g.trimFraction = gold["trim"]
        if "trans_cube" in gold:
            g.initNodeTrans = gold["trans_cube"]
        g.init_nccl(nccl_id, rank, world)
        g.BuildDT()
        g.Register()
        r = g.result
        g.close()
        mine = {"R": [float(x) for x in r["R"].reshape(-1)], "t": [float(x) for x in r["t"]], "sse": r["sse"], "exit_path": r["exit_path"],
                "rot_pops": int(r["rot_pops"]), "trans_pops": int(r["trans_pops"]), "bound_evals": int(r["bound_evals"]),
                "bound_evals_executed": int(r["bound_evals_executed"]), "bound_evals_executed_local": int(r["bound_evals_executed_local"]),
                "rounds": int(r["rounds"]), "seconds_total": r["seconds_total"]}
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        out[name] = gathered
    # sharded ICP (tolerance numerics, large cloud): the ranks split the nearest-neighbour queries and all-gather 16 moments per
    # iteration; against the same ICP on one GPU
    from bench import synth
    model, data, R_gt, t_gt = synth(60000, 24000)
    a = 0.06
    R0 = (np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]]) @ R_gt).astype(np.float32)
    t0 = (t_gt + np.array([0.02, -0.01, 0.015])).astype(np.float32)
    icp = {}
    for mode in ("sharded", "single"):
        g = pkg.GoICP(1e-4, device=local)
        g.pModel, g.pData = model, data
        g.numerics = 2
        if mode == "sharded":
            g.init_nccl(nccl_id, rank, world)
        err, R, t, iters = g.ICP(R0, t0, 200, 1e-7)
        g.close()
        icp[mode] = {"err": float(err), "R": [float(x) for x in R.reshape(-1)], "t": [float(x) for x in t], "iterations": int(iters)}
    gathered = [None] * world
    dist.all_gather_object(gathered, icp)
    out["_icp_shard"] = gathered
    out["_icp_shard_gt"] = {"R": [float(x) for x in R_gt.reshape(-1)], "t": [float(x) for x in t_gt]}
    if rank == 0:
        json.dump(out, open(sys.argv[1], "w"), indent=1)
        for name, per in out.items():
            if name.startswith("_"):
                continue
            print(name, [(p["rot_pops"], p["trans_pops"], round(p["sse"], 6), p["bound_evals_executed_local"]) for p in per])
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
