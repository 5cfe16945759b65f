"""world_size=2 test of the multi-GPU plumbing on CPU (gloo): the round-robin task sharding and the
all-gather exchange that merges every rank's inner-BnB results, and the hand-round of a contender list from each
rank in turn."""
import os
import subprocess
import sys
import textwrap

from conftest import ROOT

WORKER = textwrap.dedent('''
    import ctypes as C, importlib, os, sys
    import numpy as np
    import torch, torch.distributed as dist
    sys.path.insert(0, %r)
    pkg = importlib.import_module("cuda-go-icp_b200")
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    def allgather(send):
        s = torch.from_numpy(send.copy())
        out = [torch.empty_like(s) for _ in range(world)]
        dist.all_gather(out, s)
        return torch.cat(out).numpy()
    g = pkg.GoICP.__new__(pkg.GoICP)          # only the callback marshalling is needed (no GPU handle)
    g.rank, g.world_size = rank, world
    pkg.GoICP.set_exchange(g, allgather, rank, world)
    bad_total = 0
    for n in (1, 2, 7, 288, 289):
        bad = C.c_int(-1)
        rc = pkg.lib().goicp_selftest_shard(rank, world, n, g._exchange, None, C.byref(bad))
        assert rc == 0, rc
        bad_total += bad.value
    # the best-upper-bound exchange the search relies on: every rank must end with the global min
    ub = torch.tensor([3.5 - rank], dtype=torch.float32)
    dist.all_reduce(ub, op=dist.ReduceOp.MIN)
    assert ub.item() == 3.5 - (world - 1)
    sys.stdout.write("RANK {} mismatches {}\\n".format(rank, bad_total))   # one write: ranks must not interleave
    sys.stdout.flush()
    dist.destroy_process_group()
''')


def test_two_rank_gloo_exchange(tmp_path, pkg):
    pkg.build()
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "RANK 0 mismatches 0" in r.stdout and "RANK 1 mismatches 0" in r.stdout
