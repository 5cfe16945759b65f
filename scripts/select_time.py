#!/usr/bin/env python
"""Block-wide intro_select (goicp_intro_select) timing: n residual-like values, position n - 1 (no trimming) and 0.9 n - 1,
array in shared / global memory.  Wall time per call includes the H2D / D2H of the array."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("cuda-go-icp_b200")
g = pkg.GoICP(1e-3)
rng = np.random.default_rng(0)
for n in (3019, 100000, 1000000):
    a = np.abs(rng.normal(size=n)).astype(np.float32) * 0.05
    for k in (n - 1, int(0.9 * n) - 1):
        for in_global in (False, True):
            if not in_global and n > 15000: continue
            g.IntroSelect(a, k, in_global=in_global)
            t0 = time.perf_counter()
            for _ in range(5): g.IntroSelect(a, k, in_global=in_global)
            print(f"n {n} k {k} {'global' if in_global else 'shared'}: {1e3 * (time.perf_counter() - t0) / 5:.3f} ms per call", flush=True)
g.close()
