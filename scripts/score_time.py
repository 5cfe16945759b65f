#!/usr/bin/env python
"""Reference-order DT score (goicp_dt_score) on the sweep_100k inputs (Nd 1e5, Nm 1e6, S 512): wall time per call of the initial
(raw data) score and of a posed score.  Build with -DGOICP_SCORE_TRACE for the in-kernel select / sum cycle counts."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module("cuda-go-icp_b200")
nd = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
model, data, R, t = bench.synth(10 * nd, nd)
g = pkg.GoICP(1e-4); g.pModel, g.pData = model, data; g.dt.SIZE = 512 if nd >= 50000 else 300
g.BuildDT()
for name, args in (("raw", (None, None)), ("posed", (R, t)), ("identity pose", (np.eye(3, dtype=np.float32), np.zeros(3, np.float32)))):
    g.DTScore(*args)
    t0 = time.perf_counter()
    for _ in range(3): s = g.DTScore(*args)
    print(f"{name}: {1e3 * (time.perf_counter() - t0) / 3:.3f} ms per score, sse {s}", flush=True)
g.close()
