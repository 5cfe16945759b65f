"""scripts/proto/exact_parallel_sums.py (the scan formulation of the reference's sequential float sums)
against the plain float32 loop, bit for bit, on chains built to break it."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "scripts", "proto"))
from exact_parallel_sums import sequential_sum, sequential_sum_by_scans  # noqa: E402


def _chains():
    rng = np.random.default_rng(11)
    yield "normal", 0.0, rng.normal(size=3019)
    yield "drift", 0.0, rng.uniform(0.0, 1.0, 5000)
    yield "negative drift from a positive start", 700.0, -rng.uniform(0.0, 1.0, 4000)
    yield "cancelling pairs", 1.0, np.repeat(rng.normal(size=1500), 2) * np.tile([1, -1], 1500)
    yield "wide dynamic range", 3.0, rng.normal(size=3000) * 10.0 ** rng.integers(-12, 6, 3000)
    yield "zeros and repeats", 0.0, np.concatenate([np.zeros(100), np.full(500, 0.1), np.zeros(7), np.full(300, -0.3)])
    yield "squared residuals", 0.0, rng.uniform(0, 0.02, 3019) ** 2
    s = np.float32(1024.0)
    u = np.spacing(s)
    yield "exact half-ulp ties", s, (rng.integers(-3, 4, 2000) + 0.5) * u          # every term a tie: round-half-even on each step
    yield "ties and non-ties mixed", s, rng.choice([0.5, 1.5, -0.5, 0.25, 0.75, 2.0, -1.5], 3000) * u
    yield "binade edge", np.float32(2.0 ** 24 - 3), rng.choice([1.0, 2.0, -1.0, 3.0, -4.0], 1000)
    yield "denormal neighbourhood", 1e-38, rng.normal(size=500) * 1e-39
    yield "huge then small", 1.0, np.concatenate([[3e38, -3e38], rng.normal(size=200)])
    base = rng.normal(size=(3019, 3)) * 0.2
    yield "centroid chain from a stale mean", 0.013, base[:, 0]                    # jly_icp3d.hpp:205-206,241-247: mu is never reset
    q, m = base - base.mean(0), base[::-1] - base.mean(0)
    yield "H entry", 0.0, (q[:, 0].astype(np.float32) * m[:, 1].astype(np.float32))


@pytest.mark.parametrize("window", [1024, 64, 7])
def test_scan_formulation_equals_the_sequential_float_sum(window):
    for name, s0, a in _chains():
        a = np.asarray(a, np.float32)
        want = sequential_sum(s0, a)
        got, scans, adds = sequential_sum_by_scans(s0, a, window)
        assert got.tobytes() == want.tobytes(), (name, window, got, want)
        assert adds <= len(a) and scans <= adds + len(a) // window + 2   # one scan per window plus one per step that leaves the binade


def test_restarts_decide_whether_it_pays():
    """what decides whether a kernel pays: scans per chain (each costs a block-wide scan, ~150 cycles, against 4 cycles per term).
    A sum that drifts away from zero changes binade a few dozen times; one that hovers around zero changes it every few terms
    and the scheme degenerates to the sequential chain -- never below it."""
    G = os.path.join(ROOT, "tests", "golden")
    m = np.fromfile(os.path.join(G, "bunny_model_s0.1_seed1234.f32"), np.float32).reshape(-1, 3)
    rng = np.random.default_rng(3)
    for c in range(3):                                              # centroid chains of the bunny model (mean -0.11, -0.17, 0.12)
        a = m[rng.permutation(len(m))[:3019], c]
        got, scans, adds = sequential_sum_by_scans(m[:, c].mean(), a, 1024)
        assert got.tobytes() == sequential_sum(m[:, c].mean(), a).tobytes()
        assert scans < 100, (c, scans)                              # measured 32 / 47 / 54
    walk = (rng.normal(size=3019) * 0.2).astype(np.float32)         # zero-mean: measured 757 scans for 3019 terms
    got, scans, adds = sequential_sum_by_scans(0.013, walk, 1024)
    assert got.tobytes() == sequential_sum(0.013, walk).tobytes() and scans > 300 and adds <= len(walk)
    drift = (rng.uniform(0, 0.02, 100000) ** 2).astype(np.float32)  # the SSE sum of a 1e5-point cloud
    got, scans, adds = sequential_sum_by_scans(0.0, drift, 1024)
    assert got.tobytes() == sequential_sum(0.0, drift).tobytes()
    assert scans < 100000 // 1024 + 80                              # about one scan per window plus one per binade
