"""Executable specification of DESIGN.md section 10's "exact sequential float sums in parallel" (analysed, kernel not built).

The reference adds Nd floats one after the other (ICP3D::Run, jly_icp3d.hpp:241-254; H = ~q_t*q_m, :268; GoICP's SSE sums,
jly_goicp.cpp:293-315), and the strict kernels reproduce those sums bit for bit with a chain of dependent FADDs -- 4 cycles per
term, the floor of two of the ICP kernel's phases.  s <- fl(s + a_i) is not associative, but while the running sum stays inside
one binade [2^e, 2^(e+1)) every step is integer arithmetic on k = s / ulp:

    a_i / ulp = m_i + f_i  (m_i integer, 0 <= f_i < 1),   k <- k + m_i + r,   r = [f_i > 1/2]  or, on a tie f_i = 1/2,
                                                                             r = (k + m_i) odd   (round half to even)

i.e. k <- k + c_i[k mod 2] with a pair of integers c_i per term.  Maps of that form are closed under composition,
(G o F).c[p] = F.c[p] + G.c[(p + F.c[p]) mod 2], which is associative: a prefix scan yields every running sum exactly as long
as it stays strictly inside the binade, a range check finds the first step that leaves it, that one step is a real float add,
and the scan restarts in the new binade.  `sequential_sum_by_scans` is that algorithm with the scan spelled out as
Hillis-Steele doubling steps (what a thread block would run); tests/test_exact_parallel_sums.py checks it bit for bit against
the plain float32 loop on adversarial chains.  The number of restarts is what decides whether a kernel pays: measured with this
file, 32 / 47 / 54 per 3019 terms on the centroid chains of the bunny model (a third off those chains at ~150 cycles per
block-wide scan), about one per window plus one per binade on the SSE sum of a 1e5-point cloud (4x) -- and one every four
terms on a zero-mean chain (the off-diagonal entries of H), where the scheme degenerates to the sequential chain it
replaces, never below it."""
import numpy as np

_LO, _HI = 1 << 23, 1 << 24


def _compose(f0, f1, g0, g1):
    """(G o F) for maps k -> k + c[k & 1]; F is applied first"""
    return f0 + np.where((f0 & 1) == 1, g1, g0), f1 + np.where(((1 + f1) & 1) == 1, g1, g0)


def _prefix_maps(c0, c1):
    """inclusive scan of the maps by doubling: after the step with distance d, entry j covers the 2d terms ending at j"""
    c0, c1 = c0.copy(), c1.copy()
    d = 1
    while d < len(c0):
        n0, n1 = _compose(c0[:-d], c1[:-d], c0[d:], c1[d:])
        c0[d:], c1[d:] = n0, n1
        d *= 2
    return c0, c1


def sequential_sum_by_scans(s0, a, window=1024):
    """fl(...fl(fl(s0 + a[0]) + a[1])... + a[n-1]) in float32, bit for bit; returns (sum, scans run, real adds made)"""
    a = np.asarray(a, np.float32)
    s = np.float32(s0)
    i, n, scans, adds = 0, len(a), 0, 0
    while i < n:
        mag = abs(float(s))
        if not np.isfinite(s) or mag < 2.0 ** -100:                 # zero / tiny / non-finite: nothing to scan from, one real add
            s = np.float32(s + a[i]); i += 1; adds += 1
            continue
        sign = 1.0 if s > 0 else -1.0
        e = int(np.floor(np.log2(mag)))
        if 2.0 ** e > mag: e -= 1                                    # (log2 rounding)
        if 2.0 ** (e + 1) <= mag: e += 1
        u = 2.0 ** (e - 23)
        k0 = int(mag / u)                                            # exact: mag is a float32 in [2^e, 2^(e+1))
        assert _LO <= k0 < _HI and k0 * u == mag
        w = a[i:i + window].astype(np.float64) * sign
        t = np.clip(w / u, -2.0 ** 40, 2.0 ** 40)                    # exact scaling by a power of two; a huge term leaves the binade anyway
        m = np.floor(t)
        f = t - m                                                    # exact: |t| < 2^40 has at most 24 significant bits
        mi = m.astype(np.int64)
        up = (f > 0.5).astype(np.int64)
        tie = f == 0.5
        c0 = mi + np.where(tie, mi & 1, up)                          # k even: k + m is odd iff m is odd
        c1 = mi + np.where(tie, (mi + 1) & 1, up)                    # k odd
        p0, p1 = _prefix_maps(c0, c1)
        scans += 1
        k = k0 + (p1 if k0 & 1 else p0)
        ok = (k > _LO) & (k < _HI)                                   # strictly inside: at the edges the real add decides
        v = len(w) if ok.all() else int(np.argmin(ok))
        if v > 0:
            s = np.float32(sign * float(k[v - 1]) * u)               # exact
            i += v
        if v < len(w):
            s = np.float32(s + a[i]); i += 1; adds += 1
    return s, scans, adds


def sequential_sum(s0, a):
    s = np.float32(s0)
    for x in np.asarray(a, np.float32):
        s = np.float32(s + x)
    return s
