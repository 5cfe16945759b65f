"""ctypes bindings for the parity oracles.  TEST INFRASTRUCTURE ONLY.

Two checkers live under oracle/:

* ``Restated``  -- oracle/libgoicp_oracle.so, the plain-C restatement in goicp_oracle.c
  (always buildable: ``make -C oracle libgoicp_oracle.so``).
* ``Reference`` -- oracle/_ref/libref_goicp.so, the UNMODIFIED reference CPU Go-ICP compiled
  from /root/reference by oracle/Makefile (git-ignored; present only where it was built).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i16p = np.ctypeslib.ndpointer(np.int16, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")


def build(ref: bool = True) -> None:
    """Compile the restatement (and the reference when /root/reference exists)."""
    subprocess.run(["make", "-C", _HERE, "libgoicp_oracle.so"], check=True, capture_output=True)
    if ref:
        subprocess.run(["make", "-C", _HERE, "ref"], check=True, capture_output=True)


def _xyz(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1, 3)
    return a


def fnv1a64(buf: bytes) -> int:
    h = 1469598103934665603
    for b in buf:  # small inputs only; big ones go through Restated.fnv
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


class Restated:
    """oracle/goicp_oracle.c"""

    def __init__(self, path: str | None = None):
        path = path or os.path.join(_HERE, "libgoicp_oracle.so")
        if not os.path.exists(path):
            build(ref=False)
        L = self.L = C.CDLL(path)
        L.go_dt_build.restype = C.c_void_p
        L.go_dt_build.argtypes = [_f32p, C.c_int, C.c_int, C.c_double, C.c_int]
        L.go_dt_wrap.restype = C.c_void_p
        L.go_dt_wrap.argtypes = [_f32p, C.c_int, _f64p]
        L.go_dt_meta.argtypes = [C.c_void_p, _f64p]
        L.go_dt_grid.argtypes = [C.c_void_p, _f32p]
        L.go_dt_vectors.argtypes = [C.c_void_p, _i16p]
        L.go_dt_free.argtypes = [C.c_void_p]
        L.go_dt_frame.argtypes = [_f32p, C.c_int, C.c_int, C.c_double, _f64p]
        L.go_dt_distance_batch.argtypes = [C.c_void_p, _f32p, C.c_int, _f32p]
        L.go_dt_index.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, _i32p]
        L.go_intro_select.argtypes = [_f32p, C.c_size_t, C.c_size_t, C.c_size_t]
        L.go_kd_build.restype = C.c_void_p
        L.go_kd_build.argtypes = [_f32p, C.c_int]
        L.go_kd_free.argtypes = [C.c_void_p]
        L.go_kd_num_nodes.argtypes = [C.c_void_p]
        L.go_kd_export.argtypes = [C.c_void_p, _i32p, _i32p, _f32p]
        L.go_kd_nn.argtypes = [C.c_void_p, _f32p, C.c_int, _i32p, _f32p]
        L.go_svd3.argtypes = [_f32p, _f32p, _f32p, _f32p]
        L.go_icp_run_ex.restype = C.c_float
        L.go_icp_run_ex.argtypes = [C.c_void_p, _f32p, C.c_int, _f32p, _f32p, C.c_long, C.c_float, C.c_float, C.c_int,
                                    C.POINTER(C.c_int), C.c_int, C.c_void_p]
        L.go_create.restype = C.c_void_p
        L.go_create.argtypes = [_f32p, C.c_int, _f32p, C.c_int, C.c_float, C.c_float, C.c_int, C.c_double, C.c_void_p]
        L.go_build_dt.argtypes = [C.c_void_p]
        L.go_set_dt.argtypes = [C.c_void_p, C.c_void_p]
        L.go_get_dt.restype = C.c_void_p
        L.go_get_dt.argtypes = [C.c_void_p]
        L.go_initialize.argtypes = [C.c_void_p]
        L.go_max_rot_dis.argtypes = [C.c_void_p, C.c_int, _f32p]
        L.go_sse_thresh.restype = C.c_float
        L.go_sse_thresh.argtypes = [C.c_void_p]
        L.go_inner.restype = C.c_float
        L.go_inner.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_float, _f32p, _i64p]
        L.go_dt_score.restype = C.c_float
        L.go_dt_score.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.go_icp.restype = C.c_float
        L.go_icp.argtypes = [C.c_void_p, _f32p, _f32p]
        L.go_cube_rotation.restype = C.c_int
        L.go_cube_rotation.argtypes = [C.c_float, C.c_float, C.c_float, C.c_float, _f32p]
        L.go_register.restype = C.c_float
        L.go_register.argtypes = [C.c_void_p, _f64p, _i64p]
        L.go_improve_log.argtypes = [C.c_void_p, _f32p]
        L.go_set_budget.argtypes = [C.c_void_p, C.c_double]
        L.go_set_verbose.argtypes = [C.c_void_p, C.c_int]
        L.go_set_do_trim.argtypes = [C.c_void_p, C.c_int]
        L.go_set_rot_cube.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float]
        L.go_free.argtypes = [C.c_void_p]
        L.go_fnv1a64.restype = C.c_uint64
        L.go_fnv1a64.argtypes = [C.c_void_p, C.c_size_t]

    # -- DT -------------------------------------------------------------------------------
    def dt_build(self, model, S, expand=2.0, keep_vectors=False):
        m = _xyz(model)
        return self.L.go_dt_build(m, len(m), S, expand, int(keep_vectors))

    def dt_wrap(self, grid, S, meta):
        return self.L.go_dt_wrap(np.ascontiguousarray(grid, np.float32).ravel(), S, np.ascontiguousarray(meta, np.float64))

    def dt_meta(self, dt):
        m = np.zeros(4, np.float64)
        self.L.go_dt_meta(dt, m)
        return m

    def dt_frame(self, model, S, expand=2.0):
        m = _xyz(model)
        out = np.zeros(4, np.float64)
        self.L.go_dt_frame(m, len(m), S, expand, out)
        return out

    def dt_grid(self, dt, S):
        g = np.zeros(S * S * S, np.float32)
        self.L.go_dt_grid(dt, g)
        return g.reshape(S, S, S)

    def dt_vectors(self, dt, S):
        g = np.zeros(S * S * S * 3, np.int16)
        self.L.go_dt_vectors(dt, g)
        return g.reshape(S, S, S, 3)

    def dt_distance(self, dt, q):
        q = _xyz(q)
        out = np.zeros(len(q), np.float32)
        self.L.go_dt_distance_batch(dt, q, len(q), out)
        return out

    def dt_index(self, dt, q):
        q = _xyz(q)
        out = np.zeros((len(q), 3), np.int32)
        tmp = np.zeros(3, np.int32)
        for i, p in enumerate(q):
            self.L.go_dt_index(dt, float(p[0]), float(p[1]), float(p[2]), tmp)
            out[i] = tmp
        return out

    def dt_free(self, dt):
        self.L.go_dt_free(dt)

    def fnv(self, arr: np.ndarray) -> int:
        a = np.ascontiguousarray(arr)
        return int(self.L.go_fnv1a64(a.ctypes.data_as(C.c_void_p), a.nbytes))

    def intro_select(self, a, k):
        a = np.ascontiguousarray(a, np.float32).copy()
        self.L.go_intro_select(a, 0, len(a) - 1, k)
        return a

    # -- kd-tree / ICP ----------------------------------------------------------------------
    def kd_build(self, model):
        m = _xyz(model)
        return self.L.go_kd_build(m, len(m))

    def kd_export(self, kd, nm):
        nn = self.L.go_kd_num_nodes(kd)
        nodes = np.zeros(nn * 7, np.int32)
        vind = np.zeros(nm, np.int32)
        bbox = np.zeros(6, np.float32)
        self.L.go_kd_export(kd, nodes, vind, bbox)
        return nodes.reshape(nn, 7), vind, bbox

    def kd_nn(self, kd, q):
        q = _xyz(q)
        idx = np.zeros(len(q), np.int32)
        d2 = np.zeros(len(q), np.float32)
        self.L.go_kd_nn(kd, q, len(q), idx, d2)
        return idx, d2

    def svd3(self, H):
        H = np.ascontiguousarray(H, np.float32).reshape(9)
        U = np.zeros(9, np.float32); W = np.zeros(3, np.float32); V = np.zeros(9, np.float32)
        self.L.go_svd3(H, U, W, V)
        return U.reshape(3, 3), W, V.reshape(3, 3)

    def icp_run(self, kd, data, R, t, max_iter=10000, err_diff=1e-6, trim=0.0, do_trim=True, trace_iter=-1):
        d = _xyz(data)
        R = np.ascontiguousarray(R, np.float32).reshape(9).copy()
        t = np.ascontiguousarray(t, np.float32).reshape(3).copy()
        iters = C.c_int(0)
        trace = np.zeros(len(d), np.int32) if trace_iter >= 0 else None
        e = self.L.go_icp_run_ex(kd, d, len(d), R, t, max_iter, err_diff, trim, int(do_trim), C.byref(iters), trace_iter,
                                 trace.ctypes.data_as(C.c_void_p) if trace is not None else None)
        return float(e), R.reshape(3, 3), t, iters.value, trace

    # -- Go-ICP -----------------------------------------------------------------------------
    def create(self, model, data, mse=1e-3, trim=0.0, S=300, expand=2.0, trans_cube=None):
        m, d = _xyz(model), _xyz(data)
        tc = None
        if trans_cube is not None:
            tc = np.ascontiguousarray(trans_cube, np.float32)
            self._keep = tc
            tc = tc.ctypes.data_as(C.c_void_p)
        return self.L.go_create(m, len(m), d, len(d), mse, trim, S, expand, tc)

    def register(self, g):
        out = np.zeros(16, np.float64)
        cnt = np.zeros(8, np.int64)
        self.L.go_register(g, out, cnt)
        log = np.zeros(64, np.float32)
        self.L.go_improve_log(g, log)
        return {
            "R": out[:9].reshape(3, 3).copy(), "t": out[9:12].copy(), "sse": float(np.float32(out[12])),
            "sse_thresh": float(np.float32(out[13])), "register_s": float(out[14]), "exit_lb": float(np.float32(out[15])),
            "rot_pops": int(cnt[0]), "trans_pops": int(cnt[1]), "bound_evals": int(cnt[2]), "icp_calls": int(cnt[3]),
            "inner_calls": int(cnt[4]),
            "exit_path": {0: "none", 1: "certified", 2: "early_sse_below_thresh", 3: "queue_empty", 4: "budget"}[int(cnt[5])],
            "improvements": [float(x) for x in log[: int(cnt[6])]],
        }

    def inner(self, g, R, level, opt_error):
        R = np.ascontiguousarray(R, np.float32).reshape(9)
        out = np.zeros(5, np.float32)
        cnt = np.zeros(2, np.int64)
        v = self.L.go_inner(g, R, level, opt_error, out, cnt)
        return {"value": float(np.float32(v)), "node": out[1:5].copy(), "pops": int(cnt[0]), "evals": int(cnt[1])}

    def inner_reuse(self, g):
        """(reuse_gt, reuse_poplb, SSEThresh) of the last inner() call: the range of smaller optErrors it would have run identically for"""
        out = np.zeros(3, np.float32)
        self.L.go_inner_reuse.restype = None
        self.L.go_inner_reuse.argtypes = [C.c_void_p, _f32p]
        self.L.go_inner_reuse(g, out)
        return np.float32(out[0]), np.float32(out[1]), np.float32(out[2])

    def cube_rotation(self, a, b, c, w):
        R = np.zeros(9, np.float32)
        ok = self.L.go_cube_rotation(a, b, c, w, R)
        return bool(ok), R.reshape(3, 3)

    def max_rot_dis(self, g, level, nd):
        out = np.zeros(nd, np.float32)
        self.L.go_max_rot_dis(g, level, out)
        return out


class Reference:
    """oracle/_ref/libref_goicp.so -- the unmodified reference, when it has been built."""

    PATH = os.path.join(_HERE, "_ref", "libref_goicp.so")

    @classmethod
    def available(cls) -> bool:
        return os.path.exists(cls.PATH)

    def __init__(self):
        if not self.available():
            raise FileNotFoundError(self.PATH + " (run `make -C oracle ref` where /root/reference exists)")
        L = self.L = C.CDLL(self.PATH)
        L.ref_subsample.restype = C.c_size_t
        L.ref_subsample.argtypes = [_f32p, C.c_size_t, C.c_float, C.c_float, C.c_uint, _f32p]
        L.ref_dt_build.restype = C.c_void_p
        L.ref_dt_build.argtypes = [_f32p, C.c_int, C.c_int, C.c_double]
        L.ref_dt_meta.argtypes = [C.c_void_p, _f64p]
        L.ref_dt_grid.argtypes = [C.c_void_p, _f32p]
        L.ref_dt_vectors.argtypes = [C.c_void_p, _i16p]
        L.ref_dt_distance.argtypes = [C.c_void_p, _f32p, C.c_int, _f32p]
        L.ref_dt_free.argtypes = [C.c_void_p]
        L.ref_icp_build.restype = C.c_void_p
        L.ref_icp_build.argtypes = [_f32p, C.c_int]
        L.ref_icp_nn.argtypes = [C.c_void_p, _f32p, C.c_int, _i32p, _f32p]
        L.ref_icp_run.restype = C.c_float
        L.ref_icp_run.argtypes = [C.c_void_p, _f32p, C.c_int, _f32p, _f32p, C.c_int, C.c_float, C.c_float, C.c_int]
        L.ref_svd3.argtypes = [_f32p, _f32p, _f32p, _f32p]
        L.ref_goicp_create.restype = C.c_void_p
        L.ref_goicp_create.argtypes = [_f32p, C.c_int, _f32p, C.c_int, C.c_float, C.c_float, C.c_int, C.c_double, C.c_void_p]
        if hasattr(L, "ref_goicp_set_do_trim"):
            L.ref_goicp_set_do_trim.argtypes = [C.c_void_p, C.c_int]
        L.ref_goicp_build_dt.restype = C.c_double
        L.ref_goicp_build_dt.argtypes = [C.c_void_p]
        L.ref_goicp_dt.restype = C.c_void_p
        L.ref_goicp_dt.argtypes = [C.c_void_p]
        L.ref_goicp_register.restype = C.c_float
        L.ref_goicp_register.argtypes = [C.c_void_p, _f64p, _i64p]
        L.ref_goicp_initialize.argtypes = [C.c_void_p]
        L.ref_goicp_inner.restype = C.c_float
        L.ref_goicp_inner.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_float, _f32p, _i64p]
        L.ref_goicp_maxrotdis.argtypes = [C.c_void_p, C.c_int, _f32p]
        L.ref_goicp_icp.restype = C.c_float
        L.ref_goicp_icp.argtypes = [C.c_void_p, _f32p, _f32p]

    def subsample(self, xyz, sub, resize, seed):
        a = _xyz(xyz)
        out = np.zeros_like(a)
        k = self.L.ref_subsample(a, len(a), sub, resize, seed, out)
        return out[:k].copy()

    def dt_build(self, model, S, expand=2.0):
        m = _xyz(model)
        return self.L.ref_dt_build(m, len(m), S, expand)

    def dt_meta(self, dt):
        m = np.zeros(4, np.float64)
        self.L.ref_dt_meta(dt, m)
        return m

    def dt_grid(self, dt, S):
        g = np.zeros(S * S * S, np.float32)
        self.L.ref_dt_grid(dt, g)
        return g.reshape(S, S, S)

    def dt_vectors(self, dt, S):
        g = np.zeros(S * S * S * 3, np.int16)
        self.L.ref_dt_vectors(dt, g)
        return g.reshape(S, S, S, 3)

    def dt_distance(self, dt, q):
        q = _xyz(q)
        out = np.zeros(len(q), np.float32)
        self.L.ref_dt_distance(dt, q, len(q), out)
        return out

    def dt_free(self, dt):
        self.L.ref_dt_free(dt)

    def icp_build(self, model):
        m = _xyz(model)
        return self.L.ref_icp_build(m, len(m))

    def icp_nn(self, icp, q):
        q = _xyz(q)
        idx = np.zeros(len(q), np.int32)
        d2 = np.zeros(len(q), np.float32)
        self.L.ref_icp_nn(icp, q, len(q), idx, d2)
        return idx, d2

    def icp_run(self, icp, data, R, t, max_iter=10000, err_diff=1e-6, trim=0.0, do_trim=True):
        d = _xyz(data)
        R = np.ascontiguousarray(R, np.float32).reshape(9).copy()
        t = np.ascontiguousarray(t, np.float32).reshape(3).copy()
        e = self.L.ref_icp_run(icp, d, len(d), R, t, max_iter, err_diff, trim, int(do_trim))
        return float(e), R.reshape(3, 3), t

    def svd3(self, H):
        H = np.ascontiguousarray(H, np.float32).reshape(9)
        U = np.zeros(9, np.float32); W = np.zeros(3, np.float32); V = np.zeros(9, np.float32)
        self.L.ref_svd3(H, U, W, V)
        return U.reshape(3, 3), W, V.reshape(3, 3)

    def create(self, model, data, mse=1e-3, trim=0.0, S=300, expand=2.0, trans_cube=None):
        m, d = _xyz(model), _xyz(data)
        tc = None
        if trans_cube is not None:
            tc = np.ascontiguousarray(trans_cube, np.float32)
            self._keep = tc
            tc = tc.ctypes.data_as(C.c_void_p)
        return self.L.ref_goicp_create(m, len(m), d, len(d), mse, trim, S, expand, tc)

    def build_dt(self, g):
        return self.L.ref_goicp_build_dt(g)

    def goicp_dt(self, g):
        return self.L.ref_goicp_dt(g)

    def initialize(self, g):
        self.L.ref_goicp_initialize(g)

    def register(self, g):
        out = np.zeros(15, np.float64)
        cnt = np.zeros(5, np.int64)
        self.L.ref_goicp_register(g, out, cnt)
        return {"R": out[:9].reshape(3, 3).copy(), "t": out[9:12].copy(), "sse": float(np.float32(out[12])),
                "sse_thresh": float(np.float32(out[13])), "register_s": float(out[14]),
                "rot_pops": int(cnt[0]), "trans_pops": int(cnt[1]), "select_calls": int(cnt[4])}

    def inner(self, g, R, level, opt_error):
        R = np.ascontiguousarray(R, np.float32).reshape(9)
        out = np.zeros(5, np.float32)
        cnt = np.zeros(2, np.int64)
        v = self.L.ref_goicp_inner(g, R, level, opt_error, out, cnt)
        return {"value": float(np.float32(v)), "node": out[1:5].copy(), "pops": int(cnt[0]), "evals": int(cnt[1])}

    def max_rot_dis(self, g, level, nd):
        out = np.zeros(nd, np.float32)
        self.L.ref_goicp_maxrotdis(g, level, out)
        return out

    def goicp_icp(self, g, R, t):
        R = np.ascontiguousarray(R, np.float32).reshape(9).copy()
        t = np.ascontiguousarray(t, np.float32).reshape(3).copy()
        e = self.L.ref_goicp_icp(g, R, t)
        return float(np.float32(e)), R.reshape(3, 3), t
