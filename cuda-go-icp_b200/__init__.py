"""B200-native Go-ICP engine: thin ctypes driver over the C ABI (include/goicp_b200.h).

Python is test/bench plumbing only.  The product is ``libgoicp_b200.so`` (hand-written sm_100a
CUDA + C++ host, built in-tree by ``cuda-go-icp_b200/Makefile``); the C++ mirror of the
reference's ``class GoICP`` lives in ``host/goicp_b200.hpp``.  Nothing here falls back to the
CPU: if the library or a GPU is missing the calls raise.

The ``GoICP`` class below keeps the reference's member names (``pModel``/``Nm``/``pData``/``Nd``,
``MSEThresh``, ``trimFraction``, ``dt.SIZE``, ``BuildDT()``, ``Register()``, ``optR``/``optT``/
``optError`` -- jly_goicp.h:82-141) so the parity tests read like the reference's usage in
src/main.cpp:47-59,154-159.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GOICP_LIB_PATH") or os.path.join(_HERE, "libgoicp_b200.so")    # override: instrumented builds (profiling only)

EXIT_PATHS = {0: "none", 1: "certified", 2: "early_sse_below_thresh", 3: "queue_empty", 4: "cancelled"}


class GoicpError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"goicp status {code}: {msg}")
        self.code = code


class Params(C.Structure):
    _fields_ = [("mse_threshold", C.c_float), ("trim_fraction", C.c_float), ("do_trim", C.c_int),
                ("dt_size", C.c_int), ("dt_expand", C.c_double),
                ("rot_cube", C.c_float * 4), ("trans_cube", C.c_float * 4),
                ("icp_max_iter", C.c_int), ("device", C.c_int), ("spec_cubes", C.c_int), ("cluster_size", C.c_int), ("dt_mode", C.c_int),
                ("rank", C.c_int), ("world_size", C.c_int), ("numerics", C.c_int), ("search_mode", C.c_int)]


class Result(C.Structure):
    _fields_ = [("R", C.c_float * 9), ("t", C.c_float * 3), ("sse", C.c_float), ("sse_thresh", C.c_float),
                ("best_lb", C.c_float), ("exit_path", C.c_int),
                ("rot_pops", C.c_int64), ("trans_pops", C.c_int64), ("bound_evals", C.c_int64),
                ("bound_evals_executed", C.c_int64), ("icp_calls", C.c_int64), ("rounds", C.c_int64), ("kernel_launches", C.c_int64),
                ("seconds_total", C.c_double), ("seconds_bnb_kernels", C.c_double), ("seconds_icp", C.c_double),
                ("bound_evals_executed_local", C.c_int64), ("strict_resolves", C.c_int64), ("contender_overflows", C.c_int64),
                ("seconds_dt_score", C.c_double), ("seconds_strict", C.c_double), ("seconds_setup", C.c_double), ("seconds_host", C.c_double),
                ("bnb_kernel_variants", C.c_int64)]

    def as_dict(self):
        return {"R": np.array(self.R, np.float32).reshape(3, 3), "t": np.array(self.t, np.float32),
                "sse": float(self.sse), "sse_thresh": float(self.sse_thresh), "best_lb": float(self.best_lb),
                "exit_path": EXIT_PATHS[self.exit_path], "rot_pops": self.rot_pops, "trans_pops": self.trans_pops,
                "bound_evals": self.bound_evals, "bound_evals_executed": self.bound_evals_executed,
                "icp_calls": self.icp_calls, "rounds": self.rounds, "kernel_launches": self.kernel_launches, "seconds_total": self.seconds_total,
                "seconds_bnb_kernels": self.seconds_bnb_kernels, "seconds_icp": self.seconds_icp,
                "bound_evals_executed_local": self.bound_evals_executed_local, "strict_resolves": self.strict_resolves,
                "contender_overflows": self.contender_overflows, "seconds_dt_score": self.seconds_dt_score,
                "seconds_strict": self.seconds_strict, "seconds_setup": self.seconds_setup, "seconds_host": self.seconds_host,
                "bnb_kernel_variants": self.bnb_kernel_variants}


class IcpResult(C.Structure):
    _fields_ = [("R", C.c_float * 9), ("t", C.c_float * 3), ("err", C.c_float), ("iterations", C.c_int)]


class Snapshot(C.Structure):
    _fields_ = [("R", C.c_float * 9), ("t", C.c_float * 3), ("sse", C.c_float),
                ("rot_pops", C.c_int64), ("trans_pops", C.c_int64), ("bound_evals", C.c_int64), ("finished", C.c_int)]


class InnerResult(C.Structure):
    _fields_ = [("value", C.c_float), ("node", C.c_float * 4), ("pops", C.c_uint32), ("evals", C.c_uint32), ("status", C.c_int32),
                ("reuse_gt", C.c_float), ("reuse_poplb", C.c_float)]


ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int)

# every symbol include/goicp_b200.h declares (tests/test_abi.py checks the header against this)
ABI_SYMBOLS = ["goicp_default_params", "goicp_create", "goicp_destroy", "goicp_last_error", "goicp_set_model",
               "goicp_set_data", "goicp_build_dt", "goicp_set_dt", "goicp_get_dt", "goicp_dt_size", "goicp_dt_distance",
               "goicp_eval_bounds", "goicp_expand_bounds", "goicp_inner_bnb", "goicp_nn", "goicp_kdtree_host", "goicp_svd3", "goicp_intro_select", "goicp_bnb_shape_rule", "goicp_icp", "goicp_icp_dt", "goicp_dt_score",
               "goicp_register", "goicp_poll", "goicp_cancel", "goicp_trim_memory", "goicp_transfer_bytes", "goicp_measure_gather", "goicp_set_exchange", "goicp_nccl_unique_id", "goicp_nccl_init", "goicp_selftest_shard", "goicp_run_toml", "goicp_load_cloud", "goicp_free_cloud", "goicp_io_last_error"]


def kdtree_host(model):
    """ICP3D::Build's kd-tree layout as the library builds it on the host (no device needed): nodes (n_nodes, 7) int32
    {child1, child2, left, right, divfeat, divlow bits, divhigh bits}, vind, root box (6 floats)."""
    m = np.ascontiguousarray(model, np.float32).reshape(-1, 3)
    L = lib()
    nn = L.goicp_kdtree_host(m.ctypes.data, len(m), None, 0, None, None)
    if nn < 0:
        raise GoicpError(-nn, "goicp_kdtree_host")
    nodes = np.zeros((nn, 7), np.int32); vind = np.zeros(len(m), np.int32); bbox = np.zeros(6, np.float32)
    L.goicp_kdtree_host(m.ctypes.data, len(m), nodes.ctypes.data, nn, vind.ctypes.data, bbox.ctypes.data)
    return nodes, vind, bbox


def build(verbose: bool = False) -> str:
    """Compile libgoicp_b200.so for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", _HERE, "all"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libgoicp_b200.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    if verbose:
        print(r.stdout)
    return LIB_PATH


_lib = None


def lib():
    """The loaded C-ABI library.  Raises if it has not been built -- there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(f"{LIB_PATH} is missing: run __graft_entry__.build() (make -C cuda-go-icp_b200)")
        L = C.CDLL(LIB_PATH)
        f32p = C.POINTER(C.c_float)
        L.goicp_default_params.argtypes = [C.POINTER(Params)]
        L.goicp_create.argtypes = [C.POINTER(Params), C.POINTER(C.c_void_p)]
        L.goicp_destroy.argtypes = [C.c_void_p]
        L.goicp_last_error.argtypes = [C.c_void_p]
        L.goicp_last_error.restype = C.c_char_p
        L.goicp_set_model.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.goicp_set_data.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.goicp_build_dt.argtypes = [C.c_void_p]
        L.goicp_set_dt.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.goicp_get_dt.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.goicp_dt_size.argtypes = [C.c_void_p]
        L.goicp_dt_distance.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.goicp_eval_bounds.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.goicp_expand_bounds.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_float)]
        L.goicp_inner_bnb.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.goicp_nn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.goicp_kdtree_host.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.goicp_svd3.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.goicp_intro_select.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.goicp_icp.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.POINTER(IcpResult)]
        L.goicp_icp_dt.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, f32p]
        L.goicp_dt_score.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, f32p]
        L.goicp_register.argtypes = [C.c_void_p, C.POINTER(Result)]
        L.goicp_poll.argtypes = [C.c_void_p, C.POINTER(Snapshot)]
        L.goicp_cancel.argtypes = [C.c_void_p]
        L.goicp_transfer_bytes.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.goicp_measure_gather.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_double)]
        L.goicp_set_exchange.argtypes = [C.c_void_p, ALLGATHER_FN, C.c_void_p, C.c_int]
        L.goicp_nccl_unique_id.argtypes = [C.c_void_p]
        L.goicp_nccl_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.goicp_selftest_shard.argtypes = [C.c_int, C.c_int, C.c_int, ALLGATHER_FN, C.c_void_p, C.POINTER(C.c_int)]
        L.goicp_load_cloud.argtypes = [C.c_char_p, C.c_float, C.c_float, C.c_uint, C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.c_int)]
        L.goicp_free_cloud.argtypes = [C.POINTER(C.c_float)]
        L.goicp_io_last_error.restype = C.c_char_p
        L.goicp_run_toml.argtypes = [C.c_char_p, C.c_uint, C.c_uint, C.POINTER(Result)]
        _lib = L
    return _lib


def nccl_unique_id() -> bytes:
    """128-byte ncclUniqueId (call on rank 0, broadcast to the other ranks, pass to GoICP.init_nccl)."""
    buf = np.zeros(128, np.uint8)
    rc = lib().goicp_nccl_unique_id(buf.ctypes.data)
    if rc != 0:
        raise GoicpError(rc, "ncclGetUniqueId failed (libnccl.so.2 not loadable?)")
    return buf.tobytes()


def _f32(a, cols=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if cols:
        a = a.reshape(-1, cols)
    return a


class _DT:
    """Mirror of GoICP::dt's public knobs (jly_3ddt.h:100-111)."""
    def __init__(self):
        self.SIZE = 300
        self.expandFactor = 2.0


class GoICP:
    """Mirror of the reference's ``class GoICP`` on top of the C ABI."""

    def __init__(self, mse_threshold: float = 1e-3, device: int = 0):
        self.L = lib()
        p = Params()
        self.L.goicp_default_params(C.byref(p))
        p.mse_threshold = mse_threshold
        p.device = device
        self._p = p
        self.MSEThresh = mse_threshold
        self.trimFraction = 0.0
        self.doTrim = True
        self.dt = _DT()
        self.dt_mode = p.dt_mode          # default: exact EDT with the reference binary's corner seed (include/goicp_b200.h)
        self.numerics = p.numerics
        self.search_mode = 0              # 0: GoICP::Register (reference-identical); 1: fgoicp-style strategy (include/goicp_b200.h)
        self.spec_cubes = 0
        self.cluster_size = 0
        self.initNodeRot = [p.rot_cube[i] for i in range(4)]
        self.initNodeTrans = [p.trans_cube[i] for i in range(4)]
        self.pModel = None
        self.pData = None
        self.optR = np.eye(3, dtype=np.float32)
        self.optT = np.zeros(3, np.float32)
        self.optError = 1e10
        self.finished = False
        self._h = None
        self._dt_pending = None
        self._keep = []
        self.rank, self.world_size = 0, 1
        self._exchange = None
        self._nccl_id = None

    # -- handle management ---------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            msg = self.L.goicp_last_error(self._h).decode() if self._h else ""
            raise GoicpError(rc, msg)

    def _handle(self):
        if self._h is None:
            p = self._p
            p.mse_threshold = self.MSEThresh
            p.trim_fraction = self.trimFraction
            p.do_trim = int(self.doTrim)
            p.dt_size = int(self.dt.SIZE)
            p.dt_expand = float(self.dt.expandFactor)
            p.dt_mode = int(self.dt_mode)
            p.numerics = int(self.numerics)
            p.search_mode = int(self.search_mode)
            p.spec_cubes = int(self.spec_cubes)
            p.cluster_size = int(self.cluster_size)
            p.rank, p.world_size = self.rank, self.world_size
            for i in range(4):
                p.rot_cube[i] = self.initNodeRot[i]
                p.trans_cube[i] = self.initNodeTrans[i]
            h = C.c_void_p()
            rc = self.L.goicp_create(C.byref(p), C.byref(h))
            if rc:
                raise GoicpError(rc, "goicp_create")
            self._h = h
            if self.pModel is not None:
                m = _f32(self.pModel, 3)
                self._check(self.L.goicp_set_model(self._h, m.ctypes.data, len(m)))
            if self.pData is not None:
                d = _f32(self.pData, 3)
                self._check(self.L.goicp_set_data(self._h, d.ctypes.data, len(d)))
            if self._exchange is not None:
                self._check(self.L.goicp_set_exchange(self._h, self._exchange, None, 0))
            if self._nccl_id is not None:
                self._check(self.L.goicp_nccl_init(self._h, self._nccl_id.ctypes.data, self.rank, self.world_size))
        return self._h

    def close(self):
        if self._h is not None:
            self.L.goicp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def Nm(self):
        return 0 if self.pModel is None else len(_f32(self.pModel, 3))

    @property
    def Nd(self):
        return 0 if self.pData is None else len(_f32(self.pData, 3))

    def init_nccl(self, unique_id, rank, world_size):
        """Native multi-GPU exchange: `unique_id` = the 128 bytes of nccl_unique_id() from rank 0 (collective at handle creation)."""
        self._nccl_id = np.frombuffer(bytes(unique_id), np.uint8).copy()
        assert self._nccl_id.size == 128
        self.rank, self.world_size = rank, world_size

    def set_exchange(self, fn, rank, world_size):
        """fn(send: bytes-like np.uint8 array, world) -> np.uint8 array of world*len(send) (rank-major)."""
        self.rank, self.world_size = rank, world_size

        def _cb(user, send, recv, nbytes, is_device):
            try:
                s = np.ctypeslib.as_array(C.cast(send, C.POINTER(C.c_uint8)), shape=(nbytes,))
                out = fn(s)
                r = np.ctypeslib.as_array(C.cast(recv, C.POINTER(C.c_uint8)), shape=(nbytes * world_size,))
                r[:] = np.asarray(out, np.uint8).reshape(-1)
                return 0
            except Exception as e:  # never raise through the C frame
                print("exchange callback failed:", e)
                return 1

        self._exchange = ALLGATHER_FN(_cb)

    # -- the reference's protocol --------------------------------------------------------------
    def BuildDT(self):
        """GoICP::BuildDT (jly_goicp.cpp:75-90), on the GPU."""
        self._check(self.L.goicp_build_dt(self._handle()))

    def SetDT(self, grid, meta):
        g = _f32(grid).reshape(-1)
        S = round(g.size ** (1 / 3))
        assert S * S * S == g.size
        m = np.ascontiguousarray(meta, np.float64)
        self._check(self.L.goicp_set_dt(self._handle(), g.ctypes.data, S, m.ctypes.data))

    def GetDT(self):
        h = self._handle()
        S = self.L.goicp_dt_size(h)
        g = np.zeros(S * S * S, np.float32)
        m = np.zeros(4, np.float64)
        self._check(self.L.goicp_get_dt(h, g.ctypes.data, m.ctypes.data))
        return g.reshape(S, S, S), m

    def Register(self):
        """GoICP::Register (jly_goicp.cpp:569-585); returns optError, fills optR/optT."""
        res = Result()
        rc = self.L.goicp_register(self._handle(), C.byref(res))
        self.result = res.as_dict()          # filled even when the call was cancelled
        self._check(rc)
        self.optR, self.optT, self.optError = self.result["R"], self.result["t"], self.result["sse"]
        self.finished = True
        return self.optError

    # -- pieces of the path, exposed for parity tests and the bench -----------------------------
    def Distance(self, q, with_index=False):
        """DT3D::Distance for a batch (jly_3ddt.cpp:981-1026)."""
        q = _f32(q, 3)
        out = np.zeros(len(q), np.float32)
        idx = np.zeros((len(q), 3), np.int32) if with_index else None
        self._check(self.L.goicp_dt_distance(self._handle(), q.ctypes.data, len(q), out.ctypes.data,
                                             idx.ctypes.data if with_index else None))
        return (out, idx) if with_index else out

    def EvalBounds(self, R, level, tcube):
        R = _f32(R).reshape(-1, 9)
        level = np.ascontiguousarray(level, np.int32)
        tcube = _f32(tcube, 4)
        n = len(R)
        ub = np.zeros(n, np.float32)
        lb = np.zeros(n, np.float32)
        self._check(self.L.goicp_eval_bounds(self._handle(), n, R.ctypes.data, level.ctypes.data, tcube.ctypes.data,
                                             ub.ctypes.data, lb.ctypes.data))
        return ub, lb

    def ExpandBounds(self, R, level, tcube, repeats=0):
        """8 children of each parent translation cube: returns (ub[n,8], lb[n,8][, ms per launch])."""
        R = _f32(R).reshape(-1, 9)
        level = np.ascontiguousarray(level, np.int32)
        tcube = _f32(tcube, 4)
        n = len(R)
        out = np.zeros((n, 16), np.float32)
        ms = C.c_float(0)
        self._check(self.L.goicp_expand_bounds(self._handle(), n, R.ctypes.data, level.ctypes.data, tcube.ctypes.data,
                                               out.ctypes.data, repeats, C.byref(ms) if repeats > 0 else None))
        return (out[:, :8], out[:, 8:], ms.value) if repeats > 0 else (out[:, :8], out[:, 8:])

    def InnerBnB(self, R, level, opt_error):
        """Batch of GoICP::InnerBnB calls (jly_goicp.cpp:227-340)."""
        R = _f32(R).reshape(-1, 9)
        level = np.ascontiguousarray(level, np.int32).reshape(-1)
        oe = _f32(opt_error).reshape(-1)
        n = len(R)
        out = (InnerResult * n)()
        self._check(self.L.goicp_inner_bnb(self._handle(), n, R.ctypes.data, level.ctypes.data, oe.ctypes.data, out))
        return [{"value": float(o.value), "node": np.array(o.node, np.float32), "pops": o.pops, "evals": o.evals,
                 "reuse_gt": np.float32(o.reuse_gt), "reuse_poplb": np.float32(o.reuse_poplb)} for o in out]

    def NN(self, q):
        q = _f32(q, 3)
        idx = np.zeros(len(q), np.int32)
        d2 = np.zeros(len(q), np.float32)
        self._check(self.L.goicp_nn(self._handle(), q.ctypes.data, len(q), idx.ctypes.data, d2.ctypes.data))
        return idx, d2

    def SVD3(self, H):
        """Matrix::svd of a batch of 3x3 matrices (matrix.cpp:602-830): returns U (n,3,3), W (n,3), V (n,3,3)."""
        H = _f32(H).reshape(-1, 9)
        n = len(H)
        U = np.zeros((n, 9), np.float32); W = np.zeros((n, 3), np.float32); V = np.zeros((n, 9), np.float32)
        self._check(self.L.goicp_svd3(self._handle(), H.ctypes.data, n, U.ctypes.data, W.ctypes.data, V.ctypes.data))
        return U.reshape(n, 3, 3), W, V.reshape(n, 3, 3)

    def IntroSelect(self, a, k, threads=512, in_global=False):
        """intro_select (jly_sorting.hpp:228-313) of a copy of `a` for position k by the strict kernels' block-wide select."""
        a = np.array(a, np.float32).reshape(-1).copy()
        self._check(self.L.goicp_intro_select(self._handle(), a.ctypes.data, len(a), int(k), int(threads), 1 if in_global else 0))
        return a

    def ICP(self, R0=None, t0=None, max_iter=0, err_diff=-1.0):
        """ICP3D<float>::Run (jly_icp3d.hpp:180-295)."""
        R0 = _f32(np.eye(3) if R0 is None else R0).reshape(9)
        t0 = _f32(np.zeros(3) if t0 is None else t0).reshape(3)
        r = IcpResult()
        self._check(self.L.goicp_icp(self._handle(), R0.ctypes.data, t0.ctypes.data, max_iter, err_diff, C.byref(r)))
        return float(r.err), np.array(r.R, np.float32).reshape(3, 3), np.array(r.t, np.float32), r.iterations

    def DTScore(self, R=None, t=None):
        out = C.c_float()
        if R is None:
            self._check(self.L.goicp_dt_score(self._handle(), None, None, C.byref(out)))
        else:
            R = _f32(R).reshape(9)
            t = _f32(t).reshape(3)
            self._check(self.L.goicp_dt_score(self._handle(), R.ctypes.data, t.ctypes.data, C.byref(out)))
        return out.value

    def last_result(self):
        return self.result

    def Poll(self):
        s = Snapshot()
        self.L.goicp_poll(self._handle(), C.byref(s))
        return {"R": np.array(s.R, np.float32).reshape(3, 3), "t": np.array(s.t, np.float32), "sse": s.sse,
                "rot_pops": s.rot_pops, "trans_pops": s.trans_pops, "bound_evals": s.bound_evals, "finished": bool(s.finished)}

    def Cancel(self):
        self.L.goicp_cancel(self._handle())

    def TransferBytes(self):
        """(host->device, device->host) bytes this handle has copied so far."""
        a, b = C.c_int64(0), C.c_int64(0)
        self.L.goicp_transfer_bytes(self._handle(), C.byref(a), C.byref(b))
        return a.value, b.value

    def MeasureGather(self, nbytes, repeats=5):
        """random 4-byte gather rate (look-ups/s) over an nbytes buffer: the measured ceiling of the DT gathers"""
        out = C.c_double(0)
        self._check(self.L.goicp_measure_gather(self._handle(), nbytes, repeats, C.byref(out)))
        return out.value


def load_cloud(path, subsample=1.0, resize=1.0, seed=1234):
    """load_cloud (src/common.cpp:205-228) through the C ABI, with a seeded subsample."""
    L = lib()
    p = C.POINTER(C.c_float)()
    n = C.c_int(0)
    rc = L.goicp_load_cloud(os.fsencode(path), subsample, resize, seed, C.byref(p), C.byref(n))
    if rc:
        raise GoicpError(rc, L.goicp_io_last_error().decode())
    out = np.ctypeslib.as_array(p, shape=(n.value * 3,)).copy().reshape(-1, 3) if n.value else np.zeros((0, 3), np.float32)
    L.goicp_free_cloud(p)
    return out


def run_toml(path, seed_model=1234, seed_data=1235):
    res = Result()
    rc = lib().goicp_run_toml(os.fsencode(path), seed_model, seed_data, C.byref(res))
    if rc:
        raise GoicpError(rc, lib().goicp_io_last_error().decode())
    return res.as_dict()
