"""ICP phase cycles on the bunny Go-ICP config (GOICP_ICP_STATS=1): the first ICP of Register (from the identity), strict and fast."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("cuda-go-icp_b200")
G = os.path.join(ROOT, "tests", "golden")
ld = lambda n: np.fromfile(os.path.join(G, n), np.float32).reshape(-1, 3)
small = dict(np.load(os.path.join(G, "small_vectors.npz")))
for numerics in (0, 2):
    g = pkg.GoICP(1e-3); g.pModel, g.pData = ld("bunny_model_s0.1_seed1234.f32"), ld("bunny_data_s0.1_seed1235.f32"); g.numerics = numerics
    g.ICP(np.eye(3), np.zeros(3), 10000, 1e-7)
    t0 = time.perf_counter()
    err, R, t, iters = g.ICP(np.eye(3), np.zeros(3), 10000, 1e-7)
    dt = time.perf_counter() - t0
    print("numerics %d: %.3f ms, %d iterations, %.1f us/iter, bit-exact vs reference: %s" % (numerics, 1e3 * dt, iters, 1e6 * dt / iters,
          bool(np.float32(err) == small["icp_trim0.0_err"] and np.array_equal(R, small["icp_trim0.0_R"]) and np.array_equal(t, small["icp_trim0.0_t"]))), flush=True)
    g.close()
