"""Times ICP-only (BASELINE config 2, 40k x 40k) with both sorts; prints seconds and iterations."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("cuda-go-icp_b200")
gold = dict(np.load(os.path.join(ROOT, "tests", "golden", "bun_icp_config2.npz")))
for mode in ("1", "0"):
    os.environ["GOICP_ICP_RADIX"] = mode
    g = pkg.GoICP(1e-5)
    g.pModel, g.pData = gold["model"], gold["data"]
    g.ICP(np.eye(3), np.zeros(3), 3, 1e-9)
    t0 = time.perf_counter()
    err, R, t, iters = g.ICP(np.eye(3), np.zeros(3), 10000, 1e-9)
    dt = time.perf_counter() - t0
    print("radix" if mode == "1" else "count", "ICP 40256x40097: %.4f s, %d iterations, %.1f us/iter, err %.5f, exact=%s" %
          (dt, iters, 1e6 * dt / max(iters, 1), err, np.array_equal(R, gold["icp_R"])), flush=True)
    g.close()
