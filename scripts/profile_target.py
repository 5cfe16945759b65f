"""Short single-GPU program for ncu: one bunny Go-ICP registration in the library's default configuration (DT build
included) + batched DT-gather launches (expand_bounds_kernel) + the uniform random-gather microbenchmark."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("cuda-go-icp_b200")
G = os.path.join(ROOT, "tests", "golden")
model = np.fromfile(os.path.join(G, "bunny_model_s0.1_seed1234.f32"), np.float32).reshape(-1, 3)
data = np.fromfile(os.path.join(G, "bunny_data_s0.1_seed1235.f32"), np.float32).reshape(-1, 3)
g = pkg.GoICP(1e-3)
g.pModel, g.pData = model, data
if "PROFILE_GOLDEN" in os.environ:                    # any committed golden run instead (e.g. spanner_s0.02_mse3e-4 with GOICP_BNB_VARIANT=q5: the dense shape)
    import json
    gold = json.load(open(os.path.join(G, "goicp_runs.json")))[os.environ["PROFILE_GOLDEN"]]
    g.close()
    g = pkg.GoICP(gold["mse"])
    g.pModel, g.pData = [np.fromfile(os.path.join(G, gold[k]), np.float32).reshape(-1, 3) for k in ("model", "data")]
    g.trimFraction = gold["trim"]
    if "trans_cube" in gold:
        g.initNodeTrans = gold["trans_cube"]
if "DT_MODE" in os.environ:
    g.dt_mode = int(os.environ["DT_MODE"])          # default: the library's (exact EDT + reference corner seed)
g.BuildDT()
g.Register()
print({k: v for k, v in g.result.items() if k not in ("R", "t")})
rng = np.random.default_rng(7)
n = 148 * 16
Rs = np.stack([np.linalg.qr(rng.normal(size=(3, 3)))[0] for _ in range(n)]).astype(np.float32)
tc = np.concatenate([rng.uniform(-0.5, 0.25, (n, 3)), np.full((n, 1), 0.25)], 1).astype(np.float32)
print(g.ExpandBounds(Rs.reshape(n, 9), np.full(n, -1, np.int32), tc, repeats=3)[2])
print(g.MeasureGather(4 * 300 ** 3, 2))
g.close()
