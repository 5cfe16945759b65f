"""One committed golden registration (tests/golden/goicp_runs.json) through the C ABI, timed: python scripts/golden_run.py NAME [REPS].
GOICP_ROUND_STATS=1 adds the per-round breakdown of the inner-BnB kernels on stderr."""
import importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("cuda-go-icp_b200")
G = os.path.join(ROOT, "tests", "golden")
runs = json.load(open(os.path.join(G, "goicp_runs.json")))
name = sys.argv[1]; reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
gold = runs[name]
ld = lambda n: np.fromfile(os.path.join(G, n), np.float32).reshape(-1, 3)
g = pkg.GoICP(gold["mse"]); g.pModel, g.pData = ld(gold["model"]), ld(gold["data"]); g.trimFraction = gold["trim"]
if "trans_cube" in gold:
    g.initNodeTrans = gold["trans_cube"]
g.numerics = int(os.environ.get("NUMERICS", "0")); g.spec_cubes = int(os.environ.get("SPEC", "0")); g.cluster_size = int(os.environ.get("CLUSTER", "0"))
g.BuildDT()
for rep in range(reps):
    t0 = time.perf_counter(); g.Register(); dt = time.perf_counter() - t0
    r = g.result
    print("%s: %.4f s  sse %.6f (golden %.6f)  pops %d / %d (golden %d / %d)  evals %d executed %d  rounds %d  exit %s | bnb %.4f icp %.4f score %.4f strict %.4f host %.4f" % (
        name, dt, r["sse"], gold["sse"], r["rot_pops"], r["trans_pops"], gold["rot_pops"], gold["trans_pops"], r["bound_evals"], r["bound_evals_executed"], r["rounds"], r["exit_path"],
        r["seconds_bnb_kernels"], r["seconds_icp"], r["seconds_dt_score"], r["seconds_strict"], r["seconds_host"]), flush=True)
g.close()
