"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the golden vectors.

Tolerances (BASELINE.json north_star): DT voxel indices and NN indices bit-exact; final R within
1e-4 rad, t within 1e-4 units, SSE within 1e-5 relative, same certificate / exit path.
Per-evaluation bound sums are float32 sums taken in a different (parallel, fixed) order than the
reference's sequential loop, hence 2e-6 relative there.
"""
import numpy as np
import pytest

from conftest import rot_angle

pytestmark = pytest.mark.gpu

SUM_RTOL = 4e-6
DT_REFERENCE, DT_EXACT_EDT, DT_EXACT_EDT_REFSEED = 0, 1, 2      # goicp_dt_mode


def _close_counts(a, b):
    return abs(a - b) <= max(2, 0.005 * b)


@pytest.fixture(scope="module")
def eng_small(pkg, small, bunny):
    """engine on the golden S=64 inner-BnB fixture (data = every 2nd point of the 0.033 subsample)"""
    g = pkg.GoICP(1e-3)
    g.pModel, g.pData = bunny["model_s"], bunny["data_s"][::2].copy()
    g.dt.SIZE = 64
    g.SetDT(small["inner_grid"], small["inner_meta"])
    yield g
    g.close()


def test_dt_distance_and_voxel_indices_bit_exact(pkg, small, restated, bunny):
    g = pkg.GoICP(1e-3)
    g.pModel, g.pData = bunny["model"], bunny["data"]
    g.SetDT(small["dt48_grid"], small["dt48_meta"])
    d, idx = g.Distance(small["dt48_query"], with_index=True)
    assert np.array_equal(d.view(np.uint32), small["dt48_dist"].view(np.uint32))
    dt = restated.dt_wrap(small["dt48_grid"], 48, small["dt48_meta"])
    assert np.array_equal(idx, restated.dt_index(dt, small["dt48_query"]))
    # raw-data score = initial error of OuterBnB
    ref = restated.dt_distance(dt, bunny["data"]).astype(np.float64)
    assert g.DTScore() == pytest.approx(float((ref ** 2).sum()), rel=SUM_RTOL)
    g.close()


def test_eval_bounds_pairs(eng_small, small, restated, bunny):
    data = bunny["data_s"][::2].copy()
    dt = restated.dt_wrap(small["inner_grid"], 64, small["inner_meta"])
    cases = small["inner_cases"]
    rng = np.random.default_rng(5)
    R, lvl, tc = [], [], []
    for row in cases[:16]:
        for _ in range(4):
            w = 2.0 ** -rng.integers(0, 6)
            R.append(row[:9]); lvl.append(int(row[9]))
            tc.append([rng.uniform(-0.5, 0.5 - w), rng.uniform(-0.5, 0.5 - w), rng.uniform(-0.5, 0.5 - w), w])
    R = np.array(R, np.float32); lvl = np.array(lvl, np.int32); tc = np.array(tc, np.float32)
    ub, lb = eng_small.EvalBounds(R, lvl, tc)
    gam = small["inner_gamma"]
    for k in range(len(R)):
        Rk = R[k].reshape(3, 3)
        p = data
        # float32 op order of jly_goicp.cpp:470-476
        rot = np.stack([(Rk[i, 0] * p[:, 0] + Rk[i, 1] * p[:, 1]) + Rk[i, 2] * p[:, 2] for i in range(3)], 1).astype(np.float32)
        t = (tc[k, :3] + tc[k, 3] / np.float32(2)).astype(np.float32)
        d = restated.dt_distance(dt, (rot + t).astype(np.float32))
        if lvl[k] >= 0:
            d = d - gam[lvl[k]]
        d = np.maximum(d, 0).astype(np.float32)
        gt = np.float32(1.732050808 / 2.0 * float(tc[k, 3]))
        e = np.maximum(d - gt, 0).astype(np.float32)
        assert ub[k] == pytest.approx(float((d.astype(np.float64) ** 2).sum()), rel=SUM_RTOL, abs=1e-7)
        assert lb[k] == pytest.approx(float((e.astype(np.float64) ** 2).sum()), rel=SUM_RTOL, abs=1e-7)


def test_inner_bnb_matches_reference_known_answers(eng_small, small):
    cases = small["inner_cases"]
    out = eng_small.InnerBnB(cases[:, :9], cases[:, 9].astype(np.int32), cases[:, 10].astype(np.float32))
    for row, o in zip(cases, out):
        if int(row[9]) < 0:
            # upper-bound pass: value and arg-min cube are resolved in the reference's summation order
            assert np.float32(o["value"]) == np.float32(row[11])
            if row[11] < row[10]:
                assert np.array_equal(o["node"], row[12:16].astype(np.float32))
        else:
            assert o["value"] == pytest.approx(row[11], rel=1e-5, abs=1e-6)
        # the search itself runs on fixed-order tree sums: a borderline prune may differ by a node or two
        assert _close_counts(o["pops"], int(row[16])) and _close_counts(o["evals"], int(row[17]))


def test_nn_indices_bit_exact(pkg, small, bunny):
    g = pkg.GoICP(1e-3)
    g.pModel, g.pData = bunny["model"], bunny["data"]
    idx, d2 = g.NN(bunny["data"])
    assert np.array_equal(idx, small["nn_idx"])
    assert np.array_equal(d2.view(np.uint32), small["nn_d2"].view(np.uint32))
    g.close()
    # exact ties: lattice with duplicated points, half-integer queries
    g = pkg.GoICP(1e-3)
    g.pModel, g.pData = small["lat_model"], small["lat_query"]
    idx, d2 = g.NN(small["lat_query"])
    assert np.array_equal(idx, small["lat_idx"])
    assert np.array_equal(d2, small["lat_d2"])
    g.close()


def _bumpy_surface(n, seed):
    rng = np.random.default_rng(seed)
    u = rng.normal(size=(n, 3)); u /= np.linalg.norm(u, axis=1, keepdims=True)
    k = rng.normal(size=(4, 3)); ph = rng.uniform(0, 2 * np.pi, 4)
    r = 0.33 + sum(0.04 * np.sin(3 * (u @ kk) + p) for kk, p in zip(k, ph))
    return (u * r[:, None]).astype(np.float32)


@pytest.mark.parametrize("search", ["walk", "cooperative"])
def test_nn_far_queries_large_model_vs_oracle(pkg, restated, search, monkeypatch):
    """Both device searches -- the per-thread reference walk with subtree skipping, and the warp-cooperative
    frontier search with its tie rules (GOICP_NN_COOP=1) -- against the oracle's plain reference traversal:
    a model large enough for the tree path (60k surface points + a lattice block with duplicated points),
    queries far from the surface (where the reference walks thousands of leaves), near it, on model points,
    and at half-integer lattice positions (exact ties: the traversal ORDER decides the index)."""
    rng = np.random.default_rng(11)
    surf = _bumpy_surface(60000, 5)
    ax = np.arange(8, dtype=np.float32) * 0.0625 + 0.55
    lat = np.stack(np.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(-1, 3)
    model = np.concatenate([surf, lat, lat[::3]]).astype(np.float32)
    q = np.concatenate([rng.uniform(-1.2, 1.2, (6000, 3)),                                   # far
                        surf[rng.choice(len(surf), 4000)] + rng.normal(scale=2e-3, size=(4000, 3)),  # near
                        model[rng.choice(len(model), 1000)],                                 # on model points (d = 0, duplicates)
                        lat[rng.choice(len(lat), 1000)] + np.float32(0.03125),               # cell centres: 8-way exact ties
                        np.zeros((1, 3))]).astype(np.float32)                                 # centre of the surface: near-ties all around
    if search == "cooperative":
        monkeypatch.setenv("GOICP_NN_COOP", "1")
    else:
        monkeypatch.delenv("GOICP_NN_COOP", raising=False)
    kd = restated.kd_build(model)
    ref_idx, ref_d2 = restated.kd_nn(kd, q)
    g = pkg.GoICP(1e-3)
    g.pModel, g.pData = model, q
    idx, d2 = g.NN(q)
    g.close()
    assert np.array_equal(d2.view(np.uint32), ref_d2.view(np.uint32))
    assert np.array_equal(idx, ref_idx)


@pytest.mark.parametrize("budget", ["1", "0", "96", "12"])
def test_icp_large_model_tree_path_vs_oracle(pkg, restated, budget, monkeypatch):
    """ICP3D::Run through the tree-search branch of the ICP kernel (model > 16384 points) from a pose
    40 mrad / 0.05 off: every iteration's correspondences, sort and sequential sums follow the oracle,
    so R, t and the error agree bit for bit -- with every query answered by the warp-cooperative search (1, the
    default), with the per-thread walk only (0 = unlimited), and with mixes (96, 12)."""
    monkeypatch.setenv("GOICP_NN_BUDGET", budget)
    model = _bumpy_surface(50000, 7)
    rng = np.random.default_rng(3)
    data = (model[rng.choice(len(model), 6000, replace=False)] + rng.normal(scale=1e-3, size=(6000, 3))).astype(np.float32)
    a = 0.04
    R0 = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]], np.float32)
    t0 = np.array([0.05, -0.03, 0.02], np.float32)
    kd = restated.kd_build(model)
    e_ref, R_ref, t_ref, it_ref, _ = restated.icp_run(kd, data, R0, t0, 10000, 1e-7)
    g = pkg.GoICP(1e-3)
    g.pModel, g.pData = model, data
    err, R, t, iters = g.ICP(R0, t0, 10000, 1e-7)
    g.close()
    assert np.array_equal(R, R_ref) and np.array_equal(t, t_ref)
    assert np.float32(err) == np.float32(e_ref)


@pytest.mark.parametrize("sort", ["count", "radix"])
def test_icp_matches_reference(pkg, small, bunny, sort, monkeypatch):
    # both sorts of the ICP kernel (rank by counting / grid-wide stable radix sort) on the same input
    monkeypatch.setenv("GOICP_ICP_RADIX", "1" if sort == "radix" else "0")
    g = pkg.GoICP(1e-3)
    g.pModel, g.pData = bunny["model"], bunny["data"]
    err, R, t, iters = g.ICP(np.eye(3), np.zeros(3), 10000, 1e-7)
    # the device ICP follows the reference's arithmetic step by step (sorted, sequential float sums)
    assert np.float32(err) == small["icp_trim0.0_err"]
    assert np.array_equal(R, small["icp_trim0.0_R"]) and np.array_equal(t, small["icp_trim0.0_t"])
    g.close()
    g = pkg.GoICP(1e-3)
    g.pModel, g.pData = bunny["model"], bunny["data"]
    g.trimFraction = 0.1
    err, R, t, iters = g.ICP(np.eye(3), np.zeros(3), 10000, 1e-7)
    assert err == pytest.approx(float(small["icp_trim0.1_err"]), rel=1e-5)
    assert rot_angle(R, small["icp_trim0.1_R"]) < 1e-4 and np.abs(t - small["icp_trim0.1_t"]).max() < 1e-4
    g.close()


@pytest.mark.parametrize("S", [48])
def test_dt_build_reference_mode_bit_exact_small(pkg, small, bunny, S):
    g = pkg.GoICP(1e-3)
    g.pModel, g.pData = bunny["model"], bunny["data"]
    g.dt.SIZE = S
    g.dt_mode = DT_REFERENCE
    g.BuildDT()
    grid, meta = g.GetDT()
    assert np.array_equal(meta, small["dt48_meta"])
    assert np.array_equal(grid.view(np.uint32), small["dt48_grid"].view(np.uint32))
    g.close()


@pytest.mark.parametrize("S,unsplit", [(31, False), (61, False), (61, True), (100, True), (95, False)])
def test_dt_build_reference_mode_vs_oracle_odd_sizes_and_both_kernels(pkg, restated, bunny, S, unsplit, monkeypatch):
    """Odd grid sizes (padded row pitch), several warps with ragged tails, and both propagation
    kernels (warp-specialised / single-role) against the restated DT3D::Build, bit for bit."""
    if unsplit:
        monkeypatch.setenv("GOICP_DT_UNSPLIT", "1")
    g = pkg.GoICP(1e-3)
    g.pModel, g.pData = bunny["model_s"], bunny["data_s"]
    g.dt.SIZE = S
    g.dt_mode = DT_REFERENCE
    g.BuildDT()
    grid, meta = g.GetDT()
    g.close()
    dt = restated.dt_build(bunny["model_s"], S)
    want = restated.dt_grid(dt, S)
    restated.dt_free(dt)
    assert np.array_equal(meta, restated.dt_frame(bunny["model_s"], S))
    assert np.array_equal(grid.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("S,mode", [(48, DT_EXACT_EDT), (61, DT_EXACT_EDT_REFSEED), (100, DT_EXACT_EDT_REFSEED), (333, DT_EXACT_EDT_REFSEED)])
def test_dt_build_exact_edt_modes_vs_scipy(pkg, small, bunny, S, mode):
    """The exact-EDT builder (in-place Meijster passes; the default mode seeds the reference binary's extra voxel (0,0,0) too)
    against scipy's exact transform of the same seed set, bit for bit in the reference's float metric; S = 333 takes the
    512-entry variant of the line kernel."""
    from scipy import ndimage
    g = pkg.GoICP(1e-3)
    g.pModel, g.pData = bunny["model"], bunny["data"]
    g.dt.SIZE = S
    g.dt_mode = mode
    g.BuildDT()
    grid, meta = g.GetDT()
    g.close()
    idx = np.floor((bunny["model"].astype(np.float64) - meta[:3]) * meta[3] + 0.5).astype(int)
    idx = idx[((idx >= 0) & (idx < S)).all(1)]
    occ = np.ones((S, S, S), bool)
    occ[idx[:, 2], idx[:, 1], idx[:, 0]] = False
    if mode == DT_EXACT_EDT_REFSEED:
        occ[0, 0, 0] = False
    exact = ndimage.distance_transform_edt(occ)
    want = (np.sqrt((exact ** 2).round()).astype(np.float32).astype(np.float64) / meta[3]).astype(np.float32)
    assert np.array_equal(grid, want)


def test_default_dt_differs_from_reference_order_only_where_that_is_inexact(pkg, restated, bunny):
    """Default DT (exact EDT of the reference binary's seed set) vs the reference-order propagation on the same model: never
    above it (the propagation only ever over-estimates), identical on all but a ~1e-5 fraction of the voxels."""
    S = 100
    grids = {}
    for mode in (DT_REFERENCE, DT_EXACT_EDT_REFSEED):
        g = pkg.GoICP(1e-3)
        g.pModel, g.pData = bunny["model"], bunny["data"]
        g.dt.SIZE = S
        g.dt_mode = mode
        g.BuildDT()
        grids[mode], meta = g.GetDT()
        g.close()
    a, b = grids[DT_REFERENCE], grids[DT_EXACT_EDT_REFSEED]
    assert (b <= a).all()
    assert (a != b).mean() < 2e-4 and np.abs(a - b).max() * meta[3] < 0.5


# runs whose committed counters equal the reference's exactly (the others differ by a node or two: the search kernels prune on
# fixed-order tree sums, the reference on its sequential ones -- bunny 5e-4: 225 970 vs 225 968 translation pops)
EXACT_COUNTS = {"bunny_s0.1_mse1e-3", "bunny_s0.1_mse7e-4", "bunny_s0.033_mse1e-3", "bunny_s0.1_mse1e-3_trim0.1", "bunny_s0.3_mse1e-3", "spanner_s0.02_mse1e-3",
                "spanner_s0.02_mse1e-3_trim0.1", "spanner_s0.02_mse3e-4", "spanner_s0.02_mse3e-4_trim0.1", "skull_s0.03_mse1e-3", "face_s0.025_mse1e-3"}


def _check_run(res, gold, name=None):
    assert res["exit_path"] == gold["exit_path"]
    if name in EXACT_COUNTS:
        assert (res["rot_pops"], res["trans_pops"]) == (gold["rot_pops"], gold["trans_pops"])
    assert _close_counts(res["rot_pops"], gold["rot_pops"]) and _close_counts(res["trans_pops"], gold["trans_pops"])
    assert res["sse"] == pytest.approx(gold["sse"], rel=1e-5)
    assert rot_angle(res["R"], np.array(gold["R"]).reshape(3, 3)) < 1e-4
    assert np.abs(res["t"] - np.array(gold["t"])).max() < 1e-4
    if gold["exit_path"] == "certified":
        assert res["best_lb"] == pytest.approx(gold["exit_lb"], rel=1e-4, abs=1e-5)


def test_register_bunny_full_size_reference_dt_on_gpu(pkg, runs, bunny, restated):
    """BASELINE config 1 end to end on the GPU: build the S=300 DT (reference mode), check its
    checksum against the reference's, run Go-ICP, compare with the reference's own run."""
    g = pkg.GoICP(1e-3)
    g.pModel, g.pData = bunny["model"], bunny["data"]
    g.dt_mode = DT_REFERENCE
    g.BuildDT()
    grid, meta = g.GetDT()
    assert np.array_equal(meta, [-1.7215785086154938, -1.735278993844986, -1.7175954878330231, 87.099870706007081])
    assert "%016x" % restated.fnv(grid) == "2da64ee1865a968e"
    gold = runs["bunny_s0.1_mse1e-3"]
    g.Register()
    _check_run(g.result, gold, "bunny_s0.1_mse1e-3")
    # certificate paths reuse the same DT
    for name in ("bunny_s0.1_mse7e-4", "bunny_s0.1_mse5e-4"):
        gold = runs[name]
        g2 = pkg.GoICP(gold["mse"])
        g2.pModel, g2.pData = bunny["model"], bunny["data"]
        g2.SetDT(grid, meta)
        g2.Register()
        _check_run(g2.result, gold, name)
        g2.close()
    g.close()


def _golden_engine(pkg, gold):
    from conftest import load_cloud
    g = pkg.GoICP(gold["mse"])
    g.pModel, g.pData = load_cloud(gold["model"]), load_cloud(gold["data"])
    g.trimFraction = gold["trim"]
    if "trans_cube" in gold:
        g.initNodeTrans = gold["trans_cube"]
    return g


@pytest.mark.parametrize("name", ["bunny_s0.1_mse1e-3", "bunny_s0.1_mse7e-4", "bunny_s0.1_mse5e-4", "bunny_s0.033_mse1e-3", "bunny_s0.3_mse1e-3",
                                  "bunny_s0.1_mse1e-3_trim0.1", "spanner_s0.02_mse1e-3", "spanner_s0.02_mse1e-3_trim0.1",
                                  "spanner_s0.02_mse3e-4", "spanner_s0.02_mse3e-4_trim0.1", "spanner_s0.02_mse1e-4",
                                  "skull_s0.03_mse1e-3", "face_s0.025_mse1e-3"])
def test_every_golden_run_in_the_default_mode(pkg, runs, name):
    """All of the reference's own S = 300 runs (tests/golden/goicp_runs.json, generated from oracle/_ref) in the library's
    DEFAULT configuration -- exact-EDT distance transform, reference-order sums and ICP: pose, SSE, exit path, certificate
    and counters.  Includes the deepest committed run (bunny 0.033: 2 032 rotation pops, 3.5 M bound evaluations, rotation
    level 8), SURVEY section 6's sub-0.3 known answer (Nd 9 064), and BASELINE config 4's noisy spanner pair below its noise
    floor, where the BnB really searches: mse 3e-4 with trimming 0 / 0.1 (89 / 218 rotation pops) and the TOML's own mse 1e-4
    (test/spanner_goicp.toml:15-20: 5 305 rotation pops, 154 M bound evaluations, certified)."""
    if name not in runs:
        pytest.skip(f"{name} not in goicp_runs.json yet")
    gold = runs[name]
    g = _golden_engine(pkg, gold)
    g.BuildDT()
    g.Register()
    _check_run(g.result, gold, name)
    assert g.result["contender_overflows"] == 0
    g.close()


def test_numerics_modes_table(pkg, runs):
    """The measurement behind the defaults (scripts/parity_modes.py, profiles/r2_parity_modes.md), asserted:
      * exact-EDT DT: every golden run ends exactly where the reference-order DT ends -- same pose bits, same counters;
      * tree sums instead of the reference-order re-evaluations: inside the north-star tolerances on every run;
      * ICP with parallel sums + Jacobi solver: voxel and NN indices still exact, poses within 1e-3, but the SSE leaves the
        1e-5 band (ICP stops one iteration earlier or later on its loose test) -- which is why strict ICP stays the default."""
    names = ["bunny_s0.1_mse1e-3", "bunny_s0.1_mse7e-4", "bunny_s0.1_mse1e-3_trim0.1", "skull_s0.03_mse1e-3", "face_s0.025_mse1e-3"]
    res = {}
    for mode, (dt_mode, numerics) in {"strict": (0, 0), "default": (2, 0), "fastsums": (2, 1), "fasticp": (2, 2)}.items():
        for name in names:
            g = _golden_engine(pkg, runs[name])
            g.dt_mode, g.numerics = dt_mode, numerics
            g.BuildDT()
            g.Register()
            res[mode, name] = g.result
            g.close()
    fasticp_sse_off = 0
    for name in names:
        gold = runs[name]
        a, b = res["strict", name], res["default", name]
        assert np.array_equal(a["R"], b["R"]) and np.array_equal(a["t"], b["t"]) and a["sse"] == b["sse"]
        assert (a["rot_pops"], a["trans_pops"], a["bound_evals"], a["exit_path"]) == (b["rot_pops"], b["trans_pops"], b["bound_evals"], b["exit_path"])
        _check_run(res["fastsums", name], gold)
        f = res["fasticp", name]
        assert f["exit_path"] == gold["exit_path"]
        assert rot_angle(f["R"], np.array(gold["R"]).reshape(3, 3)) < 1e-3 and np.abs(f["t"] - np.array(gold["t"])).max() < 1e-3
        assert f["sse"] == pytest.approx(gold["sse"], rel=5e-3)
        fasticp_sse_off += abs(f["sse"] - gold["sse"]) > 1e-5 * gold["sse"]
    assert fasticp_sse_off >= 1


def test_icp_overlapped_with_dt_build_changes_nothing(pkg, bunny, monkeypatch):
    """goicp_build_dt runs Register's first ICP (from the identity, no DT needed) next to the DT build; the
    registration that follows must be the one obtained without the overlap, field for field -- and a second
    Register on the same handle (which has to run that ICP itself) as well."""
    def run(no_prefetch):
        if no_prefetch:
            monkeypatch.setenv("GOICP_NO_PREFETCH", "1")
        else:
            monkeypatch.delenv("GOICP_NO_PREFETCH", raising=False)
        g = pkg.GoICP(1e-3)
        g.pModel, g.pData = bunny["model_s"], bunny["data_s"]
        g.dt.SIZE = 100
        g.dt_mode = DT_REFERENCE
        g.BuildDT()
        g.Register(); a = dict(g.result)
        g.Register(); b = dict(g.result)
        grid, meta = g.GetDT()
        g.close()
        return a, b, grid, meta
    a0, b0, grid0, meta0 = run(True)
    a1, b1, grid1, meta1 = run(False)
    assert np.array_equal(grid0, grid1) and np.array_equal(meta0, meta1)
    for x in (a1, b1, b0):
        for k in ("sse", "rot_pops", "trans_pops", "bound_evals", "icp_calls", "exit_path"):
            assert x[k] == a0[k], k
        assert np.array_equal(x["R"], a0["R"]) and np.array_equal(x["t"], a0["t"])


@pytest.mark.parametrize("nm,nd,S,mse,tcube", [(12, 5, 24, 5e-3, None), (700, 1, 32, 5e-3, None), (40, 33, 32, 2e-2, None),
                                              (1186, 257, 40, 3e-3, None), (300, 100, 31, 1e-2, (-1.0, -1.0, -1.0, 2.0))])
def test_tiny_and_ragged_clouds_vs_oracle(pkg, restated, bunny, nm, nd, S, mse, tcube):
    """Edge sizes: fewer points than a warp, a single data point, counts that are no multiple of anything, odd
    grid sizes, the [-1,1]^3 translation domain of the Artec TOMLs -- whole registrations (several ICP calls,
    certified exits after thousands of rotation pops) against the oracle on the same inputs."""
    rng = np.random.default_rng(nm * 1000 + nd)
    m = bunny["model_s"][rng.choice(len(bunny["model_s"]), nm, replace=False)].copy()
    d = bunny["data_s"][rng.choice(len(bunny["data_s"]), nd, replace=False)].copy()
    o = restated.create(m, d, mse, 0.0, S, 2.0, tcube)
    restated.L.go_build_dt(o)
    ref = restated.register(o)
    g = pkg.GoICP(mse)
    g.pModel, g.pData = m, d
    g.dt.SIZE = S
    g.dt_mode = DT_REFERENCE
    if tcube is not None:
        g.initNodeTrans = tcube
    g.BuildDT()
    grid, _ = g.GetDT()
    assert np.array_equal(restated.dt_grid(restated.L.go_get_dt(o), S), grid)
    g.Register()
    r = g.result
    g.close()
    assert r["exit_path"] == ref["exit_path"]
    assert r["sse"] == pytest.approx(ref["sse"], rel=1e-5, abs=1e-7)
    assert rot_angle(r["R"], ref["R"]) < 1e-4 and np.abs(r["t"] - ref["t"]).max() < 1e-4
    assert r["icp_calls"] == ref["icp_calls"] and _close_counts(r["rot_pops"], ref["rot_pops"])


@pytest.mark.parametrize("nd,trim", [(130000, 0.0), (130000, 0.1), (70001, 0.0)])
def test_dt_score_of_a_large_cloud_in_reference_order(pkg, restated, nd, trim):
    """GoICP::ICP's DT re-scoring (jly_goicp.cpp:100-131) for a data cloud too large for shared memory: the residuals
    live in global memory, and the value must still be the reference's -- intro_select's permutation followed by the
    sequential float sum -- bit for bit.  Oracle = its DT lookups + its intro_select + a sequential float32 sum."""
    rng = np.random.default_rng(nd)
    model = _bumpy_surface(30000, 9)
    data = (model[rng.choice(len(model), nd)] + rng.normal(scale=3e-3, size=(nd, 3))).astype(np.float32)
    a = 0.05
    R = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]], np.float32)
    t = np.array([0.02, -0.01, 0.03], np.float32)
    S = 64
    dt = restated.dt_build(model, S)
    g = pkg.GoICP(1e-3)
    g.pModel, g.pData = model, data
    g.dt.SIZE = S
    g.trimFraction = trim
    g.SetDT(restated.dt_grid(dt, S), restated.dt_meta(dt))
    for pose in (None, (R, t)):
        if pose is None:
            q = data
            got = g.DTScore()
        else:
            f = np.float32
            q = np.stack([((R[r, 0] * data[:, 0] + R[r, 1] * data[:, 1]) + R[r, 2] * data[:, 2]) + t[r] for r in range(3)], 1).astype(f)
            got = g.DTScore(R, t)
        d = restated.dt_distance(dt, q)
        num = int(np.float32(nd) * np.float32(1 - np.float32(trim))) if trim > 0 else nd
        sel = restated.intro_select(d, num - 1)
        want = np.add.accumulate((sel[:num] * sel[:num]).astype(np.float32), dtype=np.float32)[-1]
        assert np.float32(got) == want
    g.close()


def test_run_toml_end_to_end(pkg, runs, bunny, tmp_path):
    """the reference's own workflow: TOML -> load clouds -> DT -> Go-ICP -> output file, on the GPU"""
    for name, arr in (("model.txt", bunny["model"]), ("data.txt", bunny["data"])):
        with open(tmp_path / name, "w") as f:
            f.write(f"{len(arr)}\n")
            for q in arr:
                f.write("%.9g %.9g %.9g\n" % tuple(q))
    cfg = tmp_path / "bunny.toml"
    out = tmp_path / "output.toml"
    cfg.write_text(f'[io]\ntarget = "model.txt"\nsource = "data.txt"\noutput = "{out}"\nvisualization = ""\n'
                   '[params]\nmode = 3\ntrim = true\nsubsample = 1.0\nmse_threshold = 1e-3\nresize = 1.0\n')
    res = pkg.run_toml(str(cfg))
    _check_run(res, runs["bunny_s0.1_mse1e-3"], "bunny_s0.1_mse1e-3")
    text = out.read_text()
    assert "exit_path = \"early_sse_below_thresh\"" in text and "rotation_nodes = 206" in text


def test_trimmed_inner_bnb_and_register(pkg, runs, bunny, restated, small):
    """trimFraction = 0.1 (the reference's outlier-robust variant, jly_goicp.cpp:293-315): inner BnBs
    against the oracle on a coarse DT, then the full S=300 bunny run against the reference's own."""
    data = bunny["data_s"][::2].copy()
    o = restated.create(bunny["model_s"], data, 1e-3, 0.1, 64)
    restated.L.go_set_dt(o, restated.dt_wrap(small["inner_grid"], 64, small["inner_meta"]))
    restated.L.go_initialize(o)
    g = pkg.GoICP(1e-3)
    g.pModel, g.pData = bunny["model_s"], data
    g.trimFraction = 0.1
    g.dt.SIZE = 64
    g.SetDT(small["inner_grid"], small["inner_meta"])
    cases = small["inner_cases"][:24]
    out = g.InnerBnB(cases[:, :9], cases[:, 9].astype(np.int32), cases[:, 10].astype(np.float32))
    for row, got in zip(cases, out):
        want = restated.inner(o, row[:9].astype(np.float32), int(row[9]), float(np.float32(row[10])))
        if int(row[9]) < 0:
            assert np.float32(got["value"]) == np.float32(want["value"])
            if want["value"] < row[10]:
                assert np.array_equal(got["node"], want["node"])
        else:
            assert got["value"] == pytest.approx(want["value"], rel=1e-5, abs=1e-6)
        assert _close_counts(got["pops"], want["pops"])
    g.close()
    gold = runs["bunny_s0.1_mse1e-3_trim0.1"]
    g = pkg.GoICP(1e-3)
    g.pModel, g.pData = bunny["model"], bunny["data"]
    g.trimFraction = 0.1
    g.BuildDT()
    g.Register()
    _check_run(g.result, gold)
    g.close()


@pytest.mark.parametrize("sort", ["radix", "count"])
def test_config2_full_size_nn_and_icp_on_gpu(pkg, sort, monkeypatch):
    """BASELINE config 2 (bunny ICP only) at full size: 40256 nearest-neighbour indices bit-exact with the
    reference kd-tree (incl. its 53 exact-tie queries), and ICP3D::Run reproduced bit for bit -- with the
    radix sort this size selects by itself and with the counting sort of the small clouds forced."""
    import os
    monkeypatch.setenv("GOICP_ICP_RADIX", "1" if sort == "radix" else "0")
    from conftest import GOLDEN
    gold = dict(np.load(os.path.join(GOLDEN, "bun_icp_config2.npz")))
    g = pkg.GoICP(1e-5)
    g.pModel, g.pData = gold["model"], gold["data"]
    idx, d2 = g.NN(gold["data"])
    assert np.array_equal(idx, gold["nn_idx"])
    assert np.array_equal(d2.view(np.uint32), gold["nn_d2"].view(np.uint32))
    err, R, t, iters = g.ICP(np.eye(3), np.zeros(3), 10000, 1e-9)
    assert np.float32(err) == gold["icp_err"]
    assert np.array_equal(R, gold["icp_R"]) and np.array_equal(t, gold["icp_t"])
    g.close()


@pytest.mark.parametrize("name", ["spanner_s0.02_mse1e-3", "spanner_s0.02_mse1e-3_trim0.1"])
def test_spanner_noisy_scan_with_and_without_trimming(pkg, runs, name):
    """BASELINE config 4 substitute: Artec spanner scans (binary PLY, noisy target), fgoicp's translation
    domain [-1,1]^3, with and without trimming; the reference certifies right after the first ICP."""
    from conftest import load_cloud
    gold = runs[name]
    g = pkg.GoICP(gold["mse"])
    g.pModel, g.pData = load_cloud(gold["model"]), load_cloud(gold["data"])
    g.trimFraction = gold["trim"]
    g.initNodeTrans = gold["trans_cube"]
    g.BuildDT()
    g.Register()
    _check_run(g.result, gold)
    g.close()


@pytest.mark.parametrize("name", ["skull_s0.03_mse1e-3", "face_s0.025_mse1e-3"])
def test_artec_skull_and_face_scans(pkg, runs, name):
    """BASELINE config 3 substitute: Artec skull / face scans (binary PLY, CRLF headers, RGB payload) against a
    rigidly moved copy of themselves (150 / 110 degrees: ICP alone does not solve it), fgoicp's translation
    domain [-1,1]^3.  Same pose, SSE, exit path and search counters as the reference; the pose is also the
    ground-truth motion the target was generated with."""
    import os
    from conftest import GOLDEN, load_cloud
    gold = runs[name]
    g = pkg.GoICP(gold["mse"])
    g.pModel, g.pData = load_cloud(gold["model"]), load_cloud(gold["data"])
    g.initNodeTrans = gold["trans_cube"]
    g.BuildDT()
    g.Register()
    _check_run(g.result, gold)
    gt = np.load(os.path.join(GOLDEN, name.split("_")[0] + "_pose_gt.npz"))
    assert rot_angle(g.result["R"], gt["R"]) < 2e-2 and np.abs(g.result["t"] - gt["t"]).max() < 2e-2
    g.close()
