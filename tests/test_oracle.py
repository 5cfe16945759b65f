"""CPU tests of the oracle: the C restatement (oracle/goicp_oracle.c) against the golden vectors
that were produced by the UNMODIFIED reference (tests/golden/make_golden.py), and -- where
oracle/_ref has been built -- against the reference itself, bit for bit."""
import numpy as np
import pytest


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_dt_grid_vectors_distance_bit_exact_vs_golden(restated, small, bunny):
    dt = restated.dt_build(bunny["model"], 48, 2.0, keep_vectors=True)
    assert np.array_equal(restated.dt_meta(dt), small["dt48_meta"])
    assert np.array_equal(bits(restated.dt_grid(dt, 48)), bits(small["dt48_grid"]))
    assert np.array_equal(restated.dt_vectors(dt, 48), small["dt48_vec"])
    assert np.array_equal(bits(restated.dt_distance(dt, small["dt48_query"])), bits(small["dt48_dist"]))
    # the reference binary's extra seed at voxel (0,0,0) (see goicp_oracle.c) and the out-of-grid branch are exercised
    assert small["dt48_grid"][0, 0, 0] == 0.0
    idx = restated.dt_index(dt, small["dt48_query"])
    assert ((idx < 0) | (idx >= 48)).any()
    restated.dt_free(dt)


def test_dt_round_is_truncation_toward_zero(restated, small):
    dt = restated.dt_wrap(small["dt48_grid"], 48, small["dt48_meta"])
    m = small["dt48_meta"]
    # (q - min)*scale = -0.7  ->  int(-0.2) = 0 (inside!), -1.7 -> int(-1.2) = -1 (outside)
    q = np.array([[m[0] - 0.7 / m[3], m[1], m[2]], [m[0] - 1.7 / m[3], m[1], m[2]]], np.float64)
    idx = restated.dt_index(dt, q.astype(np.float32))
    assert idx[0, 0] == 0 and idx[1, 0] == -1


@pytest.mark.slow
def test_dt_full_size_checksum(restated, bunny):
    """S=300 grid of the BASELINE config: FNV-1a-64 pinned by the reference."""
    dt = restated.dt_build(bunny["model"], 300, 2.0)
    assert np.array_equal(restated.dt_meta(dt), [-1.7215785086154938, -1.735278993844986, -1.7175954878330231, 87.099870706007081])
    g = restated.dt_grid(dt, 300)
    assert "%016x" % restated.fnv(g) == "2da64ee1865a968e"
    assert float(g.max()) == pytest.approx(2.4384243, abs=1e-6)
    restated.dt_free(dt)


def test_intro_select_properties(restated):
    rng = np.random.default_rng(0)
    for n, k in [(6, 2), (50, 49), (3019, 3018), (3019, 2716), (1000, 0)]:
        for kind in range(4):
            a = rng.random(n).astype(np.float32)
            if kind == 1:
                a[rng.random(n) < 0.6] = 0          # many clamped zeros: the median-of-medians fallback
            if kind == 2:
                a = np.sort(a)
            if kind == 3:
                a[:] = 0.25
            out = restated.intro_select(a, k)
            assert np.array_equal(np.sort(out), np.sort(a))          # a permutation
            assert out[:k + 1].max() <= out[k:].min() and out[k] == np.sort(a)[k]


def test_nn_indices_vs_golden_including_ties(restated, small, bunny):
    kd = restated.kd_build(bunny["model"])
    idx, d2 = restated.kd_nn(kd, bunny["data"])
    assert np.array_equal(idx, small["nn_idx"]) and np.array_equal(bits(d2), bits(small["nn_d2"]))
    kd2 = restated.kd_build(small["lat_model"])
    idx, d2 = restated.kd_nn(kd2, small["lat_query"])
    assert np.array_equal(idx, small["lat_idx"]) and np.array_equal(d2, small["lat_d2"])
    brute = np.array([np.argmin(((small["lat_model"] - q) ** 2).sum(1)) for q in small["lat_query"]])
    assert (idx != brute).sum() > 100       # the tie winner is the first VISITED, not the lowest index
    assert np.array_equal(d2, ((small["lat_model"][brute] - small["lat_query"]) ** 2).sum(1).astype(np.float32))


def test_svd_and_icp_vs_golden(restated, small, bunny):
    for H, U, W, V in zip(small["svd_H"], small["svd_U"], small["svd_W"], small["svd_V"]):
        u, w, v = restated.svd3(H)
        assert np.array_equal(bits(u), bits(U)) and np.array_equal(bits(w), bits(W)) and np.array_equal(bits(v), bits(V))
    kd = restated.kd_build(bunny["model"])
    for trim in (0.0, 0.1):
        e, R, t, iters, _ = restated.icp_run(kd, bunny["data"], np.eye(3), np.zeros(3), 10000, 1e-7, trim)
        assert np.float32(e) == small[f"icp_trim{trim}_err"]
        assert np.array_equal(R, small[f"icp_trim{trim}_R"]) and np.array_equal(t, small[f"icp_trim{trim}_t"])


def test_inner_bnb_known_answers(restated, small, bunny):
    data = bunny["data_s"][::2].copy()
    g = restated.create(bunny["model_s"], data, 1e-3, 0.0, 64)
    restated.L.go_set_dt(g, restated.dt_wrap(small["inner_grid"], 64, small["inner_meta"]))
    restated.L.go_initialize(g)
    for lvl in range(20):
        assert np.array_equal(restated.max_rot_dis(g, lvl, len(data)), small["inner_gamma"][lvl])
    for row in small["inner_cases"]:
        r = restated.inner(g, row[:9].astype(np.float32), int(row[9]), float(np.float32(row[10])))
        assert np.float32(r["value"]) == np.float32(row[11])
        assert np.array_equal(r["node"], row[12:16].astype(np.float32))
        assert (r["pops"], r["evals"]) == (int(row[16]), int(row[17]))


def test_register_small_matches_reference_when_built(restated, reference, bunny):
    """Full Go-ICP on a small case, restatement vs the unmodified reference: identical pose, error,
    node counts and bound-evaluation count."""
    data = bunny["data_s"][::4].copy()
    a = restated.create(bunny["model_s"], data, 3e-3, 0.0, 40)
    b = reference.create(bunny["model_s"], data, 3e-3, 0.0, 40)
    restated.L.go_build_dt(a)
    reference.build_dt(b)
    ra, rb = restated.register(a), reference.register(b)
    assert np.array_equal(ra["R"], rb["R"]) and np.array_equal(ra["t"], rb["t"]) and ra["sse"] == rb["sse"]
    assert (ra["rot_pops"], ra["trans_pops"]) == (rb["rot_pops"], rb["trans_pops"])
    assert ra["bound_evals"] == rb["select_calls"] - 1 - ra["icp_calls"]


def test_do_trim_false_restatement_matches_reference_when_built(restated, reference, bunny):
    """GoICP::doTrim = false (a public field, jly_goicp.h:118): no qsort in ICP3D::Run (jly_icp3d.hpp:236-239), no intro_select
    in the bound evaluation or the scores (jly_goicp.cpp:109,293,361) -- sums run in data order.  Restatement vs the
    unmodified reference: the ICP alone, then a small full registration."""
    kd = restated.kd_build(bunny["model_s"])
    icp = reference.icp_build(bunny["model_s"])
    ea, Ra, ta, _, _ = restated.icp_run(kd, bunny["data_s"], np.eye(3), np.zeros(3), 10000, 1e-7, 0.0, False)
    eb, Rb, tb = reference.icp_run(icp, bunny["data_s"], np.eye(3), np.zeros(3), 10000, 1e-7, 0.0, False)
    assert np.float32(ea) == np.float32(eb) and np.array_equal(Ra, Rb) and np.array_equal(ta, tb)
    et, Rt, tt, _, _ = restated.icp_run(kd, bunny["data_s"], np.eye(3), np.zeros(3), 10000, 1e-7, 0.0, True)
    assert not (np.array_equal(Ra, Rt) and np.float32(ea) == np.float32(et))       # the summation order is visible in the result
    data = bunny["data_s"][::4].copy()
    a = restated.create(bunny["model_s"], data, 3e-3, 0.0, 40)
    b = reference.create(bunny["model_s"], data, 3e-3, 0.0, 40)
    restated.L.go_set_do_trim(a, 0)
    reference.L.ref_goicp_set_do_trim(b, 0)
    restated.L.go_build_dt(a)
    reference.build_dt(b)
    ra, rb = restated.register(a), reference.register(b)
    assert np.array_equal(ra["R"], rb["R"]) and np.array_equal(ra["t"], rb["t"]) and ra["sse"] == rb["sse"]
    assert (ra["rot_pops"], ra["trans_pops"]) == (rb["rot_pops"], rb["trans_pops"])


def test_golden_runs_file_is_the_survey_known_answers(runs):
    r = runs["bunny_s0.1_mse1e-3"]
    assert (r["Nm"], r["Nd"], r["rot_pops"], r["trans_pops"]) == (3594, 3019, 206, 31896)
    assert r["sse"] == pytest.approx(2.29739285, abs=1e-7) and r["select_calls"] - 3 == 235552
    assert runs["bunny_s0.1_mse5e-4"]["exit_lb"] == pytest.approx(0.796166, abs=1e-6)
    assert runs["bunny_s0.1_mse1e-3_trim0.1"]["sse"] == pytest.approx(1.65017891, abs=1e-7)


def test_config2_full_size_nn_and_icp(restated):
    """BASELINE config 2 at full size (40097 x 40256 points): the survey's known answers -- sum d^2 =
    4733.98282, FNV-1a-64 of the indices dcfe927c373eaf21, 53 exact-tie queries -- and the ICP result."""
    import os
    from conftest import GOLDEN
    g = dict(np.load(os.path.join(GOLDEN, "bun_icp_config2.npz")))
    assert "%016x" % restated.fnv(g["nn_idx"].astype(np.uint32)) == "dcfe927c373eaf21"
    assert float(g["nn_d2"].astype(np.float64).sum()) == pytest.approx(4733.98282, abs=1e-4)
    assert float(g["icp_err"]) == pytest.approx(86.54953, abs=1e-4)
    kd = restated.kd_build(g["model"])
    idx, d2 = restated.kd_nn(kd, g["data"])
    assert np.array_equal(idx, g["nn_idx"]) and np.array_equal(bits(d2), bits(g["nn_d2"]))
    e, R, t, iters, _ = restated.icp_run(kd, g["data"], np.eye(3), np.zeros(3), 10000, 1e-9, 0.0)
    assert np.float32(e) == g["icp_err"] and np.array_equal(R, g["icp_R"]) and np.array_equal(t, g["icp_t"])


def test_restatement_reproduces_the_skull_golden_run(restated, runs):
    """Artec skull scan vs its moved copy (BASELINE config 3 substitute, S=300, [-1,1]^3 translation domain): the C
    restatement follows the unmodified reference's run -- same pose, SSE, pops and evaluation count (about 50 s)."""
    from conftest import load_cloud
    gold = runs["skull_s0.03_mse1e-3"]
    g = restated.create(load_cloud(gold["model"]), load_cloud(gold["data"]), gold["mse"], 0.0, 300, trans_cube=gold["trans_cube"])
    restated.L.go_build_dt(g)
    r = restated.register(g)
    assert (r["rot_pops"], r["trans_pops"], r["exit_path"]) == (gold["rot_pops"], gold["trans_pops"], gold["exit_path"])
    assert r["bound_evals"] + r["icp_calls"] == gold["bound_evals_plus_icp"]
    assert np.float32(r["sse"]) == np.float32(gold["sse"])
    assert np.allclose(r["R"].reshape(-1), gold["R"], atol=2e-7) and np.allclose(r["t"], gold["t"], atol=2e-7)
