"""GPU parity tests, second file: the kernel variants and code paths the first round left uncovered (VERDICT r1 "What's
weak" 1-5): every cluster size and both inner-BnB kernels against the reference's known answers, the large-cloud variants
(rotated points / trimming keys in global memory) against the oracle, the reference-order DT at S = 512, GoICP::doTrim =
false, poll / cancel from a second thread, and a 2-rank NCCL registration (skipped with one device).
"""
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import pytest

from conftest import ROOT, rot_angle

pytestmark = pytest.mark.gpu

DT_REFERENCE = 0


def _close_counts(a, b):
    return abs(a - b) <= max(2, 0.005 * b)


def _bumpy_surface(n, seed):
    rng = np.random.default_rng(seed)
    u = rng.normal(size=(n, 3)); u /= np.linalg.norm(u, axis=1, keepdims=True)
    k = rng.normal(size=(4, 3)); ph = rng.uniform(0, 2 * np.pi, 4)
    r = 0.33 + sum(0.04 * np.sin(3 * (u @ kk) + p) for kk, p in zip(k, ph))
    return (u * r[:, None]).astype(np.float32)


@pytest.mark.parametrize("cluster,variant", [(1, "lat"), (2, "thr"), (4, "lat"), (8, "lat"), (8, "thr"), (16, "lat"), (16, "thr"), (4, "legacy"), (4, "lat-notex"), (4, "thr-notex"),
                                             (2, "q5"), (4, "q5"), (16, "q5"), (4, "q3")])
def test_inner_bnb_every_cluster_size_and_kernel_variant(pkg, small, bunny, cluster, variant, monkeypatch):
    """The 48 known-answer InnerBnB calls of the reference (tests/golden/small_vectors.npz) through clusters of 1..16 CTAs,
    the low-latency, the two-CTAs-per-SM and the dense (192 threads x 5 / 128 x 8 CTAs per SM: "q5" / "q3") shapes of the
    pipelined kernel, the unpipelined kernel, and the DT look-ups as plain loads (GOICP_DT_TEX=0) instead of texture fetches."""
    if variant.endswith("-notex"):                     # plain loads instead of the texture path for the DT look-ups
        monkeypatch.setenv("GOICP_DT_TEX", "0")
        variant = variant[:-6]
    if variant == "legacy":
        monkeypatch.setenv("GOICP_NO_PIPELINE", "1")
    else:
        monkeypatch.setenv("GOICP_BNB_VARIANT", variant)
    code = ("import importlib, sys, numpy as np\n"
            "sys.path.insert(0, %r)\n"
            "pkg = importlib.import_module('cuda-go-icp_b200')\n"
            "small = dict(np.load(%r)); ld = lambda n: np.fromfile(%r + '/' + n, np.float32).reshape(-1, 3)\n"
            "g = pkg.GoICP(1e-3); g.pModel, g.pData = ld('bunny_model_s0.033_seed1234.f32'), ld('bunny_data_s0.033_seed1235.f32')[::2].copy()\n"
            "g.dt.SIZE = 64; g.cluster_size = %d; g.SetDT(small['inner_grid'], small['inner_meta'])\n"
            "c = small['inner_cases']; out = g.InnerBnB(c[:, :9], c[:, 9].astype(np.int32), c[:, 10].astype(np.float32))\n"
            "np.save(sys.argv[1], np.array([[o['value'], *o['node'], o['pops'], o['evals']] for o in out], np.float64))\n"
            % (ROOT, os.path.join(ROOT, "tests", "golden", "small_vectors.npz"), os.path.join(ROOT, "tests", "golden"), cluster))
    # the variant switches are read once per process (static): run each combination in its own interpreter
    out_path = os.path.join(ROOT, "gpurun_out", f"_inner_{cluster}_{variant}.npy")
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    subprocess.run([sys.executable, "-c", code, out_path], check=True, env=dict(os.environ))
    got = np.load(out_path)
    os.remove(out_path)
    for row, o in zip(small["inner_cases"], got):
        if int(row[9]) < 0:
            assert np.float32(o[0]) == np.float32(row[11])
            if row[11] < row[10]:
                assert np.array_equal(o[1:5].astype(np.float32), row[12:16].astype(np.float32))
        else:
            assert o[0] == pytest.approx(row[11], rel=1e-5, abs=1e-6)
        assert _close_counts(o[5], int(row[16])) and _close_counts(o[6], int(row[17]))


@pytest.mark.parametrize("name", ["spanner_s0.02_mse3e-4", "bunny_s0.033_mse1e-3", "bunny_s0.1_mse5e-4"])
def test_kernel_shape_and_result_reuse_change_nothing(pkg, runs, name):
    """The engine picks the inner-BnB kernel shape per round from cost forecasts and measured cycles, and keeps speculative
    results across an improvement of the incumbent when the kernel's validity range (InnerResult::reuse_*) says the call
    would have run identically.  Neither may change anything: every forced shape, and reuse switched off (GOICP_NO_REUSE=1:
    every cached result dropped at an improvement, as the reference-order semantics trivially allow), must end on the same
    pose bits, SSE, exit path and committed counters as the default."""
    code = ("import importlib, json, sys, numpy as np\n"
            "sys.path.insert(0, %r)\n"
            "pkg = importlib.import_module('cuda-go-icp_b200')\n"
            "gold = json.load(open(%r))[%r]; ld = lambda n: np.fromfile(%r + '/' + n, np.float32).reshape(-1, 3)\n"
            "g = pkg.GoICP(gold['mse']); g.pModel, g.pData = ld(gold['model']), ld(gold['data']); g.trimFraction = gold['trim']\n"
            "if 'trans_cube' in gold: g.initNodeTrans = gold['trans_cube']\n"
            "g.BuildDT(); g.Register(); r = g.result\n"
            "json.dump({'R': r['R'].tobytes().hex(), 't': r['t'].tobytes().hex(), 'sse': float(r['sse']), 'exit': r['exit_path'], 'rot_pops': int(r['rot_pops']),\n"
            "           'trans_pops': int(r['trans_pops']), 'evals': int(r['bound_evals']), 'executed': int(r['bound_evals_executed'])}, open(sys.argv[1], 'w'))\n"
            % (ROOT, os.path.join(ROOT, "tests", "golden", "goicp_runs.json"), name, os.path.join(ROOT, "tests", "golden")))
    got = {}
    for tag, env in {"default": {}, "lat": {"GOICP_BNB_VARIANT": "lat"}, "thr": {"GOICP_BNB_VARIANT": "thr"}, "dense": {"GOICP_BNB_VARIANT": "q5"},
                     "no_reuse": {"GOICP_NO_REUSE": "1"}}.items():
        out_path = os.path.join(ROOT, "gpurun_out", f"_shape_{name}_{tag}.json")
        os.makedirs(os.path.dirname(out_path), exist_ok=True)
        subprocess.run([sys.executable, "-c", code, out_path], check=True, env=dict(os.environ, **env))
        got[tag] = json.load(open(out_path))
        os.remove(out_path)
    ref = {k: v for k, v in got["default"].items() if k != "executed"}
    for tag, g in got.items():
        assert {k: v for k, v in g.items() if k != "executed"} == ref, (tag, g, ref)
    assert got["default"]["executed"] <= got["no_reuse"]["executed"]
    assert _close_counts(got["default"]["rot_pops"], runs[name]["rot_pops"]) and got["default"]["exit"] == runs[name]["exit_path"]


def test_reuse_range_of_the_kernel_is_sound_and_matches_the_oracle(pkg, restated, small, bunny):
    """InnerResult::reuse_gt / reuse_poplb (goicp_inner_result): (i) the kernel's values equal the ones the restated reference
    InnerBnB records in sequential form (tests/test_reuse_criterion.py; bounds are tree sums on the device, so to 1e-5);
    (ii) the property itself, on the device: every known-answer call re-run from a smaller optError inside its range returns
    the same value (or the new optError where the value was the old one), arg-min cube, pops and evaluations."""
    data = bunny["data_s"][::2].copy()
    og = restated.create(bunny["model_s"], data, 1e-3, 0.0, 64)
    restated.L.go_set_dt(og, restated.dt_wrap(small["inner_grid"], 64, small["inner_meta"]))
    restated.L.go_initialize(og)
    g = pkg.GoICP(1e-3); g.pModel, g.pData = bunny["model_s"], data
    g.dt.SIZE = 64; g.SetDT(small["inner_grid"], small["inner_meta"])
    c = small["inner_cases"]
    R, lvl, E = c[:, :9].astype(np.float32), c[:, 9].astype(np.int32), c[:, 10].astype(np.float32)
    base = g.InnerBnB(R, lvl, E)
    thresh = np.float32(np.float32(1e-3) * np.float32(len(data)))
    close = 0
    for k, b in enumerate(base):
        restated.inner(og, R[k], int(lvl[k]), float(E[k]))
        gt, poplb, th = restated.inner_reuse(og)
        assert th == thresh
        close += b["reuse_gt"] == pytest.approx(gt, rel=1e-5, abs=1e-6) and b["reuse_poplb"] == pytest.approx(poplb, rel=1e-5, abs=1e-6)
    assert close >= 0.8 * len(base), close                          # (a last-bit difference of a bound may move a prune decision and with it the range)
    rows = []
    for k, b in enumerate(base):
        gt, e = np.float32(b["reuse_gt"]), E[k]
        if not gt < e:
            continue
        cand = [np.nextafter(gt, np.float32(np.inf)), np.nextafter(e, np.float32(0)), np.float32(e * (1 - 8e-4))] + [np.float32(gt + (e - gt) * f) for f in (0.02, 0.5, 0.97)]
        rows += [(k, e2) for e2 in cand if gt < e2 < e and not np.float32(e2 - b["reuse_poplb"]) < thresh]
    assert len(rows) > 60, len(rows)
    ks = np.array([k for k, _ in rows])
    again = g.InnerBnB(R[ks], lvl[ks], np.array([e2 for _, e2 in rows], np.float32))
    was_E = 0
    for (k, e2), a in zip(rows, again):
        b = base[k]
        want = np.float32(e2) if np.float32(b["value"]) == E[k] else np.float32(b["value"])
        was_E += np.float32(b["value"]) == E[k]
        assert (a["pops"], a["evals"]) == (b["pops"], b["evals"]), (k, e2, E[k], b, a)
        # (an upper-bound pass that ends within the contenders' band of its starting optError has its value settled by the strict
        # re-evaluation, which starts from that optError: not the kernel's business here)
        if lvl[k] >= 0 or abs(b["value"] - E[k]) > 1e-3 * E[k] or np.float32(b["value"]) == E[k]:
            assert np.float32(a["value"]) == want, (k, e2, E[k], b, a)
            if np.float32(b["value"]) < E[k]:
                assert np.array_equal(a["node"], b["node"])
    assert was_E > 5
    g.close()


@pytest.mark.parametrize("nd,trim,cluster", [(100000, 0.0, 0), (100000, 0.1, 0), (30000, 0.1, 2), (90000, 0.0, 8)])
def test_inner_bnb_large_cloud_variants_vs_oracle(pkg, restated, nd, trim, cluster):
    """Clouds whose slice of rotated points -- and, when trimming, of residual keys -- no longer fits in shared memory
    (inner_bnb_*<PTS_SMEM = false>, key slabs in global memory, clusters of 8 and 16) against the oracle's InnerBnB on
    the same inputs: upper-bound passes settle in the reference's summation order, so value and arg-min cube are exact."""
    model = _bumpy_surface(20000, 21)
    rng = np.random.default_rng(nd + int(trim * 10))
    data = (model[rng.choice(len(model), nd)] + rng.normal(scale=4e-3, size=(nd, 3))).astype(np.float32)
    S = 48
    dt = restated.dt_build(model, S)
    o = restated.create(model, data, 2e-4, trim, S)
    restated.L.go_set_dt(o, dt)
    restated.L.go_initialize(o)
    g = pkg.GoICP(2e-4)
    g.pModel, g.pData = model, data
    g.trimFraction = trim
    g.dt.SIZE = S
    g.cluster_size = cluster
    g.SetDT(restated.dt_grid(dt, S), restated.dt_meta(dt))
    a = 0.03
    Rz = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]], np.float32)
    f = nd / 1e5 * (0.65 if trim else 1.0)
    cases = [(np.eye(3, dtype=np.float32), -1, 1e10), (Rz, -1, 75.0 * f), (Rz, 3, 60.0 * f), (np.eye(3, dtype=np.float32), 4, 1e10)]
    if trim and nd >= 100000:
        cases = cases[1:]                    # the oracle's intro_select needs ~20 s per such pass on 1e5 quantised residuals
    R = np.stack([c[0].reshape(9) for c in cases]); lvl = np.array([c[1] for c in cases], np.int32); oe = np.array([c[2] for c in cases], np.float32)
    got = g.InnerBnB(R, lvl, oe)
    g.close()
    for (Rk, l, e), out in zip(cases, got):
        want = restated.inner(o, Rk, l, float(np.float32(e)))
        if l < 0:
            assert np.float32(out["value"]) == np.float32(want["value"])
            if want["value"] < e:
                assert np.array_equal(out["node"], want["node"])
        else:
            assert out["value"] == pytest.approx(want["value"], rel=2e-5, abs=1e-6)
        assert _close_counts(out["pops"], want["pops"])


def test_dt_reference_mode_at_512_vs_restatement(pkg, restated, bunny):
    """DT3D::Build's sequential propagation at S = 512 (18 warps of 30 voxels, rows that need three voxels per producer
    thread) against the restated reference, bit for bit (FNV of the 512^3 floats and a full comparison)."""
    S = 512
    g = pkg.GoICP(1e-3)
    g.pModel, g.pData = bunny["model"], bunny["data"]
    g.dt.SIZE = S
    g.dt_mode = DT_REFERENCE
    g.BuildDT()
    grid, meta = g.GetDT()
    g.close()
    dt = restated.dt_build(bunny["model"], S)
    want = restated.dt_grid(dt, S)
    restated.dt_free(dt)
    assert np.array_equal(meta, restated.dt_frame(bunny["model"], S))
    assert np.array_equal(grid.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("numerics", [0, 2])
def test_do_trim_false_icp_and_register_vs_oracle(pkg, restated, bunny, numerics):
    """GoICP::doTrim = false: ICP3D::Run accumulates in data order (no qsort, jly_icp3d.hpp:236-239) and the bounds / scores
    skip intro_select.  Strict numerics reproduce the oracle bit for bit (ICP) / to the usual tolerances (Register); the
    fast ICP stays within 1e-3 of it."""
    kd = restated.kd_build(bunny["model_s"])
    e_ref, R_ref, t_ref, _, _ = restated.icp_run(kd, bunny["data_s"], np.eye(3), np.zeros(3), 10000, 1e-7, 0.0, False)
    g = pkg.GoICP(1e-3)
    g.pModel, g.pData = bunny["model_s"], bunny["data_s"]
    g.doTrim = False
    g.numerics = numerics
    err, R, t, iters = g.ICP(np.eye(3), np.zeros(3), 10000, 1e-7)
    g.close()
    if numerics == 0:
        assert np.float32(err) == np.float32(e_ref) and np.array_equal(R, R_ref) and np.array_equal(t, t_ref)
    else:
        assert rot_angle(R, R_ref) < 1e-3 and np.abs(t - t_ref).max() < 1e-3 and err == pytest.approx(e_ref, rel=5e-3)
        return
    data = bunny["data_s"][::4].copy()
    o = restated.create(bunny["model_s"], data, 3e-3, 0.0, 40)
    restated.L.go_set_do_trim(o, 0)
    restated.L.go_build_dt(o)
    ref = restated.register(o)
    g = pkg.GoICP(3e-3)
    g.pModel, g.pData = bunny["model_s"], data
    g.doTrim = False
    g.dt.SIZE = 40
    g.dt_mode = DT_REFERENCE
    g.BuildDT()
    g.Register()
    r = g.result
    g.close()
    assert r["exit_path"] == ref["exit_path"] and r["icp_calls"] == ref["icp_calls"]
    assert r["sse"] == pytest.approx(ref["sse"], rel=1e-5, abs=1e-7)
    assert rot_angle(r["R"], ref["R"]) < 1e-4 and np.abs(r["t"] - ref["t"]).max() < 1e-4
    assert _close_counts(r["rot_pops"], ref["rot_pops"]) and _close_counts(r["trans_pops"], ref["trans_pops"])


def test_poll_and_cancel_from_a_second_thread(pkg, runs, bunny):
    """goicp_poll / goicp_cancel next to a running goicp_register (replaces the GL thread's unlocked reads of optR / optT /
    optError and the global goicp_finished, goicp_kernel.cu:82-149, jly_goicp.cpp:36,400): snapshots are consistent, the
    error only ever decreases, the counters only grow, `finished` flips at the end; a cancel ends the call with
    GOICP_ERR_CANCELLED and the best pose so far."""
    gold = runs["bunny_s0.033_mse1e-3"]            # certified exit after 2 032 rotation pops: the longest committed search
    from conftest import load_cloud
    g = pkg.GoICP(gold["mse"])
    g.pModel, g.pData = load_cloud(gold["model"]), load_cloud(gold["data"])
    g.BuildDT()
    snaps, done = [], threading.Event()

    def poller():
        while not done.is_set():
            snaps.append(g.Poll())
            time.sleep(0.001)

    th = threading.Thread(target=poller)
    th.start()
    g.Register()
    done.set(); th.join()
    final = g.Poll()
    assert final["finished"] and final["sse"] == g.result["sse"] and final["rot_pops"] == g.result["rot_pops"] == gold["rot_pops"]
    assert np.array_equal(final["R"], g.result["R"]) and np.array_equal(final["t"], g.result["t"])
    live = [s for s in snaps if s["rot_pops"] > 0 or s["sse"] > 0]
    assert len(live) >= 3
    for a, b in zip(live, live[1:]):
        assert b["sse"] <= a["sse"] and b["rot_pops"] >= a["rot_pops"] and b["trans_pops"] >= a["trans_pops"] and b["bound_evals"] >= a["bound_evals"]
    assert not live[0]["finished"]
    # cancel: a second registration on the same handle, stopped from another thread shortly after it starts
    killer = threading.Timer(0.02, g.Cancel)
    killer.start()
    with pytest.raises(pkg.GoicpError) as ei:
        g.Register()
    killer.join()
    assert ei.value.code == 6                                        # GOICP_ERR_CANCELLED
    part = g.result
    assert part["exit_path"] == "cancelled" and 0 < part["rot_pops"] < gold["rot_pops"]
    assert part["sse"] >= gold["sse"] * (1 - 1e-6) and part["sse"] < 1e9
    assert abs(np.linalg.det(part["R"].astype(np.float64)) - 1) < 1e-4
    # and the handle is still usable: the next call runs to the certificate again
    g.Register()
    assert g.result["exit_path"] == "certified" and g.result["rot_pops"] == gold["rot_pops"]
    g.close()


def test_two_rank_nccl_registration_vs_golden(pkg, runs):
    """One registration sharded over 2 GPUs with the native NCCL exchange (ncclAllGather of the result records on the
    engine stream, hand-round of contender lists): both ranks must end with the reference's run, field for field the same
    as a 1-GPU run.  Needs two devices (gpurun --gpus 2); skipped otherwise."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 CUDA devices")
    out = os.path.join(ROOT, "gpurun_out", "_two_rank.json")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29631",
           os.path.join(ROOT, "scripts", "two_rank_check.py"), out]
    subprocess.run(cmd, check=True, timeout=600)
    import json
    res = json.load(open(out))
    os.remove(out)
    shard = res.pop("_icp_shard"); gt = res.pop("_icp_shard_gt")
    # sharded ICP: both ranks end with the same pose bit for bit, and it is the single-GPU pose up to summation order
    assert shard[0]["sharded"] == shard[1]["sharded"]
    a, b = shard[0]["sharded"], shard[0]["single"]
    assert abs(a["iterations"] - b["iterations"]) <= 2 and a["err"] == pytest.approx(b["err"], rel=2e-3)
    assert rot_angle(np.array(a["R"]).reshape(3, 3), np.array(b["R"]).reshape(3, 3)) < 2e-4 and np.abs(np.array(a["t"]) - np.array(b["t"])).max() < 2e-4
    assert rot_angle(np.array(a["R"]).reshape(3, 3), np.array(gt["R"]).reshape(3, 3)) < 5e-3
    for name, per_rank in res.items():
        gold = runs[name]
        for r in per_rank:
            assert r["exit_path"] == gold["exit_path"] and (r["rot_pops"], r["trans_pops"]) == (gold["rot_pops"], gold["trans_pops"])
            assert r["sse"] == pytest.approx(gold["sse"], rel=1e-5)
            assert rot_angle(np.array(r["R"]).reshape(3, 3), np.array(gold["R"]).reshape(3, 3)) < 1e-4
            assert np.abs(np.array(r["t"]) - np.array(gold["t"])).max() < 1e-4
        assert per_rank[0]["R"] == per_rank[1]["R"] and per_rank[0]["sse"] == per_rank[1]["sse"]


def test_cta_select_matches_reference_permutation(pkg, restated):
    """intro_select by a whole thread block (csrc/strict_sum.cuh: ranked stop lists + K independent exchanges per partition
    sweep) leaves exactly the reference's permutation (oracle go_intro_select, pinned to jly_sorting.hpp in test_oracle.py):
    random, tie-heavy, constant, sorted and reverse-sorted arrays, DT-like residuals with many exact zeros, 1 .. 200 000
    elements, k at both ends / the middle / the untrimmed n - 1, block sizes 64 .. 1024, arrays in shared and in global memory."""
    g = pkg.GoICP(1e-3)
    rng = np.random.default_rng(2024)
    cases = []
    for n in (1, 2, 5, 6, 7, 63, 64, 65, 127, 128, 129, 500, 3019, 4097, 20000, 130001, 200000):
        ks = sorted({0, n // 2, max(0, int(0.9 * n) - 1), n - 1})
        kinds = [(rng.random(n), ks), (np.arange(n), ks), (np.arange(n)[::-1], ks), (np.repeat(rng.random((n + 9) // 10), 10)[:n], ks),
                 (np.maximum(rng.normal(size=n), -0.3) + 0.3, ks[2:])]          # DT-like: a third exact zeros, selected above them
        kinds.append((np.round(np.abs(rng.normal(size=n)) * 3) / 3, ks[2:]))     # a handful of distinct values: the median-of-medians fallback at every size
        if n <= 4097:                                                           # (a selection INSIDE a run of equal values costs the reference O(n^2))
            kinds += [(rng.integers(0, 4, n), ks), (np.zeros(n), ks), (np.maximum(rng.normal(size=n), 0.0), ks)]
        for a, kk in kinds:
            for k in kk:
                cases.append((np.asarray(a, np.float32), k))
    checked = 0
    for i, (a, k) in enumerate(cases):
        want = restated.intro_select(a, k)
        threads = (64, 256, 512, 1024)[i % 4]
        for in_global in ((False, True) if len(a) <= 15000 else (True,)):
            if len(a) >= 100000 and i % 3: continue                    # the largest sizes: a third of the (kind, k) pairs
            got = g.IntroSelect(a, k, threads=threads, in_global=in_global)
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (len(a), k, threads, in_global, int((got != want).sum()))
            checked += 1
    g.close()
    assert checked > 500


def test_svd3_matches_reference_bit_for_bit(pkg, small, restated):
    """Matrix::svd on the device (register-resident restatement inside the ICP kernel) against the reference's outputs for the
    32 golden matrices (scales 1e-3 .. 1e3, every eighth rank-deficient) and against the oracle on 4 000 random ones,
    incl. near-singular, diagonal and zero matrices: U, W, V identical to the last bit."""
    g = pkg.GoICP(1e-3)
    U, W, V = g.SVD3(small["svd_H"])
    assert np.array_equal(U.view(np.uint32), small["svd_U"].view(np.uint32))
    assert np.array_equal(W.view(np.uint32), small["svd_W"].view(np.uint32))
    assert np.array_equal(V.view(np.uint32), small["svd_V"].view(np.uint32))
    rng = np.random.default_rng(99)
    H = (rng.normal(size=(4000, 3, 3)) * 10 ** rng.uniform(-4, 4, (4000, 1, 1))).astype(np.float32)
    H[::7, :, 1] = H[::7, :, 0] * np.float32(0.5)                   # rank 2
    H[::11, 2, :] = 0                                               # zero row
    H[::13] = np.diag([3.0, 2.0, 1.0]).astype(np.float32) * H[::13, :1, :1]      # diagonal
    H[5] = 0
    H[6] = np.eye(3, dtype=np.float32)
    # covariance-like matrices as ICP forms them
    P = rng.normal(size=(500, 40, 3)).astype(np.float32); Q = P + 0.05 * rng.normal(size=P.shape).astype(np.float32)
    H = np.concatenate([H, np.einsum("nki,nkj->nij", P, Q).astype(np.float32)])
    U, W, V = g.SVD3(H)
    g.close()
    for k in range(len(H)):
        u, w, v = restated.svd3(H[k])
        ok = np.array_equal(U[k].view(np.uint32), u.view(np.uint32)) and np.array_equal(W[k].view(np.uint32), w.view(np.uint32)) and np.array_equal(V[k].view(np.uint32), v.view(np.uint32))
        nan_both = np.isnan(U[k]).any() and np.isnan(u).any()
        assert ok or nan_both, (k, H[k], U[k], u)


@pytest.mark.parametrize("name", ["bunny_s0.1_mse1e-3", "skull_s0.03_mse1e-3", "face_s0.025_mse1e-3"])
def test_fgoicp_style_search_reaches_the_reference_optimum(pkg, runs, name):
    """GOICP_SEARCH_FGOICP (the strategy of the reference's GPU path, src/fgoicp/fgoicp.cpp: quaternion cube, span cut-offs,
    `ub < 2 best => ICP`): no executable oracle exists for it (it needs the CUDA + Eigen build, and its distance table only
    covers [0,1]^3), so the test pins what can be pinned -- it ends in the optimum the reference's CPU Go-ICP certifies for
    the same clouds (pose within 2e-2: its error is the nearest-neighbour SSE, not the DT's), the search stays inside its
    bounded tree (<= 1 + 8 + 64 + 512 rotation nodes), it is reproducible, and its nearest-neighbour SSE is not above the
    NN error of the golden pose."""
    from conftest import load_cloud
    gold = runs[name]
    res = []
    for _ in range(2):
        g = pkg.GoICP(gold["mse"])
        g.pModel, g.pData = load_cloud(gold["model"]), load_cloud(gold["data"])
        g.search_mode = 1
        g.BuildDT()
        g.Register()
        res.append(dict(g.result))
        # nearest-neighbour SSE of the reference's pose, by one ICP iteration's error from it
        e_gold = g.ICP(np.array(gold["R"], np.float32).reshape(3, 3), np.array(gold["t"], np.float32), 1, 1e-7)[0]
        g.close()
    a, b = res
    assert np.array_equal(a["R"], b["R"]) and a["sse"] == b["sse"] and a["rot_pops"] == b["rot_pops"]
    assert a["exit_path"] in ("certified", "queue_empty", "early_sse_below_thresh")
    assert a["rot_pops"] <= 585
    assert rot_angle(a["R"], np.array(gold["R"]).reshape(3, 3)) < 2e-2 and np.abs(a["t"] - np.array(gold["t"])).max() < 2e-2
    assert a["sse"] <= e_gold * 1.02


def _write_txt(path, arr):
    with open(path, "w") as f:
        f.write(f"{len(arr)}\n")
        for q in arr:
            f.write("%.9g %.9g %.9g\n" % tuple(q))


def test_cli_and_cpp_mirror_binaries(pkg, runs, bunny, tmp_path):
    """The two host programs the Makefile builds, run as a user would: `goicp_b200_cli <config.toml>` (the reference's
    `cis5650_fgo_icp <config.toml>` without the window: TOML -> clouds -> DT -> Go-ICP -> output.toml + viz.ply) and
    `example_mirror` (the C++ mirror of class GoICP with the reference's member names, main.cpp:47-59)."""
    import json
    pkg_dir = os.path.join(ROOT, "cuda-go-icp_b200")
    _write_txt(tmp_path / "model.txt", bunny["model"]); _write_txt(tmp_path / "data.txt", bunny["data"])
    cfg = tmp_path / "bunny.toml"
    cfg.write_text('[info]\ndescription = "bunny"\n[io]\ntarget = "model.txt"\nsource = "data.txt"\noutput = "output.toml"\nvisualization = "viz.ply"\n'
                   '[params]\nmode = 3\ntrim = true\nsubsample = 1.0\nmse_threshold = 1e-3\nresize = 1.0\n')
    r = subprocess.run([os.path.join(pkg_dir, "goicp_b200_cli"), str(cfg)], capture_output=True, text=True, cwd=tmp_path, timeout=300)
    assert r.returncode == 0, r.stderr
    out = json.loads(r.stdout.strip().splitlines()[-1])
    gold = runs["bunny_s0.1_mse1e-3"]
    assert out["exit_path"] == gold["exit_path"] and (out["rot_pops"], out["trans_pops"]) == (gold["rot_pops"], gold["trans_pops"])
    assert out["sse"] == pytest.approx(gold["sse"], rel=1e-5)
    assert rot_angle(np.array(out["R"]).reshape(3, 3), np.array(gold["R"]).reshape(3, 3)) < 1e-4
    text = (tmp_path / "output.toml").read_text()
    assert "rotation_nodes = 206" in text and "exit_path = \"early_sse_below_thresh\"" in text
    ply = (tmp_path / "viz.ply").read_text().splitlines()
    assert ply[0] == "ply" and f"element vertex {len(bunny['model']) + len(bunny['data'])}" in ply[:4]
    body = ply[ply.index("end_header") + 1:]
    assert len(body) == len(bunny["model"]) + len(bunny["data"])
    # the registered source points sit on the target: last vertex = R * last data point + t
    R, t = np.array(out["R"]).reshape(3, 3), np.array(out["t"])
    last = np.array([float(x) for x in body[-1].split()[:3]])
    assert np.abs(last - (R @ bunny["data"][-1].astype(np.float64) + t)).max() < 1e-4
    # a missing key falls back to the reference's Config defaults (mode 1 = ICP only, mse 1e-5): common.cpp:11-13,56-60
    cfg2 = tmp_path / "icp_only.toml"
    cfg2.write_text('[io]\ntarget = "model.txt"\nsource = "data.txt"\n')
    r2 = subprocess.run([os.path.join(pkg_dir, "goicp_b200_cli"), str(cfg2)], capture_output=True, text=True, cwd=tmp_path, timeout=300)
    assert r2.returncode == 0, r2.stderr
    o2 = json.loads(r2.stdout.strip().splitlines()[-1])
    assert o2["rot_pops"] == 0 and o2["exit_path"] == "none" and o2["sse"] > 0
    # unreadable input: an error message and a non-zero exit code, no crash
    r3 = subprocess.run([os.path.join(pkg_dir, "goicp_b200_cli"), str(tmp_path / "nope.toml")], capture_output=True, text=True, cwd=tmp_path, timeout=60)
    assert r3.returncode != 0 and "error" in r3.stderr.lower()
    # the C++ mirror (subsamples its inputs by 0.1 with the golden seeds: a 359 x 301 point registration)
    r4 = subprocess.run([os.path.join(pkg_dir, "example_mirror"), str(tmp_path / "model.txt"), str(tmp_path / "data.txt"), "5e-3"], capture_output=True, text=True, timeout=300)
    assert r4.returncode == 0 and r4.stdout.startswith("optError "), r4.stderr


@pytest.mark.parametrize("nm,nd,S,mse,seed_pose", [(50000, 2000, 100, 3e-4, 3), (30000, 1500, 80, 3e-4, 11)])
def test_synthetic_sweep_cases_vs_oracle(pkg, restated, nm, nd, S, mse, seed_pose):
    """BASELINE config 5 (synthetic closed surface, noisy subset under a random SE(3)) at sizes the oracle finishes in
    seconds: whole registrations -- model beyond the linear-scan range (tree NN path), BnB with several ICP refinements --
    against the oracle on the same clouds, in the reference-order DT mode (grid bit-exact) and in the default mode; and the
    recovered pose is the ground-truth motion the data was generated with."""
    sys.path.insert(0, ROOT)
    from bench import synth
    model, data, R_gt, t_gt = synth(nm, nd, seed_pose=seed_pose)
    o = restated.create(model, data, mse, 0.0, S)
    restated.L.go_build_dt(o)
    ref = restated.register(o)
    for dt_mode in (DT_REFERENCE, None):
        g = pkg.GoICP(mse)
        g.pModel, g.pData = model, data
        g.dt.SIZE = S
        if dt_mode is not None:
            g.dt_mode = dt_mode
        g.BuildDT()
        if dt_mode == DT_REFERENCE:
            assert np.array_equal(restated.dt_grid(restated.L.go_get_dt(o), S), g.GetDT()[0])
        g.Register()
        r = g.result
        g.close()
        assert r["exit_path"] == ref["exit_path"] and r["icp_calls"] == ref["icp_calls"]
        assert r["sse"] == pytest.approx(ref["sse"], rel=1e-5, abs=1e-7)
        assert rot_angle(r["R"], ref["R"]) < 1e-4 and np.abs(r["t"] - ref["t"]).max() < 1e-4
        assert _close_counts(r["rot_pops"], ref["rot_pops"]) and _close_counts(r["trans_pops"], ref["trans_pops"])
    if ref["exit_path"] != "certified" or ref["rot_pops"] > 1:
        assert rot_angle(ref["R"], R_gt) < 5e-3 and np.abs(ref["t"] - t_gt).max() < 5e-3
