"""CPU tests of the host logic that mirrors src/common.cpp: TOML subset, TXT / PLY loaders (ascii and
binary little-endian, CRLF headers, extra properties, trailing elements), seeded subsample, errors."""
import os
import struct

import numpy as np
import pytest

from conftest import GOLDEN, ROOT


def write_txt(path, pts):
    with open(path, "w") as f:
        f.write(f"{len(pts)}\n")
        for p in pts:
            f.write("%.6f %.6f %.6f\n" % tuple(p))


def test_txt_loader_and_seeded_subsample_match_reference_rule(pkg, tmp_path, reference):
    rng = np.random.default_rng(3)
    pts = rng.uniform(-1, 1, (5000, 3)).astype(np.float32)
    pts = np.round(pts, 6).astype(np.float32)
    p = tmp_path / "cloud.txt"
    write_txt(p, pts)
    full = pkg.load_cloud(str(p), 1.0, 1.0, 7)
    assert np.array_equal(full, np.loadtxt(p, skiprows=1, dtype=np.float32))
    for sub, resize, seed in [(0.1, 1.0, 1234), (0.33, 0.02, 99)]:
        got = pkg.load_cloud(str(p), sub, resize, seed)
        want = reference.subsample(full, sub, resize, seed)      # same rule as the golden fixtures (libstdc++ mt19937)
        assert np.array_equal(got, want) and len(got) <= int(5000 * sub)


def test_golden_fixture_is_what_the_loader_produces(pkg, tmp_path, bunny):
    # round trip: the committed fixture written as TXT with full float precision, subsample 1.0
    p = tmp_path / "m.txt"
    with open(p, "w") as f:
        f.write(f"{len(bunny['model'])}\n")
        for q in bunny["model"]:
            f.write("%.9g %.9g %.9g\n" % tuple(q))
    assert np.array_equal(pkg.load_cloud(str(p), 1.0, 1.0, 1), bunny["model"])


def test_ply_ascii_with_trailing_element_and_binary_with_rgb_and_crlf(pkg, tmp_path):
    rng = np.random.default_rng(5)
    pts = rng.normal(size=(200, 3)).astype(np.float32)
    a = tmp_path / "a.ply"       # Stanford style: ascii, extra property, trailing list element
    with open(a, "w") as f:
        f.write("ply\nformat ascii 1.0\nobj_info demo\nelement vertex 200\nproperty float x\nproperty float y\nproperty float z\nproperty float confidence\n"
                "element range_grid 2\nproperty list uchar int vertex_indices\nend_header\n")
        for q in pts:
            f.write("%.9g %.9g %.9g 0.5\n" % tuple(q))
        f.write("1 0\n0\n")
    assert np.array_equal(pkg.load_cloud(str(a), 1.0, 1.0, 1), pts)
    b = tmp_path / "b.ply"       # Artec style: binary LE, CRLF header, uchar rgb after xyz
    with open(b, "wb") as f:
        f.write(b"ply\r\nformat binary_little_endian 1.0\r\nelement vertex 200\r\nproperty float x\r\nproperty float y\r\nproperty float z\r\n"
                b"property uchar red\r\nproperty uchar green\r\nproperty uchar blue\r\nend_header\r\n")
        for q in pts:
            f.write(struct.pack("<fffBBB", *q, 1, 2, 3))
    got = pkg.load_cloud(str(b), 1.0, 2.0, 1)
    assert np.array_equal(got, (np.float32(2.0) * pts).astype(np.float32))


def test_loader_errors_are_statuses_not_exceptions_or_exits(pkg, tmp_path):
    with pytest.raises(pkg.GoicpError) as e:
        pkg.load_cloud(str(tmp_path / "missing.txt"))
    assert e.value.code == 5 and "Unable to open" in str(e.value)
    bad = tmp_path / "x.xyz"
    bad.write_text("1\n0 0 0\n")
    with pytest.raises(pkg.GoicpError) as e:
        pkg.load_cloud(str(bad))
    assert "Unsupported file extension" in str(e.value)
    trunc = tmp_path / "t.txt"
    trunc.write_text("3\n0 0 0\n1 1 1\n")
    with pytest.raises(pkg.GoicpError) as e:
        pkg.load_cloud(str(trunc))
    assert "Error reading point data" in str(e.value)


def test_run_toml_reports_bad_config_and_needs_a_gpu(pkg, tmp_path, bunny):
    import torch
    cfg = tmp_path / "bad.toml"
    cfg.write_text('[io]\nsource = "nope.txt"\n')
    with pytest.raises(pkg.GoicpError) as e:
        pkg.run_toml(str(cfg))
    assert e.value.code == 5 and "missing TOML key: io.target" in str(e.value)
    # a valid config (the reference's keys, incl. tables it never reads) reaches the engine
    write_txt(tmp_path / "m.txt", bunny["model_s"]); write_txt(tmp_path / "d.txt", bunny["data_s"])
    cfg = tmp_path / "ok.toml"
    cfg.write_text('# comment\n[info]\ndescription = "x # not a comment"\n[io]\ntarget = "m.txt"   # model\nsource = "d.txt"\noutput = ""\nvisualization = ""\n'
                   '[params]\nmode = 3\ntrim = true\nsubsample = 0.5\nmse_threshold = 1e-3\nresize = 1.0\n[params.rotation]\nxmin = -180\nsearch_depth = 12\n')
    if not torch.cuda.is_available():
        with pytest.raises(pkg.GoicpError) as e:
            pkg.run_toml(str(cfg))
        assert "no CUDA device" in str(e.value)


def _same_tree(a_nodes, b_nodes, ia=0, ib=0):
    """structural equality of two flattened kd-trees whose node numbering may differ"""
    stack = [(ia, ib)]
    while stack:
        x, y = stack.pop()
        a, b = a_nodes[x], b_nodes[y]
        leaf_a, leaf_b = a[0] < 0 and a[1] < 0, b[0] < 0 and b[1] < 0
        if leaf_a != leaf_b:
            return False
        if leaf_a:
            if a[2] != b[2] or a[3] != b[3]:
                return False
        else:
            if a[4] != b[4] or a[5] != b[5] or a[6] != b[6]:        # split axis and the bit patterns of divlow / divhigh
                return False
            stack.append((a[0], b[0])); stack.append((a[1], b[1]))
    return True


@pytest.mark.parametrize("case", ["bunny", "lattice_with_duplicates", "random_50k", "planar", "tiny"])
def test_host_kdtree_layout_is_the_reference_layout(pkg, restated, bunny, case):
    """The library's host kd-tree builder (no device involved) against the oracle's ICP3D::Build restatement, itself
    pinned to the unmodified reference: same splits (bit patterns), same leaf ranges, same permutation -- the layout
    that decides which of several equidistant model points a nearest-neighbour query returns."""
    rng = np.random.default_rng(5)
    if case == "bunny":
        m = bunny["model"]
    elif case == "lattice_with_duplicates":
        ax = np.arange(9, dtype=np.float32) * 0.125
        lat = np.stack(np.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(-1, 3)
        m = np.concatenate([lat, lat[::2], lat[::7]]).astype(np.float32)
    elif case == "random_50k":
        m = rng.uniform(-1, 1, (50000, 3)).astype(np.float32)
    elif case == "planar":
        m = np.concatenate([rng.uniform(-1, 1, (3000, 2)), np.full((3000, 1), 0.25)], 1).astype(np.float32)   # zero spread on one axis
    else:
        m = rng.uniform(-1, 1, (7, 3)).astype(np.float32)                                                     # a single leaf
    nodes, vind, bbox = pkg.kdtree_host(m)
    kd = restated.kd_build(m)
    ref_nodes, ref_vind, ref_bbox = restated.kd_export(kd, len(m))
    assert len(nodes) == len(ref_nodes)
    assert np.array_equal(vind, ref_vind)
    assert np.array_equal(bbox.view(np.uint32), ref_bbox.view(np.uint32))
    assert _same_tree(nodes, ref_nodes)


def test_bnb_shape_rule_on_the_committed_round_measurements(pkg):
    """Host logic, no device: the engine's per-round choice of inner-BnB kernel shape (engine.cu: bnb_shape_rule, exported as
    goicp_bnb_shape_rule) scored on the measurements it was fitted on -- profiles/r2x_rounds/: the GOICP_ROUND_STATS lines of six
    golden runs, each forced into every shape.  The rounds of a run are the same in every shape (same results), so a rule's cost
    is the sum of the measured kernel time of the shape it picks.  It must stay within 6 % of the per-round oracle on average,
    beat every fixed shape on average, and never be worse than 12 % above the best fixed shape of a run."""
    import ctypes as C
    import re
    L = pkg.lib()
    L.goicp_bnb_shape_rule.restype = C.c_int
    L.goicp_bnb_shape_rule.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int]
    pat = re.compile(r"\[round\] shape (\d+) forecast max (\d+) sum (\d+); \[round\] tasks (\d+) kernel ([\d.]+) ms; slowest task ([\d.]+) Mcyc "
                     r"\(pops (\d+), level (-?\d+)\); max pops (\d+);")
    cfgs = ["bunny_s0.1_mse1e-3", "bunny_s0.1_mse5e-4", "bunny_s0.033_mse1e-3", "spanner_s0.02_mse3e-4", "skull_s0.03_mse1e-3", "spanner_s0.02_mse1e-4"]
    shape_of = {1: "lat", 0: "thr", 2: "q5"}
    ratios, fixed = [], {"lat": [], "thr": [], "q5": []}
    for c in cfgs:
        rows = {}
        for v in ("lat", "thr", "q5"):
            with open(os.path.join(ROOT, "profiles", "r2x_rounds", f"r2x_rounds_{c}_{v}.txt")) as f:
                rows[v] = [[float(x) for x in m.groups()] for m in map(pat.match, f) if m]
        n = len(rows["lat"])
        assert n > 0 and len(rows["thr"]) == n and len(rows["q5"]) == n
        total, oracle, prev_max = 0.0, 0.0, 0.0
        for i in range(n):
            fmax, fsum = rows["lat"][i][1], rows["lat"][i][2]
            pick = "lat" if fsum <= 0 else shape_of[L.goicp_bnb_shape_rule(fmax, fsum, prev_max, 148 // 4)]
            total += rows[pick][i][4]
            oracle += min(rows[v][i][4] for v in rows)
            prev_max = rows["lat"][i][8]
        best_fixed = min(sum(r[4] for r in rows[v]) for v in rows)
        assert total <= 1.12 * best_fixed, (c, total, best_fixed)
        ratios.append(total / oracle)
        for v in rows:
            fixed[v].append(sum(r[4] for r in rows[v]) / oracle)
    assert np.mean(ratios) < 1.06, ratios
    assert all(np.mean(ratios) < np.mean(fixed[v]) for v in fixed), (ratios, fixed)
