"""Soundness of the rule by which the engine keeps speculative InnerBnB results across an improvement of the incumbent
(InnerResult::reuse_gt / reuse_poplb in cuda-go-icp_b200/csrc/goicp_types.h), checked on the CPU against the
restatement of the reference's own InnerBnB (jly_goicp.cpp:227-340): the oracle records the same validity range in the
reference's sequential form, and every call re-run with a smaller optError inside that range must come out identical --
value (or the new optError where the value was the old one), arg-min cube, pops and bound evaluations."""
import numpy as np
import pytest


def _setup(restated, small, bunny):
    data = bunny["data_s"][::2].copy()
    g = restated.create(bunny["model_s"], data, 1e-3, 0.0, 64)
    restated.L.go_set_dt(g, restated.dt_wrap(small["inner_grid"], 64, small["inner_meta"]))
    restated.L.go_initialize(g)
    return g


def _candidates(E, gt, rng):
    """optErrors below E: right above the range's lower end, spread over the range, right below E -- and some below the range"""
    E, gt = np.float32(E), np.float32(gt)
    inside = [np.nextafter(gt, np.float32(np.inf)), np.nextafter(E, np.float32(0))]
    inside += [np.float32(gt + (E - gt) * f) for f in (0.01, 0.25, 0.5, 0.9, 0.999)]
    inside += [np.float32(E * (1 - f)) for f in (8e-4, 1e-5)]
    below = [gt, np.nextafter(gt, np.float32(0)), np.float32(gt * 0.9), np.float32(gt * rng.uniform(0.3, 0.99))]
    return [e for e in inside if gt < e < E], [e for e in below if 0 < e <= gt]


def test_reuse_range_is_sound_on_the_known_answer_calls(restated, small, bunny):
    g = _setup(restated, small, bunny)
    rng = np.random.default_rng(5)
    checked = reused_value_is_E = differs_below = below = 0
    cases = [(row[:9].astype(np.float32), int(row[9]), np.float32(row[10])) for row in small["inner_cases"]]
    # the committed calls start from one optError each: add, for every rotation, starts spread over two decades
    cases = cases + [(R, lvl, np.float32(E * f)) for (R, lvl, E) in cases[::4] for f in (0.5, 0.05, 2.0)]
    for R, lvl, E in cases:
        a = restated.inner(g, R, lvl, float(E))
        gt, poplb, thresh = restated.inner_reuse(g)
        if a["pops"] > 500:                                      # keep the CPU suite short: the long searches are re-run a dozen times below
            continue
        inside, outside = _candidates(E, gt, rng)
        for e2 in inside:
            if np.float32(e2 - poplb) < thresh:                 # a node expanded before the first improvement would have ended the call
                continue
            b = restated.inner(g, R, lvl, float(e2))
            want = np.float32(e2) if np.float32(a["value"]) == E else np.float32(a["value"])
            reused_value_is_E += np.float32(a["value"]) == E
            assert np.float32(b["value"]) == want, (lvl, E, e2, gt, poplb, a, b)
            assert (b["pops"], b["evals"]) == (a["pops"], a["evals"]), (lvl, E, e2, gt, poplb, a, b)
            if np.float32(a["value"]) < E:
                assert np.array_equal(a["node"], b["node"])
            checked += 1
        for e2 in outside:                                      # no claim below the range -- but the range must not be vacuous
            b = restated.inner(g, R, lvl, float(e2))
            below += 1
            differs_below += (b["pops"], b["evals"]) != (a["pops"], a["evals"]) or (np.float32(b["value"]) != np.float32(a["value"]) and np.float32(a["value"]) != E)
    print(f'checked {checked}, value-was-E {reused_value_is_E}, changed below the range {differs_below} of {below}')
    assert checked > 100 and reused_value_is_E > 10
    assert differs_below >= below // 8                           # right at / below the bound the calls really do change


@pytest.mark.parametrize("level", [-1, 2, 5])
def test_reuse_range_random_rotations(restated, small, bunny, level):
    """random rotation-cube centres, upper- and lower-bound passes, starts near the values such calls return"""
    g = _setup(restated, small, bunny)
    rng = np.random.default_rng(100 + level)
    checked = 0
    for _ in range(14):
        ok, R = restated.cube_rotation(*(rng.uniform(-2.5, 2.0, 3)), 0.4)
        if not ok:
            continue
        E = np.float32(rng.choice([3.0, 10.0, 40.0, 150.0]))
        a = restated.inner(g, R, level, float(E))
        gt, poplb, thresh = restated.inner_reuse(g)
        if a["pops"] > 250:
            continue
        for e2 in _candidates(E, gt, rng)[0]:
            if np.float32(e2 - poplb) < thresh:
                continue
            b = restated.inner(g, R, level, float(e2))
            want = np.float32(e2) if np.float32(a["value"]) == E else np.float32(a["value"])
            assert np.float32(b["value"]) == want and (b["pops"], b["evals"]) == (a["pops"], a["evals"]), (level, E, e2, gt, poplb, a, b)
            checked += 1
    print(f'level {level}: checked {checked}')
    assert checked > 20
