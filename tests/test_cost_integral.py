"""CostPlanExpMapIntegral (advanced_planner_modules/src/cost_computer/cost_planexp_map_integral.cpp:21-73): the oracle against
cases derived by hand from the cited source (nothing in the reference pins this evaluator: parity unpinned), and the CUDA path
(sipp_cost_integral) against the oracle, bit for bit."""
import numpy as np
import pytest

import sipp
from sipp import synth
from oracle import oracle as orc


def _uniform_world(mk, label=60, W=64, H=48):
    d = mk(W, H, W * synth.MPP, H * synth.MPP, (0.5, 0.5), (W * synth.MPP - 0.5, H * synth.MPP - 0.5), 30, 120, 30)
    lab = np.full((H, W), label, np.int32)
    return d, lab


def test_oracle_hand_cases():
    d, lab = _uniform_world(orc.make_desc)
    o = orc.Oracle(d)
    o.map_upload_full(lab)                   # every cell observed with cost 60 -> mean cost 60
    cp = orc.make_cost_params()
    # one straight 1 m segment inside a constant-cost map: the integral is distance * cost, whatever the sampling
    c, v = o.cost_integral(cp, [0, 2], [[0.5, 0.5, 0.0], [1.5, 0.5, 0.0]])
    assert v[0] == 1 and abs(c[0] - 60.0) < 1e-9
    # fewer than two points: cost 0, computeCost returns false (:22-25)
    c, v = o.cost_integral(cp, [0, 1, 1], [[0.5, 0.5, 0.0]])
    assert list(c) == [0.0, 0.0] and list(v) == [0, 0]
    # a pair closer than 5 mm takes the else branch (:55-61): distance * weight of the END point
    c, v = o.cost_integral(cp, [0, 2], [[0.5, 0.5, 0.0], [0.503, 0.5, 0.0]])
    assert c[0] == (0.503 - 0.5) * 60.0 or abs(c[0] - 0.003 * 60.0) < 1e-12
    # accumulate adds the parent's cost (:68-70)
    c, _ = o.cost_integral(orc.make_cost_params(accumulate=True), [0, 2], [[0.5, 0.5, 0.0], [1.5, 0.5, 0.0]], parent_cost=[7.0])
    assert abs(c[0] - 67.0) < 1e-9
    # off-map samples read -1 (getCostAtLocation :757-762) and are used as they are: a segment lying completely outside the map
    c, _ = o.cost_integral(cp, [0, 2], [[-2.0, -2.0, 0.0], [-1.0, -2.0, 0.0]])
    assert abs(c[0] - (-1.0)) < 1e-9
    # unobserved cells have weight 0 -> the sample pair is charged the MEAN map cost (:50-51) — whatever the map server's running
    # mean is (its update divides int by int, map_server_planexp.cpp:246-248, so it is not the arithmetic mean)
    o2 = orc.Oracle(d)
    obs = np.zeros(lab.shape, np.uint8)
    obs[:, :32] = 1
    o2.map_upload_full(lab, obs)
    c, _ = o2.cost_integral(cp, [0, 2], [[2.5, 0.5, 0.0], [3.5, 0.5, 0.0]])      # entirely in the unobserved half (x >= 2 m)
    mean = o2.map_stats()[0]
    assert mean > 0 and abs(c[0] - 1.0 * mean) < 1e-9


@pytest.mark.gpu
def test_cost_integral_matches_oracle_bit_for_bit():
    W, H = 320, 240
    lab, obs, dg = synth.world(W, H, 21, 10, sipp.make_desc, init_cost="max")
    _, _, do = synth.world(W, H, 21, 10, orc.make_desc, init_cost="max")
    g, o = sipp.Context(dg), orc.Oracle(do)
    g.map_upload_full(lab, obs)
    o.map_upload_full(lab, obs)
    rng = np.random.default_rng(3)
    n = 3000
    npts = rng.integers(0, 7, n)             # ragged: empty, single-point and multi-point trajectories
    off = np.concatenate([[0], np.cumsum(npts)]).astype(np.int32)
    pts = np.empty((off[-1], 3))
    k = 0
    for s in range(n):
        p = np.array([rng.uniform(-1.0, W * synth.MPP + 1.0), rng.uniform(-1.0, H * synth.MPP + 1.0), 0.0])    # some leave the map
        for _ in range(npts[s]):
            pts[k] = p
            k += 1
            step = rng.choice([0.0, 0.003, 0.05, 0.2, 1.7]) * rng.uniform(0.5, 1.5)
            a = rng.uniform(0, 2 * np.pi)
            p = p + np.array([step * np.cos(a), step * np.sin(a), rng.choice([0.0, 0.1])])
    parent = rng.uniform(0, 50, n)
    for cpk in (dict(), dict(accumulate=True), dict(max_sample_distance=0.0173)):
        cg, vg = g.cost_integral(sipp.make_cost_params(**cpk), off, pts, parent)
        co, vo = o.cost_integral(orc.make_cost_params(**cpk), off, pts, parent)
        assert np.array_equal(vg, vo)
        assert np.array_equal(cg, co), np.abs(cg - co).max()
    with pytest.raises(sipp.SippError):
        g.cost_integral(sipp.make_cost_params(max_sample_distance=0.0), off, pts)
    assert g.cost_integral(sipp.make_cost_params(), np.zeros(1, np.int32), np.zeros((0, 3)))[0].shape == (0,)
