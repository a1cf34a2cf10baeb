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


# ---- MavSimulator::computeCost (mav_simulator/src/mav_simulator.cpp:785-821): the simulator's travel cost over the ground truth ----
def _truth_world(mk, W=96, H=64, seed=5):
    d = mk(W, H, W * synth.MPP, H * synth.MPP, (0.5, 0.5), (W * synth.MPP - 0.5, H * synth.MPP - 0.5), 30, 120, 30)
    truth = np.random.default_rng(seed).integers(0, 121, (H, W)).astype(np.int32)
    return d, truth


def test_travel_cost_oracle_hand_cases():
    d, truth = _truth_world(orc.make_desc)
    o = orc.Oracle(d)
    const = np.full_like(truth, 60)
    # constant map, 1 m flown: MapCost = 60 * 1 m; DistanceCost = 1 m; TimeCost = 1 m / 2 m/s (sums of 20 equal steps of 0.05 m)
    a, b = [[0.5, 0.5, 0.0]], [[1.5, 0.5, 0.0]]
    assert abs(o.travel_cost(const, orc.make_travel_params(orc.TRAVEL_MAP_COST), a, b)[0] - 60.0) < 1e-9
    assert abs(o.travel_cost(const, orc.make_travel_params(orc.TRAVEL_DISTANCE_COST), a, b)[0] - 1.0) < 1e-12
    assert abs(o.travel_cost(const, orc.make_travel_params(orc.TRAVEL_TIME_COST, mav_speed=2.0), a, b)[0] - 0.5) < 1e-12
    # start == goal: n_elements = ceil(0) = 0 -> cost 0 (:790)
    assert o.travel_cost(const, orc.make_travel_params(), a, a)[0] == 0.0
    # a hop shorter than one step is ONE step charged the mean of its two end cells (:789 ceil, :799)
    t2 = const.copy()
    t2[8, 9] = 100                                           # cell (x=9, y=8); (int)(x * 96 / 6.0) = 16 x
    c = o.travel_cost(t2, orc.make_travel_params(), [[0.53, 0.53, 0.0]], [[0.57, 0.53, 0.0]])[0]      # cell 8 -> cell 9 in x
    assert c == ((60 + 100) / 2.0) * np.sqrt((0.57 - 0.53) ** 2)
    # z counts in the flown distance (3-D norm, :788) but not in the pixel lookup
    c3 = o.travel_cost(const, orc.make_travel_params(orc.TRAVEL_DISTANCE_COST), [[0.5, 0.5, 0.0]], [[0.5, 0.5, 2.0]])[0]
    assert abs(c3 - 2.0) < 1e-12


@pytest.mark.gpu
def test_travel_cost_gpu_bit_identical_to_oracle():
    d, truth = _truth_world(sipp.make_desc)
    do, _ = _truth_world(orc.make_desc)
    g, o = sipp.Context(d), orc.Oracle(do)
    g.truth_upload(truth)
    rng = np.random.default_rng(8)
    n = 3000
    W, H = 96, 64
    a = np.zeros((n, 3)); b = np.zeros((n, 3))
    a[:, 0], a[:, 1] = rng.uniform(0, W * synth.MPP - 1e-6, n), rng.uniform(0, H * synth.MPP - 1e-6, n)
    b[:, 0], b[:, 1] = rng.uniform(0, W * synth.MPP - 1e-6, n), rng.uniform(0, H * synth.MPP - 1e-6, n)
    b[::5, 2] = rng.uniform(0, 1.0, len(b[::5]))             # some climbs
    b[1::9] = a[1::9]                                        # zero-length hops
    b[2::11] = a[2::11] + 1e-4                               # hops far below one step
    a[3::50, 0] = -0.2; b[4::50, 1] = H * synth.MPP + 0.3    # out-of-map samples read label 0 on both sides
    for form in (sipp.TRAVEL_MAP_COST, sipp.TRAVEL_DISTANCE_COST, sipp.TRAVEL_TIME_COST):
        for step, speed in ((0.05, 1.0), (0.013, 0.7)):
            cg = g.travel_cost(sipp.make_travel_params(form, step, speed), a, b)
            co = o.travel_cost(truth, orc.make_travel_params(form, step, speed), a, b)
            assert np.array_equal(cg, co), (form, step, int((cg != co).sum()))
    with pytest.raises(sipp.SippError):
        sipp.Context(d).travel_cost(sipp.make_travel_params(), a[:1], b[:1])          # no truth map uploaded
    with pytest.raises(sipp.SippError):
        g.travel_cost(sipp.make_travel_params(7), a[:1], b[:1])                        # unknown formulation (mav_simulator.cpp:813-815)
