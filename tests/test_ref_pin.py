"""Pins the restated oracle (oracle/sipp_oracle.cpp) to the REFERENCE'S OWN SOURCES.

oracle/_ref/libsipp_ref.so is built by `make -C oracle _ref` from the reference's unmodified evaluator / updater / cost / map-adapter
translation units (advanced_planner_modules/src/{map/map_planexp, sensor_model/*, trajectory_evaluator/{rrt_star, goal_distance,
expand_reachable_area, average_view_cost, expand_low_cost_area}_evaluator, evaluator_updater/*, cost_computer/cost_planexp_map_integral}.cpp)
against stand-in headers (oracle/ref_shim/). Two layers:
  * live: oracle vs that library on fresh random scenarios — needs the library (built here from /root/reference; it travels to the GPU
    box as a prebuilt file), skipped where it is neither present nor buildable;
  * committed vectors: tests/golden/ref_evaluators.npz (made by tests/golden/make_ref_golden.py from the same library) — always runs.
Everything is compared BIT FOR BIT: gains, unknown-voxel counts, terminal flags, updater passes (gain / value / terminal of every
node, the root included), re-integrated costs, trajectory cost integrals."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_ref_golden as G  # noqa: E402  (the scenario of the committed vectors)
from oracle import oracle as orc  # noqa: E402
from oracle import ref  # noqa: E402

needs_ref = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libsipp_ref.so missing and /root/reference not here to build it")


def same(a, b):
    return np.array_equal(a, b, equal_nan=True)


@pytest.fixture(scope="module")
def scenario():
    o = orc.Oracle(G.make_desc(orc.make_desc))
    for x0, y0, patch in G.patches():
        o.map_update_patch(x0, y0, patch)
    o.plan()
    return o


@pytest.fixture(scope="module")
def vectors():
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_evaluators.npz"))


# ---- committed vectors (always) ---------------------------------------------------------------------------------------------------
def test_scenario_planes_match_committed_vectors(scenario, vectors):
    state, cost, reach, path = scenario.map_download()
    closest, _ = scenario.closest_download()
    mean, mx, _ = scenario.map_stats()
    assert same(state, vectors["state"]) and same(cost, vectors["cost"]) and same(reach, vectors["reach"]) and same(path, vectors["path"])
    assert same(closest, vectors["closest"]) and same(np.array([mean, mx]), vectors["mean_max"])


def test_oracle_matches_committed_reference_vectors(scenario, vectors):
    o = scenario
    xy, sxy = G.poses()
    k = 0
    for kw in G.EVAL_SETS:
        for hf in G.HALF_FOVS:
            for bv in G.BVS:
                for variant in (0, 1):      # faithful (quadratic erase loop) and linear filter: same numbers
                    g, c, t = o.gain_batch(orc.make_params(half_fov=hf, bounding_volume=bv, **kw), xy, sxy, variant=variant)
                    assert same(g, vectors["gain_%03d" % k]), (kw, hf, bv)
                    assert same(c[:, 1], vectors["unk_%03d" % k]) and same(t, vectors["term_%03d" % k]), (kw, hf, bv)
                k += 1
    par, end, start, tcost, gain0, term0, value0 = G.tree()
    k = 0
    for upd, fol in G.UPDATER_SETS:
        for kw in (G.EVAL_SETS[0], G.EVAL_SETS[3], G.EVAL_SETS[5], G.EVAL_SETS[6], G.EVAL_SETS[7]):
            for pc in (False, True):
                for flags in ((True, False, True), (False, False, True), (True, False, False)):
                    u = orc.make_updater(upd, fol, 0.001, 4.0, *flags)
                    g, v, t = o.update_tree(orc.make_params(half_fov=10, **kw), u, par, end, gain0, tcost, value0, term0, (3.1, 2.7), pc, start_xy=start, variant=0)
                    first = 1 if upd == 0 else 0     # an RRTStarUpdater chain never touches the root
                    assert same(g[first:], vectors["tg_%03d" % k][first:]) and same(v[first:], vectors["tv_%03d" % k][first:]), (upd, fol, kw, pc, flags)
                    assert same(t[first:], vectors["tt_%03d" % k][first:])
                    k += 1
    off, pts, pcost = G.segments()
    k = 0
    for acc in (False, True):
        for msd in (0.05, 0.3):
            c, vd = o.cost_integral(orc.make_cost_params(acc, msd), off, pts, pcost if acc else None)
            assert same(c, vectors["ci_%d" % k]) and same(vd, vectors["civ_%d" % k]), (acc, msd)
            k += 1


def test_update_cost_pass_matches_committed_reference_vectors(scenario, vectors):
    """update_cost (UpdateAll / UpdateAllNonTerminal call computeCost before computeValue, update_all_non_terminal.cpp:20-22) with the
    map-integral cost computer: the caller re-integrates the followed segments and passes old and new costs (sipp_update_tree's
    cost / cost_below contract); the reference does both inside its pass."""
    o = scenario
    par, end, start, tcost, gain0, term0, value0 = G.tree()
    n = len(par)
    off = np.arange(0, 2 * n + 1, 2, dtype=np.int32)
    pts = np.zeros((2 * n, 3))
    pts[0::2, :2], pts[1::2, :2] = start, end
    integral, _ = o.cost_integral(orc.make_cost_params(False, 0.05), off, pts)
    p = orc.make_params(half_fov=10, **G.EVAL_SETS[5])
    for k, (upd, fol) in enumerate(G.UPDATER_SETS):
        u = orc.make_updater(upd, fol, 0.001, 4.0, True, True, True)
        rg, rv, rt, rc = vectors["cg_%03d" % k], vectors["cv_%03d" % k], vectors["ct_%03d" % k], vectors["cc_%03d" % k]
        followed = rc != np.where(np.arange(n) == 0, 0.0, tcost)      # nodes whose cost the reference re-integrated ...
        cost_new = np.where(np.arange(n) == 0, 0.0, tcost)
        cost_new[followed] = integral[followed]
        assert same(cost_new, rc), (upd, fol)                         # ... got exactly the oracle's integral
        g, v, t = o.update_tree(p, u, par, end, gain0, cost_new, value0, term0, (3.1, 2.7), True, start_xy=start, variant=0, cost_below=tcost)
        first = 1 if upd == 0 else 0
        assert same(g[first:], rg[first:]) and same(v[first:], rv[first:]) and same(t[first:], rt[first:]), (upd, fol)


# ---- live against the library ---------------------------------------------------------------------------------------------------------
@needs_ref
def test_reference_sources_registered_their_module_types():
    for name in ("MapPlanExp", "CameraModelPlanExp", "RRTStarEvaluator", "GoalDistanceEvaluator", "ExpandReachableAreaEvaluator",
                 "AverageViewCostEvaluator", "ExpandLowCostAreaEvaluator", "RRTStarUpdater", "UpdateAllNonTerminal", "CostPlanExpMapIntegral"):
        assert ref.has_module(name), name          # the static Registration objects of the reference's own .cpp files ran
    assert not ref.has_module("CombinedEvaluator")  # declares a registration it never defines (combined_evaluator.h:43): cannot be created


@needs_ref
def test_committed_vectors_are_what_the_reference_library_produces(scenario, vectors):
    r = ref.Reference(G.make_desc(orc.make_desc)).snapshot(scenario)
    xy, sxy = G.poses()
    k = 0
    for kw in G.EVAL_SETS:
        for hf in G.HALF_FOVS:
            for bv in G.BVS:
                g, u, t = r.gain_batch(orc.make_params(half_fov=hf, bounding_volume=bv, **kw), xy, sxy)
                assert same(g, vectors["gain_%03d" % k]) and same(u, vectors["unk_%03d" % k]) and same(t, vectors["term_%03d" % k])
                k += 1


@needs_ref
def test_voxel_area_centre_twice_and_clipping():
    """MapPlanExp::getVoxelArea as compiled from the reference: the centre voxel first (centre coordinates), then the window x-outer /
    y-inner with corner coordinates, off-map cells skipped, nothing for an off-map centre."""
    d = orc.make_desc(8, 8, 0.5, 0.5, (0.03, 0.03), (0.4375, 0.4375), 30, 120, 30)
    r = ref.Reference(d)
    v = r.visible_voxels(1, 0.20, 0.21)
    assert len(v) == 10 and tuple(v[0]) == (3.5 * 0.0625, 3.5 * 0.0625, 0.0)
    assert [tuple(p[:2] / 0.0625) for p in v[1:]] == [(x, y) for x in (2, 3, 4) for y in (2, 3, 4)]
    assert len(r.visible_voxels(1, 0.01, 0.01)) == 5            # x = -1 and y = -1 clipped
    assert len(r.visible_voxels(1, 0.50, 0.20)) == 0            # x == width: centre off the map
    assert len(r.visible_voxels(1, -0.03, 0.20)) == 7           # (int) truncation toward zero: (-0.0625, 0) maps to cell 0


@needs_ref
@pytest.mark.parametrize("seed", [11, 12])
def test_oracle_matches_reference_on_random_scenarios(seed):
    rng = np.random.default_rng(seed)
    W, H = int(rng.integers(60, 200)), int(rng.integers(60, 160))
    mk = lambda f: f(W, H, W * 0.0625, H * 0.0625, (0.5, 0.5), (W * 0.0625 - 0.5, H * 0.0625 - 0.5), 30, 120, 30 if seed % 2 else 120,
                     compute_closest_observed=True)
    o = orc.Oracle(mk(orc.make_desc))
    for _ in range(10):
        rows, cols = int(rng.integers(5, 66)), int(rng.integers(5, 66))
        patch = rng.integers(0, 121, (rows, cols)).astype(np.int32)
        patch[rng.random((rows, cols)) < 0.15] = 0
        o.map_update_patch(int(rng.integers(-20, W)), int(rng.integers(-20, H)), patch)
    o.plan()
    r = ref.Reference(mk(orc.make_desc)).snapshot(o)
    xy = rng.uniform(-0.4, W * 0.0625 + 0.4, (200, 2))
    xy[:, 1] = rng.uniform(-0.4, H * 0.0625 + 0.4, 200)
    xy[:8] = [[0.0, 0.0], [W * 0.0625, 1.0], [1.0, H * 0.0625], [-0.0624, 0.5], [0.5, -0.0624], [W * 0.0625 - 1e-9, H * 0.0625 - 1e-9], [-0.0626, 0.5], [0.03125, 0.03125]]
    sxy = xy[::-1].copy()
    for kw in G.EVAL_SETS:
        for hf in (0, 3, 32):
            for bv in (None, (0.5, W * 0.03, 0.3, H * 0.05)):
                p = orc.make_params(half_fov=hf, bounding_volume=bv, **kw)
                og, oc, ot = o.gain_batch(p, xy, sxy, variant=0)
                rg, ru, rt = r.gain_batch(p, xy, sxy)
                assert same(og, rg), (kw, hf, bv, np.argwhere(~((og == rg) | (np.isnan(og) & np.isnan(rg)))).ravel()[:5])
                assert same(oc[:, 1], ru) and same(ot, rt), (kw, hf, bv)


# ---- segments viewed from several sampled viewpoints (sensor_model/sampling_time > 0) ----------------------------------------------------
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_ref_views_golden as GV  # noqa: E402


def views_of(traj, times, sampling_time):
    """what the host side of the product does: the reference's sampleViewpoints rule picks trajectory points, their positions are the views"""
    idx = orc.sample_viewpoints(times, sampling_time)
    off = np.arange(0, (len(traj) + 1) * len(idx), len(idx), dtype=np.int32)
    return off, np.ascontiguousarray(traj[:, idx, :]).reshape(-1, 2), idx


def test_sample_viewpoints_rule():
    t = (np.arange(9) * 0.4e9).astype(np.int64)
    assert orc.sample_viewpoints(t, 0.0).tolist() == [8] and orc.sample_viewpoints(t, -1.0).tolist() == [8]
    assert orc.sample_viewpoints(t, 10.0).tolist() == [8] and orc.sample_viewpoints(t, 3.2).tolist() == [8]       # interval >= the segment's duration
    assert orc.sample_viewpoints(t, 0.5).tolist() == [2, 3, 4, 5, 7, 8]          # first point at or after 0.5, 1.0, 1.5, ... (one per point at most)
    assert orc.sample_viewpoints(t, 1.0).tolist() == [3, 5, 8]
    assert orc.sample_viewpoints(t, 0.1).tolist() == [1, 2, 3, 4, 5, 6, 7, 8]    # a denser rate than the trajectory: every point but the first
    assert orc.sample_viewpoints(np.zeros(0, np.int64), 0.5).tolist() == []


def test_multi_view_oracle_matches_committed_reference_vectors(scenario):
    vec = np.load(os.path.join(ROOT, "tests", "golden", "ref_views.npz"))
    traj, times = GV.trajectories()
    assert same(traj, vec["traj"]) and same(times, vec["times"])
    k = 0
    for st in GV.SAMPLING:
        off, vxy, idx = views_of(traj, times, st)
        for kw in G.EVAL_SETS:
            for hf in GV.HALF_FOVS:
                for bv in GV.BVS:
                    og, oc, ot = scenario.gain_batch_views(orc.make_params(half_fov=hf, bounding_volume=bv, **kw), off, vxy, traj[:, 0, :], variant=0)
                    assert same(og, vec["gain_%03d" % k]), (st, kw, hf, bv, idx)
                    assert same(oc[:, 1], vec["unk_%03d" % k]) and same(ot, vec["term_%03d" % k]), (st, kw, hf, bv)
                    k += 1


@needs_ref
def test_multi_view_vectors_are_what_the_reference_library_produces(scenario):
    vec = np.load(os.path.join(ROOT, "tests", "golden", "ref_views.npz"))
    r = ref.Reference(G.make_desc(orc.make_desc)).snapshot(scenario)
    traj, times = GV.trajectories()
    k = 0
    for st in GV.SAMPLING:
        for kw in G.EVAL_SETS:
            for hf in GV.HALF_FOVS:
                for bv in GV.BVS:
                    if k % 7 == 0:      # a sample of the sets: the generator runs all of them
                        g, u, t = r.gain_batch(orc.make_params(half_fov=hf, bounding_volume=bv, **kw), traj[:, -1, :], traj[:, 0, :], sampling_time=st,
                                               traj_xy=traj, times_ns=times)
                        assert same(g, vec["gain_%03d" % k]) and same(u, vec["unk_%03d" % k]) and same(t, vec["term_%03d" % k]), (st, kw, hf, bv)
                    k += 1
