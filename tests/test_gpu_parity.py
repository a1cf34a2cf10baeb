"""GPU parity: libsipp.so (CUDA, through the C ABI) against the CPU oracle on the same inputs.
Bit-exact for counts, terminal flags, cost fields, paths, masks and selected candidates; floating-point gains
within 1e-5 relative (north_star tolerance) — exact where the reference's arithmetic is integer valued."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import sipp
from sipp import synth
from oracle import oracle as orc

RTOL = 1e-5   # north_star: "within 1e-5 relative for floating-point utilities"


def both(desc_kw_fn):
    return sipp.Context(desc_kw_fn(sipp.make_desc)), orc.Oracle(desc_kw_fn(orc.make_desc))


def eval_planner_desc(cfg, mk, **kw):
    obst = list(cfg["obstacle_ids_rov"]) + [cfg["unknown_id"]]
    ids = [i for i in cfg["present_ids"] if i not in obst]
    return mk(cfg["W"], cfg["H"], cfg["width"], cfg["height"], cfg["start"], cfg["goal"], min(ids), max(ids), min(ids),
              obstacle_ids_mav=cfg["obstacle_ids_mav"], obstacle_ids_rov=obst, unknown_id=cfg["unknown_id"], **kw)


def inloop_desc(cfg, mk, init="max", **kw):
    ids = [i for i in cfg["present_ids"] if i not in cfg["obstacle_ids_rov"]]
    return mk(cfg["W"], cfg["H"], cfg["width"], cfg["height"], cfg["start"], cfg["goal"], min(ids), max(ids),
              max(ids) if init == "max" else min(ids), obstacle_ids_mav=cfg["obstacle_ids_mav"],
              obstacle_ids_rov=cfg["obstacle_ids_rov"], unknown_id=cfg["unknown_id"], **kw)


ALL_PARAM_SETS = [
    dict(evaluator=sipp.EVAL_RRT_STAR, non_path_voxel_gain=0.0),
    dict(evaluator=sipp.EVAL_RRT_STAR, non_path_voxel_gain=0.001),
    dict(evaluator=sipp.EVAL_RRT_STAR, do_binary_check=True),
    dict(evaluator=sipp.EVAL_RRT_STAR, do_binary_check=True, use_distance_discount=True),
    dict(evaluator=sipp.EVAL_RRT_STAR, use_distance_discount=True, non_path_voxel_gain=0.005),
    dict(evaluator=sipp.EVAL_GOAL_DISTANCE, max_gain=100.0),
    dict(evaluator=sipp.EVAL_EXPAND_REACHABLE, discount_factor=0.1),
    dict(evaluator=sipp.EVAL_EXPAND_REACHABLE, discount_factor=0.0),
    dict(evaluator=sipp.EVAL_AVERAGE_VIEW_COST),
    dict(evaluator=sipp.EVAL_NAIVE),
]
EXACT_GAIN = lambda kw: (kw["evaluator"] in (sipp.EVAL_NAIVE, sipp.EVAL_EXPAND_REACHABLE)
                         or (kw["evaluator"] == sipp.EVAL_RRT_STAR and not kw.get("use_distance_discount")
                             and kw.get("non_path_voxel_gain", 0.0) == 0.0)
                         or (kw["evaluator"] == sipp.EVAL_RRT_STAR and kw.get("do_binary_check")))


def assert_gain_parity(g, o, kw, xy, start_xy=None, half_fov=32, bv=None):
    pg = sipp.make_params(half_fov=half_fov, bounding_volume=bv, **kw)
    po = orc.make_params(half_fov=half_fov, bounding_volume=bv, **kw)
    gg, gc, gt = g.gain_batch(pg, xy, start_xy)
    og, oc, ot = o.gain_batch(po, xy, start_xy, variant=1, threads=8)
    assert np.array_equal(gc, oc), "counts differ: %s" % (np.argwhere(gc != oc)[:5],)
    assert np.array_equal(gt, ot)
    if EXACT_GAIN(kw):
        assert np.array_equal(gg, og, equal_nan=True)
    else:
        assert np.allclose(gg, og, rtol=RTOL, atol=0.0, equal_nan=True)
        fin = np.isfinite(og) & (og != 0)
        if fin.any():
            assert np.max(np.abs(gg[fin] - og[fin]) / np.abs(og[fin])) < 1e-9   # what we actually achieve


# ------------------------------------------------------------------------------------------------------------
def small_world_pair(observed_cols=(0,)):
    W = H = 8
    mk = lambda f: f(W, H, 0.5, 0.5, (0.03, 0.03), (0.4375, 0.4375), 30, 120, 30)
    g, o = both(mk)
    labels = np.array([[30 + x + 8 * y for x in range(W)] for y in range(H)], np.int32)
    obs = np.zeros((H, W), np.uint8)
    for c in observed_cols:
        obs[:, c] = 1
    mask = np.zeros((H, W), np.uint8)
    mask[3, :] = 1
    for c in (g, o):
        c.map_upload_full(labels, obs)
        c.set_path_mask(mask)
    return g, o


@pytest.mark.parametrize("kw", ALL_PARAM_SETS)
def test_gain_hand_cases(kw):
    """SURVEY appendix D positions, incl. clipped corners, off-map centres and the negative-coordinate quirks."""
    xy = np.array([[0.20, 0.21], [0.01, 0.01], [0.03, 0.20], [0.50, 0.20], [-0.03, 0.20], [-0.07, 0.20], [-0.13, 0.20],
                   [0.20, -0.07], [0.4999, 0.4999], [0.25, 0.25], [1e9, 0.1], [np.nan, 0.1], [0.1, np.inf]])
    sxy = xy[::-1].copy()
    sxy[~np.isfinite(sxy)] = 0.1
    for cols in ((0,), (0, 1), tuple(range(8))):
        g, o = small_world_pair(cols)
        ok = np.isfinite(xy).all(axis=1) & (np.abs(xy) < 1e8).all(axis=1)
        assert_gain_parity(g, o, kw, xy[ok], sxy[ok], half_fov=1)
        # non-finite / absurd poses are undefined behaviour in the reference ((int) of NaN); we define them as off-map
        p = sipp.make_params(half_fov=1, **kw)
        gg, gc, gt = g.gain_batch(p, xy[~ok], sxy[~ok])
        assert not gc.any()
        assert_gain_parity(g, o, kw, xy[ok], sxy[ok], half_fov=1, bv=(0.0, 0.5, 0.0, 0.5))
        assert_gain_parity(g, o, kw, xy[ok], sxy[ok], half_fov=2, bv=(0.07, 0.3, 0.0, 0.26))


@pytest.mark.parametrize("half_fov", [0, 1, 7, 10, 32, 40])
def test_gain_random_small_map(half_fov):
    """W, H not multiples of 16, windows larger than the map, every evaluator."""
    W, H = 100, 70
    rng = np.random.default_rng(11)
    mk = lambda f: f(W, H, W * 0.0625, H * 0.0625, (0.5, 0.5), (5.5, 3.9), 30, 120, 30)
    g, o = both(mk)
    for _ in range(5):
        x0, y0 = (int(v) for v in rng.integers(-10, 90, 2))
        patch = rng.integers(0, 121, (33, 33)).astype(np.int32)
        patch[rng.random((33, 33)) < 0.2] = 0
        g.map_update_patch(x0, y0, patch)
        o.map_update_patch(x0, y0, patch)
    cg, sg, pg, ng = g.plan()
    co, so, po, no = o.plan()
    assert (cg, sg, ng) == (co, so, no) and np.array_equal(pg, po)
    xy = rng.uniform(-0.3, W * 0.0625 + 0.3, (257, 2))
    xy[:, 1] = rng.uniform(-0.3, H * 0.0625 + 0.3, 257)
    for kw in ALL_PARAM_SETS:
        assert_gain_parity(g, o, kw, xy, xy[::-1].copy(), half_fov=half_fov)


def test_gain_and_cycle_config2_1024(golden):
    """BASELINE config 2: 4096 candidate poses on a 1024^2 label map, RRTStarEvaluator count mode (npg 0.001 and 0.0)."""
    W = H = 1024
    lab, obs, dg = synth.world(W, H, 1002, 100, sipp.make_desc, init_cost="max")
    _, _, do = synth.world(W, H, 1002, 100, orc.make_desc, init_cost="max")
    g, o = sipp.Context(dg), orc.Oracle(do)
    g.map_upload_full(lab, obs)
    o.map_upload_full(lab, obs)
    assert g.map_stats() == o.map_stats()
    rg, ro = g.plan(), o.plan()
    assert rg[0] == ro[0] and rg[1] == ro[1] and rg[3] == ro[3] and np.array_equal(rg[2], ro[2])
    assert np.array_equal(g.field_download(), o.field_download())
    sg, _, _, pg = g.map_download()
    so, _, _, po = o.map_download()
    assert np.array_equal(sg, so) and np.array_equal(pg, po)
    xy, cost = synth.make_candidates(W, H, 1002, 4096)
    for kw in ALL_PARAM_SETS:
        assert_gain_parity(g, o, kw, xy, xy[::-1].copy())
    for npg in (0.001, 0.0):
        pgp = sipp.make_params(sipp.EVAL_RRT_STAR, non_path_voxel_gain=npg)
        pop = orc.make_params(orc.EVAL_RRT_STAR, non_path_voxel_gain=npg)
        gg, gt, gv, gi, gbv = g.cycle(pgp, xy, cost)
        og, ot, ov, oi, obv = o.cycle(pop, xy, cost, variant=1, threads=8)
        assert gi == oi, "selected candidate differs"
        assert np.array_equal(gt, ot)
        assert np.allclose(gv, ov, rtol=RTOL, atol=0) and abs(gbv - obv) <= RTOL * abs(obv)
        if npg == 0.0:
            assert np.array_equal(gg, og) and np.array_equal(gv, ov) and gbv == obv


def test_config3_and_bench_workload_4096():
    """BASELINE config 3 (cost-to-go field on a 4096^2 synthetic map, partial map, init cost max) and the bench.py workload
    (65536 candidates on that map): the whole field, the canonical path, the mask, every count and the selected candidate
    against the oracle — the oracle's Dijkstra takes ~3 s at this size, so the full comparison is affordable."""
    W = H = 4096
    lab, obs, dg = synth.world(W, H, 1003, 1500, sipp.make_desc, init_cost="max")
    _, _, do = synth.world(W, H, 1003, 1500, orc.make_desc, init_cost="max")
    g, o = sipp.Context(dg), orc.Oracle(do)
    g.map_upload_full(lab, obs)
    o.map_upload_full(lab, obs)
    rg, ro = g.plan(), o.plan()
    assert rg[0] == ro[0] and rg[1] == ro[1] and rg[3] == ro[3] and np.array_equal(rg[2], ro[2])
    fg, fo = g.field_download(), o.field_download()
    assert np.array_equal(fg, fo), "field differs in %d cells" % int((fg != fo).sum())
    assert np.array_equal(g.map_download()[3], o.map_download()[3])
    # planning twice gives the same bits (the relaxation order differs from run to run, the fixed point does not)
    rg2 = g.plan()
    assert rg2[0] == rg[0] and np.array_equal(rg2[2], rg[2]) and np.array_equal(g.field_download(), fg)
    xy, cost = synth.make_candidates(W, H, 1003, 65536)
    for npg in (0.001, 0.0):
        pgp = sipp.make_params(sipp.EVAL_RRT_STAR, non_path_voxel_gain=npg)
        pop = orc.make_params(orc.EVAL_RRT_STAR, non_path_voxel_gain=npg)
        gg, gc, gt = g.gain_batch(pgp, xy)
        og, oc, ot = o.gain_batch(pop, xy, variant=1, threads=16)
        assert np.array_equal(gc, oc) and np.array_equal(gt, ot)
        gg2, gt2, gv, gi, gbv = g.cycle(pgp, xy, cost)
        og2, ot2, ov, oi, obv = o.cycle(pop, xy, cost, variant=1, threads=16)
        assert gi == oi and np.array_equal(gt2, ot2) and np.array_equal(gg2, gg)
        assert np.allclose(gv, ov, rtol=RTOL, atol=0) and abs(gbv - obv) <= RTOL * abs(obv)
        if npg == 0.0:
            assert np.array_equal(gg, og) and np.array_equal(gv, ov) and gbv == obv


def test_config4_field_properties_8192():
    """BASELINE config 4 map size (8192^2): size-independent properties instead of a second full oracle run — (1) the field is a
    fixed point: no edge can lower any cell, and every reachable cell but the start has a neighbour it is exactly derived from
    (checked on the host with numpy over the axial edges, whose weights do not depend on the diagonal admission); (2) the path
    is a chain of neighbouring cells from start to goal whose costs increase monotonically to d[goal]; (3) idempotence."""
    W = H = 8192
    lab, obs, dg = synth.world(W, H, 1004, 6000, sipp.make_desc, init_cost="max")
    g = sipp.Context(dg)
    g.map_upload_full(lab, obs)
    cost, status, path, n = g.plan()
    assert cost > 0 and n > 2 * 8000 // 2
    f = g.field_download()
    assert f[8, 8] == 0.0 and np.isfinite(cost)
    # planner costs: observed label, init cost (max label) where unobserved, obstacle (label 0) = no node
    ids = sorted(set(np.unique(lab).tolist()) - {0})
    c = np.where(obs != 0, lab, max(ids)).astype(np.float64)
    removed = (obs != 0) & (lab == 0)
    s = (W * 0.0625) / (W + 1)
    X = (W * 0.0625 - (W + 1) * s) / 2.0 + s * np.arange(W, dtype=np.float64)
    dx = np.sqrt((X[1:] - X[:-1]) ** 2 + 0.0 * 0.0)
    # (1a) horizontal edges: d[y, x+1] <= d[y, x] + dist * (c_a + c_b) / 2 and the reverse, wherever both ends are nodes
    ok = ~removed[:, 1:] & ~removed[:, :-1]
    w = dx[None, :] * (c[:, 1:] + c[:, :-1]) / 2.0
    with np.errstate(invalid="ignore"):
        assert not np.any(ok & (f[:, 1:] > f[:, :-1] + w)) and not np.any(ok & (f[:, :-1] > f[:, 1:] + w))
    del w, ok
    # (1b) vertical edges (the reference uses the x spacing for y as well)
    Y = (H * 0.0625 - (H + 1) * s) / 2.0 + s * np.arange(H, dtype=np.float64)
    dy = np.sqrt(0.0 * 0.0 + (Y[1:] - Y[:-1]) ** 2)
    ok = ~removed[1:, :] & ~removed[:-1, :]
    w = dy[:, None] * (c[1:, :] + c[:-1, :]) / 2.0
    with np.errstate(invalid="ignore"):
        assert not np.any(ok & (f[1:, :] > f[:-1, :] + w)) and not np.any(ok & (f[:-1, :] > f[1:, :] + w))
    assert np.all(np.isinf(f[removed]))
    # (2) the path: neighbouring nodes, strictly increasing cost-to-come, ends at d[goal]
    px = np.rint((path[:, 0] - X[0]) / s).astype(int)
    py = np.rint((path[:, 1] - Y[0]) / s).astype(int)
    snap = lambda p, off: int(np.floor((p - off) / s + 0.5))            # std::round of the reference (prm_star_planner.h:81), p > 0
    assert (px[0], py[0]) == (snap(0.5, X[0]), snap(0.5, Y[0])) == (8, 8)
    assert (px[-1], py[-1]) == (snap(W * 0.0625 - 0.5, X[0]), snap(H * 0.0625 - 0.5, Y[0]))
    assert np.all(np.abs(np.diff(px)) <= 1) and np.all(np.abs(np.diff(py)) <= 1)
    along = f[py, px]
    assert np.all(np.diff(along) > 0) and along[-1] == cost
    # (3) idempotence: planning again reproduces the field bit for bit
    cost2, _, path2, n2 = g.plan()
    assert cost2 == cost and n2 == n and np.array_equal(path2, path) and np.array_equal(g.field_download(), f)


def test_config4_full_cycle_8192_against_oracle():
    """BASELINE config 4 (full planning cycle, 65536 candidates on an 8192^2 map) compared cell by cell with the oracle: field,
    path, status, mask, every count, terminal flag and the selected candidate (the oracle's Dijkstra takes ~12 s at this size)."""
    W = H = 8192
    lab, obs, dg = synth.world(W, H, 1004, 6000, sipp.make_desc, init_cost="max")
    _, _, do = synth.world(W, H, 1004, 6000, orc.make_desc, init_cost="max")
    g, o = sipp.Context(dg), orc.Oracle(do)
    g.map_upload_full(lab, obs)
    o.map_upload_full(lab, obs)
    del lab, obs
    rg, ro = g.plan(), o.plan()
    assert rg[0] == ro[0] and rg[1] == ro[1] and rg[3] == ro[3] and np.array_equal(rg[2], ro[2])
    fg, fo = g.field_download(), o.field_download()
    assert np.array_equal(fg, fo), "field differs in %d cells" % int((fg != fo).sum())
    del fg, fo
    assert np.array_equal(g.map_download()[3], o.map_download()[3])
    xy, cost = synth.make_candidates(W, H, 1004, 65536)
    for npg in (0.001, 0.0):
        pgp = sipp.make_params(sipp.EVAL_RRT_STAR, non_path_voxel_gain=npg)
        pop = orc.make_params(orc.EVAL_RRT_STAR, non_path_voxel_gain=npg)
        gg, gc, gt = g.gain_batch(pgp, xy)
        og, oc, ot = o.gain_batch(pop, xy, variant=1, threads=16)
        assert np.array_equal(gc, oc) and np.array_equal(gt, ot)
        gg2, gt2, gv, gi, gbv = g.cycle(pgp, xy, cost)
        og2, ot2, ov, oi, obv = o.cycle(pop, xy, cost, variant=1, threads=16)
        assert gi == oi, "selected candidate differs"
        assert np.array_equal(gt2, ot2) and np.array_equal(gg2, gg)
        assert np.allclose(gv, ov, rtol=RTOL, atol=0) and abs(gbv - obv) <= RTOL * abs(obv)
        if npg == 0.0:
            assert np.array_equal(gg, og) and np.array_equal(gv, ov) and gbv == obv


FP_PARAM_SETS = [
    dict(evaluator=sipp.EVAL_GOAL_DISTANCE, max_gain=100.0),
    dict(evaluator=sipp.EVAL_RRT_STAR, use_distance_discount=True, non_path_voxel_gain=0.0),
    dict(evaluator=sipp.EVAL_RRT_STAR, use_distance_discount=True, non_path_voxel_gain=0.005),
    dict(evaluator=sipp.EVAL_AVERAGE_VIEW_COST),
    dict(evaluator=sipp.EVAL_EXPAND_LOW_COST),
    dict(evaluator=sipp.EVAL_EXPAND_LOW_COST, augment_unknown_with_mean=True),
    dict(evaluator=sipp.EVAL_EXPAND_REACHABLE, discount_factor=0.1),
    dict(evaluator=sipp.EVAL_NAIVE),
]


def test_selected_candidate_is_bit_exact_for_every_evaluator_config2():
    """north_star: "bit-exact ... selected candidate". Config 2 (4096 poses, 1024^2 map): the candidate sipp_cycle selects equals the
    oracle's for the floating-point evaluators too (their gains agree to ~1e-12 relative, so only an exact tie between the two best
    values could flip the winner; the winner's index is asserted, not only its value)."""
    W = H = 1024
    kwm = dict(compute_closest_observed=True)
    lab, obs, dg = synth.world(W, H, 1002, 100, sipp.make_desc, init_cost="max", **kwm)
    _, _, do = synth.world(W, H, 1002, 100, orc.make_desc, init_cost="max", **kwm)
    g, o = sipp.Context(dg), orc.Oracle(do)
    # patches (not one masked upload) so that the closest-observed plane of ExpandLowCostArea is maintained
    from sipp import episode
    rng = np.random.default_rng(5)
    for k in range(60):
        x0, y0 = (0, 0) if k == 0 else (int(rng.integers(-20, W)), int(rng.integers(-20, H)))
        rows = cols = 200 if k == 0 else 65
        patch = episode.extract_patch(lab, x0, y0, rows, cols)
        g.map_update_patch(x0, y0, patch)
        o.map_update_patch(x0, y0, patch)
    rg, ro = g.plan(), o.plan()
    assert rg[0] == ro[0] and rg[3] == ro[3]
    xy, cost = synth.make_candidates(W, H, 1002, 4096)
    for kw in FP_PARAM_SETS:
        pg, po = sipp.make_params(half_fov=32, **kw), orc.make_params(half_fov=32, **kw)
        gg, gt, gv, gi, gbv = g.cycle(pg, xy, cost, start_xy=xy[::-1].copy())
        og, ot, ov, oi, obv = o.cycle(po, xy, cost, start_xy=xy[::-1].copy(), variant=1, threads=8)
        assert gi == oi, ("selected candidate differs", kw, gi, oi, gbv, obv)
        assert np.array_equal(gt, ot)
        assert np.allclose(gv, ov, rtol=RTOL, atol=0, equal_nan=True) and abs(gbv - obv) <= RTOL * abs(obv)
        # the runner-up is not within rounding distance of the winner: the selection does not hang on the last bits of a sum
        srt = np.sort(ov[np.isfinite(ov)])
        if len(srt) > 1 and srt[-1] > 0:
            assert (srt[-1] - srt[-2]) / srt[-1] > 1e-9 or srt[-1] == srt[-2], kw


def test_config1_example_launch_tree_cycle(golden):
    """BASELINE config 1 (example.launch defaults: map_212, GoalDistanceEvaluator max_gain 100, half_fov 32, prm_star follower,
    unknown_init_cost min; launch/example.launch:33,38-39) on an RRT-like tree of 1024 nodes: one updater pass
    (RRTStarUpdater -> UpdateAll) + value + next-best selection. Gains within RTOL, terminal flags / re-scored set / selected child
    identical to the oracle's literal emulation of the reference pass."""
    from sipp import episode
    labels, cfg = golden["map_212"]
    lab = labels.astype(np.int32)
    W, H = cfg["W"], cfg["H"]
    g = sipp.Context(inloop_desc(cfg, sipp.make_desc, "min"))
    o = orc.Oracle(inloop_desc(cfg, orc.make_desc, "min"))
    obs = synth.make_observed(W, H, 1001, 40)
    ys, xs = np.nonzero(obs)
    rng = np.random.default_rng(1001)
    # the observed region arrives as patches (65 x 65 footprints + the initial square), like in the loop
    cover = np.zeros_like(obs)
    x0s = [(0, 0, 320)] + [(int(xs[j]) - 32, int(ys[j]) - 32, 65) for j in rng.integers(0, len(xs), 60)]
    for x0, y0, sz in x0s:
        patch = episode.extract_patch(lab, x0, y0, sz, sz)
        g.map_update_patch(x0, y0, patch)
        o.map_update_patch(x0, y0, patch)
    rg, ro = g.plan(), o.plan()
    assert rg[0] == ro[0] and rg[1] == ro[1] and rg[3] == ro[3] and np.array_equal(rg[2], ro[2])
    T = 1024
    parent, end, start, cost = synth.make_tree(W, H, 1001, T)
    for kw in (dict(evaluator=sipp.EVAL_GOAL_DISTANCE, max_gain=100.0), dict(evaluator=sipp.EVAL_RRT_STAR, non_path_voxel_gain=0.001)):
        pg, po = sipp.make_params(half_fov=32, **kw), orc.make_params(half_fov=32, **kw)
        gain0, value0, term0 = np.zeros(T), np.zeros(T), np.zeros(T, np.uint8)
        ug, uo = sipp.make_updater(sipp.UPD_RRT_STAR, sipp.FOLLOW_UPDATE_ALL, -1e9, 1e9, True, False, True), orc.make_updater(0, 0, -1e9, 1e9, True, False, True)
        gg, gv, gt, nres = g.update_tree(pg, ug, parent, end, gain0, cost, value0, term0, (0.5, 0.5), True, start_xy=start)
        og, ov, ot = o.update_tree(po, uo, parent, end, gain0, cost, value0, term0, (0.5, 0.5), True, start_xy=start)
        assert np.array_equal(gt, ot) and nres >= T - 1
        assert np.allclose(gg, og, rtol=RTOL, atol=0, equal_nan=True) and np.allclose(gv, ov, rtol=RTOL, atol=0, equal_nan=True)
        # next-best selection over the updated tree ([upstream] SubsequentBest): the same child of the root
        vg, bg, bvg = g.value_select(gg, cost, parent)
        vo, bo, bvo = orc.value_select(og, cost, parent)
        assert bg == bo, ("selected child differs", kw, bg, bo, bvg, bvo)
        assert abs(bvg - bvo) <= RTOL * abs(bvo)


def test_cuda_matches_committed_reference_vectors():
    """CUDA against tests/golden/ref_evaluators.npz — vectors produced by the reference's OWN evaluator / updater / cost sources
    (compiled unmodified into oracle/_ref, tests/golden/make_ref_golden.py): unknown-voxel counts, terminal flags, updater passes'
    terminal flags and trajectory cost integrals bit for bit; gains / values exact where they are integer valued, else within RTOL."""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_ref_golden as G
    vec = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_evaluators.npz"))
    g = sipp.Context(G.make_desc(sipp.make_desc))
    for x0, y0, patch in G.patches():
        g.map_update_patch(x0, y0, patch)
    g.plan()
    state, cost, reach, path = g.map_download()
    assert np.array_equal(state, vec["state"]) and np.array_equal(cost, vec["cost"]) and np.array_equal(reach, vec["reach"])
    assert np.array_equal(path, vec["path"]) and g.map_stats()[:2] == tuple(vec["mean_max"])
    xy, sxy = G.poses()
    close = lambda a, b: np.allclose(a, b, rtol=RTOL, atol=0.0, equal_nan=True)
    k = 0
    for kw in G.EVAL_SETS:
        for hf in G.HALF_FOVS:
            for bv in G.BVS:
                gg, gc, gt = g.gain_batch(sipp.make_params(half_fov=hf, bounding_volume=bv, **kw), xy, sxy)
                assert np.array_equal(gc[:, 1], vec["unk_%03d" % k]) and np.array_equal(gt, vec["term_%03d" % k]), (kw, hf, bv)
                if EXACT_GAIN(kw):
                    assert np.array_equal(gg, vec["gain_%03d" % k], equal_nan=True), (kw, hf, bv)
                else:
                    assert close(gg, vec["gain_%03d" % k]), (kw, hf, bv)
                k += 1
    par, end, start, tcost, gain0, term0, value0 = G.tree()
    k = 0
    for upd, fol in G.UPDATER_SETS:
        for kw in (G.EVAL_SETS[0], G.EVAL_SETS[3], G.EVAL_SETS[5], G.EVAL_SETS[6], G.EVAL_SETS[7]):
            for pc in (False, True):
                for flags in ((True, False, True), (False, False, True), (True, False, False)):
                    u = sipp.make_updater(upd, fol, 0.001, 4.0, *flags)
                    tg, tv, tt, _ = g.update_tree(sipp.make_params(half_fov=10, **kw), u, par, end, gain0, tcost, value0, term0, (3.1, 2.7), pc, start_xy=start)
                    first = 1 if upd == 0 else 0
                    assert np.array_equal(tt[first:], vec["tt_%03d" % k][first:]), (upd, fol, kw, pc, flags)
                    assert close(tg[first:], vec["tg_%03d" % k][first:]) and close(tv[first:], vec["tv_%03d" % k][first:]), (upd, fol, kw, pc, flags)
                    k += 1
    # update_cost with the map-integral cost: re-integrate the followed segments (one batched sipp_cost_integral), pass old + new costs
    n = len(par)
    off = np.arange(0, 2 * n + 1, 2, dtype=np.int32)
    pts = np.zeros((2 * n, 3))
    pts[0::2, :2], pts[1::2, :2] = start, end
    integral, _ = g.cost_integral(sipp.make_cost_params(False, 0.05), off, pts)
    for k, (upd, fol) in enumerate(G.UPDATER_SETS):
        rc = vec["cc_%03d" % k]
        base = np.where(np.arange(n) == 0, 0.0, tcost)
        followed = rc != base
        cost_new = base.copy()
        cost_new[followed] = integral[followed]
        assert np.array_equal(cost_new, rc), (upd, fol)            # the CUDA line integral IS the reference's, bit for bit
        u = sipp.make_updater(upd, fol, 0.001, 4.0, True, True, True)
        tg, tv, tt, _ = g.update_tree(sipp.make_params(half_fov=10, **G.EVAL_SETS[5]), u, par, end, gain0, cost_new, value0, term0, (3.1, 2.7), True,
                                      start_xy=start, cost_below=tcost)
        first = 1 if upd == 0 else 0
        assert np.array_equal(tt[first:], vec["ct_%03d" % k][first:])
        assert close(tg[first:], vec["cg_%03d" % k][first:]) and close(tv[first:], vec["cv_%03d" % k][first:]), (upd, fol)
    off, pts, pcost = G.segments()
    k = 0
    for acc in (False, True):
        for msd in (0.05, 0.3):
            c, vd = g.cost_integral(sipp.make_cost_params(acc, msd), off, pts, pcost if acc else None)
            assert np.array_equal(c, vec["ci_%d" % k], equal_nan=True) and np.array_equal(vd, vec["civ_%d" % k]), (acc, msd)
            k += 1


def _random_tree(rng, n, W, H):
    """pre-order parent array of a random RRT-like tree (children in index order) + end / start points of the segments"""
    par = np.full(n, -1, np.int32)
    kids = [[] for _ in range(n)]
    for i in range(1, n):                       # random attachment, then renumber in pre-order
        p = int(rng.integers(0, i)) if rng.random() < 0.8 else int(rng.integers(max(0, i - 5), i))
        kids[p].append(i)
    order, stack = [], [0]
    while stack:
        v = stack.pop()
        order.append(v)
        stack.extend(reversed(kids[v]))
    new = {v: k for k, v in enumerate(order)}
    for p in range(n):
        for c in kids[p]:
            par[new[c]] = new[p]
    end = np.empty((n, 2))
    end[:, 0] = rng.uniform(-0.2, W * 0.0625 + 0.2, n)
    end[:, 1] = rng.uniform(-0.2, H * 0.0625 + 0.2, n)
    start = np.where(par[:, None] >= 0, end[np.maximum(par, 0)], end)
    return par, end, start


@pytest.mark.parametrize("upd,fol", [(0, 0), (0, 1), (1, 0), (2, 1), (2, 0)])
def test_update_tree_pass_matches_preorder_emulation(upd, fol):
    """Row A13 on the device: sipp_update_tree (predicate + compaction + one batched gain launch + pre-order values) against
    the oracle's literal node-by-node emulation of the reference pass, for every updater chain the shipped cfgs use."""
    rng = np.random.default_rng(100 + 10 * upd + fol)
    W, H = 320, 240
    lab, obs, dg = synth.world(W, H, 77, 14, sipp.make_desc, init_cost="max")
    _, _, do = synth.world(W, H, 77, 14, orc.make_desc, init_cost="max")
    g, o = sipp.Context(dg), orc.Oracle(do)
    g.map_upload_full(lab, obs)
    o.map_upload_full(lab, obs)
    g.plan(); o.plan()
    n = 1500
    par, end, start = _random_tree(rng, n, W, H)
    cost = rng.uniform(0.2, 5.0, n); cost[0] = 0.0
    for kw in (dict(evaluator=sipp.EVAL_RRT_STAR, non_path_voxel_gain=0.0), dict(evaluator=sipp.EVAL_RRT_STAR, do_binary_check=True, use_distance_discount=True),
               dict(evaluator=sipp.EVAL_GOAL_DISTANCE), dict(evaluator=sipp.EVAL_EXPAND_REACHABLE, discount_factor=0.0), dict(evaluator=sipp.EVAL_AVERAGE_VIEW_COST)):
        pg, po = sipp.make_params(half_fov=32, **kw), orc.make_params(half_fov=32, **kw)
        # state before the pass: last cycle's gains / values / terminal flags (from an older, smaller map: different numbers)
        gain0 = np.where(rng.random(n) < 0.3, 0.0, rng.uniform(0.0, 50.0, n)); gain0[0] = 0.0
        term0 = (rng.random(n) < 0.15).astype(np.uint8); term0[0] = 0
        value0 = rng.uniform(0.0, 10.0, n)
        for path_changed in (False, True):
            for flags in ((True, False, True), (True, True, True), (False, False, True), (True, False, False)):
                ug = sipp.make_updater(upd, fol, 0.001, 4.0, *flags)
                uo = orc.make_updater(upd, fol, 0.001, 4.0, *flags)
                cur = (rng.uniform(1.0, W * 0.0625 - 1.0), rng.uniform(1.0, H * 0.0625 - 1.0))
                gg, gv, gt, nres = g.update_tree(pg, ug, par, end, gain0, cost, value0, term0, cur, path_changed, start_xy=start)
                og, ov, ot = o.update_tree(po, uo, par, end, gain0, cost, value0, term0, cur, path_changed, start_xy=start)
                assert np.array_equal(gt, ot)
                assert nres == int(np.sum((gg != gain0) | (og != gain0))) or nres >= int(np.sum(og != gain0))
                if EXACT_GAIN(kw):
                    assert np.array_equal(gg, og, equal_nan=True)
                else:
                    assert np.allclose(gg, og, rtol=RTOL, atol=0.0, equal_nan=True)
                assert np.allclose(gv, ov, rtol=RTOL, atol=0.0, equal_nan=True)
                # nodes the pass must not touch keep their numbers bit for bit
                same = (og == gain0) & (ov == value0)
                assert np.array_equal(gv[same], value0[same])


def test_merge_best_records_matches_sequential_scan():
    """sipp_merge_best_dev (the multi-GPU merge of the per-shard winners) against a sequential scan over the concatenated
    candidate list: strictly greater wins, ties keep the lowest global index, NaN never beats a number, empty shards never win."""
    import torch
    g = sipp.Context(sipp.make_desc(64, 64, 4.0, 4.0, (0.5, 0.5), (3.5, 3.5), 30, 120, 30))
    rng = np.random.default_rng(9)
    dev = torch.device("cuda", 0)
    for trial in range(200):
        world = int(rng.integers(1, 9))
        stride = int(rng.integers(1, 1000))
        vals = rng.choice([0.0, 1.5, 1.5, 2.25, -3.0, np.nan, np.inf], size=world)
        idx = np.array([int(rng.integers(0, stride)) if rng.random() > 0.25 else -1 for _ in range(world)], np.int64)
        recs = np.empty((world, 2), np.int64)
        recs[:, 0] = vals.view(np.int64)
        recs[:, 1] = idx
        # sequential scan (what one rank holding every candidate would do), first record kept when nothing is comparable
        best = (vals[0], idx[0] if idx[0] >= 0 else idx[0])
        have = idx[0] >= 0 and not np.isnan(vals[0])
        for r in range(1, world):
            if idx[r] < 0 or np.isnan(vals[r]):
                continue
            gi = idx[r] + r * stride
            if not have or vals[r] > best[0] or (vals[r] == best[0] and gi < best[1]):
                best, have = (vals[r], gi), True
        d = torch.from_numpy(recs).to(dev)
        out = torch.zeros(2, dtype=torch.int64, device=dev)
        g.merge_best_dev(d.data_ptr(), world, stride, out.data_ptr())
        torch.cuda.synchronize()
        o = out.cpu().numpy()
        ov = o[:1].view(np.float64)[0]
        assert int(o[1]) == int(best[1]) and (ov == best[0] or (np.isnan(ov) and np.isnan(best[0]))), (trial, recs, o, best)


def test_expand_low_cost_area_with_closest_observed_plane():
    """A8: ExpandLowCostAreaEvaluator over the closest-observed plane that the patch ingest maintains
    (setClosestObservedMaps, map_server_planexp.cpp:803-823), with and without augment_unknown_with_mean, with a bounding volume,
    and without the plane (compute_closest_observed false: every lookup returns -1)."""
    rng = np.random.default_rng(11)
    W, H = 160, 128
    for closest in (True, False):
        mk = lambda f: f(W, H, W * 0.0625, H * 0.0625, (0.5, 0.5), (W * 0.0625 - 0.5, H * 0.0625 - 0.5), 30, 120, 30,
                         compute_closest_observed=closest)
        g, o = both(mk)
        for k in range(14):
            rows, cols = (33, 33) if k % 3 else (17, 41)
            x0, y0 = int(rng.integers(-12, W - 10)), int(rng.integers(-12, H - 10))
            patch = rng.integers(0, 121, (rows, cols)).astype(np.int32)
            g.map_update_patch(x0, y0, patch)
            o.map_update_patch(x0, y0, patch)
        assert g.map_stats() == o.map_stats()
        xy = rng.uniform(-0.3, W * 0.0625 + 0.3, (300, 2))
        xy[:, 1] = rng.uniform(-0.3, H * 0.0625 + 0.3, 300)
        for half_fov in (1, 10, 32):
            for aug in (False, True):
                for bv in (None, (1.0, 7.5, 0.7, 6.1)):
                    pg = sipp.make_params(sipp.EVAL_EXPAND_LOW_COST, half_fov=half_fov, augment_unknown_with_mean=aug, bounding_volume=bv)
                    po = orc.make_params(orc.EVAL_EXPAND_LOW_COST, half_fov=half_fov, augment_unknown_with_mean=aug, bounding_volume=bv)
                    gg, gc, gt = g.gain_batch(pg, xy)
                    og, oc, ot = o.gain_batch(po, xy, variant=1, threads=4)
                    assert np.array_equal(gc[:, :2], oc[:, :2]), (closest, half_fov, aug, bv)
                    assert np.array_equal(gt, ot)
                    assert np.allclose(gg, og, rtol=RTOL, atol=0.0, equal_nan=True), (closest, half_fov, aug, bv)
        # the other evaluators still see their own views after the low-cost view was built (and vice versa)
        g.set_path_mask(np.zeros((H, W), np.uint8)); o.set_path_mask(np.zeros((H, W), np.uint8))
        assert_gain_parity(g, o, dict(evaluator=sipp.EVAL_NAIVE), xy)
        assert_gain_parity(g, o, dict(evaluator=sipp.EVAL_AVERAGE_VIEW_COST), xy, bv=(1.0, 7.5, 0.7, 6.1))


# ------------------------------------------------------------------------------------------------------------
GOLDEN_NAMES = ["map_102", "map_110", "map_111", "map_112", "map_113", "map_115", "map_116", "map_122", "map_212", "map_312"]


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_field_golden_maps(golden, name):
    """Fully known shipped maps (eval-planner mode): golden cost, bit-exact field, identical path, mask and status."""
    labels, cfg = golden[name]
    g = sipp.Context(eval_planner_desc(cfg, sipp.make_desc))
    o = orc.Oracle(eval_planner_desc(cfg, orc.make_desc))
    lab = labels.astype(np.int32)
    g.map_upload_full(lab)
    o.map_upload_full(lab)
    cg, sg, pg, ng = g.plan()
    co, so, po, no = o.plan()
    if cfg["golden_prm_cost"] < 0:
        assert cg == -1.0 and sg == sipp.PLAN_TERMINAL_INVALID and ng == 0
    else:
        assert float("%.6g" % cg) == cfg["golden_prm_cost"]
    assert cg == co and sg == so and ng == no
    assert np.array_equal(pg, po)
    fg, fo = g.field_download(), o.field_download()
    assert np.array_equal(fg, fo), "field differs in %d cells" % int((fg != fo).sum())
    assert np.array_equal(g.map_download()[3], o.map_download()[3])
    assert g.path_mask_generation() == 1 and o.path_mask_generation() == 1


@pytest.mark.parametrize("name,init", [("map_212", "max"), ("map_112", "min"), ("map_116", "max")])
def test_field_partial_map_in_loop_mode(golden, name, init):
    """In-loop mode: unknown cells cost init_cost, ROV obstacle ids only, patches arrive one by one; first observation wins
    in the planner (prm_star_planner.cpp:111) while the map server overwrites (map_server_planexp.cpp:242)."""
    labels, cfg = golden[name]
    g = sipp.Context(inloop_desc(cfg, sipp.make_desc, init))
    o = orc.Oracle(inloop_desc(cfg, orc.make_desc, init))
    rng = np.random.default_rng(3)
    lab = labels.astype(np.int32)
    for k in range(30):
        px, py = rng.uniform(0, cfg["width"]), rng.uniform(0, cfg["height"])
        rows = cols = 65
        x0, y0 = g.patch_origin(px, py, rows, cols)
        assert (x0, y0) == o.patch_origin(px, py, rows, cols)
        patch = np.zeros((rows, cols), np.int32)
        ys, xs = np.clip(np.arange(y0, y0 + rows), 0, cfg["H"] - 1), np.clip(np.arange(x0, x0 + cols), 0, cfg["W"] - 1)
        patch[:, :] = lab[np.ix_(ys, xs)]
        if k % 7 == 3:
            patch = (patch + 5) % 121      # a re-observation with different labels: planner keeps the first, server the last
        g.map_update_patch(x0, y0, patch)
        o.map_update_patch(x0, y0, patch)
        if k % 10 == 9:
            rg, ro = g.plan(), o.plan()
            assert rg[0] == ro[0] and rg[1] == ro[1] and rg[3] == ro[3] and np.array_equal(rg[2], ro[2])
            assert np.array_equal(g.field_download(), o.field_download())
    assert g.map_stats() == o.map_stats()
    for a, b in zip(g.map_download(), o.map_download()):
        assert np.array_equal(a, b)
    # goal-explored switch (map_server_planexp.cpp:276): observe the goal, init cost drops to min
    gx, gy = cfg["goal"]
    x0, y0 = g.patch_origin(gx, gy, 65, 65)
    ys, xs = np.clip(np.arange(y0, y0 + 65), 0, cfg["H"] - 1), np.clip(np.arange(x0, x0 + 65), 0, cfg["W"] - 1)
    patch = lab[np.ix_(ys, xs)].copy()
    g.map_update_patch(x0, y0, patch)
    o.map_update_patch(x0, y0, patch)
    rg, ro = g.plan(), o.plan()
    assert rg[0] == ro[0] and rg[1] == ro[1] and rg[3] == ro[3] and np.array_equal(rg[2], ro[2])
    assert np.array_equal(g.field_download(), o.field_download())
    assert g.path_mask_generation() == o.path_mask_generation() == 4


def test_field_degenerate_cases():
    W = H = 64
    mk = lambda f, s=(0.5, 0.5), t=(3.5, 3.5): f(W, H, 4.0, 4.0, s, t, 30, 120, 30)
    lab = np.full((H, W), 30, np.int32)
    # start == goal node
    g, o = sipp.Context(mk(sipp.make_desc, (1.0, 1.0), (1.01, 1.01))), orc.Oracle(mk(orc.make_desc, (1.0, 1.0), (1.01, 1.01)))
    for c in (g, o):
        c.map_upload_full(lab)
    rg, ro = g.plan(), o.plan()
    assert rg[0] == ro[0] == 0.0 and rg[3] == ro[3] == 1 and rg[1] == ro[1]
    # start on an obstacle: nothing reachable
    lab2 = lab.copy()
    lab2[0:20, 0:20] = 0
    g, o = both(mk)
    for c in (g, o):
        c.map_upload_full(lab2)
    rg, ro = g.plan(), o.plan()
    assert rg[0] == ro[0] == -1.0 and rg[1] == ro[1] == sipp.PLAN_TERMINAL_INVALID
    assert np.array_equal(g.field_download(), o.field_download())
    # a wall with a one-cell gap, zero-cost cells (label 0 is not an obstacle here)
    mk0 = lambda f: f(W, H, 4.0, 4.0, (0.5, 0.5), (3.5, 3.5), 0, 120, 0, obstacle_ids_mav=(7,), obstacle_ids_rov=(7,))
    lab3 = np.random.default_rng(1).integers(0, 3, (H, W)).astype(np.int32) * 10
    lab3[:, 30] = 7
    lab3[41, 30] = 20
    g, o = both(mk0)
    for c in (g, o):
        c.map_upload_full(lab3)
    rg, ro = g.plan(), o.plan()
    assert rg[0] == ro[0] and rg[1] == ro[1] and rg[3] == ro[3] and np.array_equal(rg[2], ro[2])
    assert np.array_equal(g.field_download(), o.field_download())
    # the same wall with strictly positive labels: a real path through the gap
    lab4 = lab3 + 10
    lab4[:, 30] = 7
    lab4[41, 30] = 20
    g, o = both(mk0)
    for c in (g, o):
        c.map_upload_full(lab4)
    rg, ro = g.plan(), o.plan()
    assert rg[0] == ro[0] > 0 and rg[1] == ro[1] and rg[3] == ro[3] > 0 and np.array_equal(rg[2], ro[2])
    assert np.array_equal(g.field_download(), o.field_download())
    assert np.array_equal(g.map_download()[3], o.map_download()[3])
    # untouched map: everything unknown at init cost
    g, o = both(mk)
    rg, ro = g.plan(), o.plan()
    assert rg[0] == ro[0] and rg[1] == ro[1] == sipp.PLAN_VALID and np.array_equal(rg[2], ro[2])
    assert np.array_equal(g.field_download(), o.field_download())
    g.update_unknown_cost(77.5)
    o.update_unknown_cost(77.5)
    rg, ro = g.plan(), o.plan()
    assert rg[0] == ro[0] and np.array_equal(g.field_download(), o.field_download())


# ------------------------------------------------------------------------------------------------------------
def test_value_select_flat_and_tree():
    rng = np.random.default_rng(9)
    W = H = 64
    g = sipp.Context(sipp.make_desc(W, H, 4.0, 4.0, (0.5, 0.5), (3.5, 3.5), 30, 120, 30))
    for n in (1, 2, 3, 33, 1000, 5000):
        gain = rng.integers(0, 50, n).astype(np.float64) * rng.choice([1.0, 0.001], n)
        cost = rng.uniform(0.2, 5.0, n)
        cost[rng.random(n) < 0.05] = 0.0
        vg, bg, bvg = g.value_select(gain, cost, None)
        vo, bo, bvo = orc.value_select(gain, cost, None)
        assert np.array_equal(vg, vo) and bg == bo and (bg < 0 or bvg == bvo)
        parent = np.zeros(n, np.int32)
        parent[0] = -1
        for i in range(1, n):
            parent[i] = rng.integers(max(0, i - 40), i)
        vg, bg, bvg = g.value_select(gain, cost, parent)
        vo, bo, bvo = orc.value_select(gain, cost, parent)
        assert np.array_equal(vg, vo), np.argwhere(vg != vo)[:5]
        assert bg == bo and (bg < 0 or bvg == bvo)
    # a chain deeper than 512 levels is refused, not silently mis-evaluated
    n = 700
    with pytest.raises(sipp.SippError) as e:
        g.value_select(np.ones(n), np.ones(n), np.arange(-1, n - 1, dtype=np.int32))
    assert e.value.code == sipp.ERR_UNSUPPORTED
    # ties: lowest index wins
    gain = np.array([0, 3, 6, 3, 6.0]); cost = np.array([0, 1, 2, 1, 2.0])
    assert g.value_select(gain, cost)[1] == 1 == orc.value_select(gain, cost)[1]


def test_empty_and_error_paths():
    W = H = 64
    g = sipp.Context(sipp.make_desc(W, H, 4.0, 4.0, (0.5, 0.5), (3.5, 3.5), 30, 120, 30))
    p = sipp.make_params(sipp.EVAL_NAIVE, half_fov=3)
    gg, gc, gt = g.gain_batch(p, np.zeros((0, 2)))
    assert gg.shape == (0,)
    with pytest.raises(sipp.SippError) as e:
        g.gain_batch(sipp.make_params(sipp.EVAL_RRT_STAR), np.zeros((1, 2)))   # no path mask yet
    assert e.value.code == sipp.ERR_NO_PATH_MASK
    with pytest.raises(sipp.SippError):
        g.gain_batch(sipp.make_params(sipp.EVAL_NAIVE, half_fov=200), np.zeros((1, 2)))
    with pytest.raises(sipp.SippError):
        sipp.Context(sipp.make_desc(100, 100, 3.3, 3.3, (0.5, 0.5), (2.5, 2.5), 30, 120, 30))   # inexact voxel size
    with pytest.raises(sipp.SippError):
        g.map_update_patch(0, 0, np.full((4, 4), 70000, np.int32))
    assert g.launch_count() > 0


def test_episode_loop_matches_oracle_trace():
    """config 5 in small: episodes of {ingest the 65x65 patch at the scout, re-plan, score, select} on 640x480-style maps; the
    trace of every cycle — path cost, plan status, path length, selected candidate bit-identical, the winner's value within RTOL
    (bit-identical too where the gain is integer valued) — matches the oracle's"""
    from sipp import episode
    W, H = 320, 240
    for seed in (2000, 2001):
        lab = synth.make_labels(W, H, seed)
        ids = sorted(set(np.unique(lab).tolist()) - {0})
        mk = lambda f: f(W, H, W * synth.MPP, H * synth.MPP, (0.5, 0.5), (W * synth.MPP - 0.5, H * synth.MPP - 0.5), min(ids), max(ids), max(ids))
        kw = dict(half_fov=32, non_path_voxel_gain=0.001 if seed % 2 else 0.0)
        tg = episode.run_episode(sipp.Context(mk(sipp.make_desc)), lab, sipp.make_params(sipp.EVAL_RRT_STAR, **kw), seed, cycles=8, T=256, init=96)
        to = episode.run_episode(orc.Oracle(mk(orc.make_desc)), lab, orc.make_params(orc.EVAL_RRT_STAR, **kw), seed, cycles=8, T=256, init=96)
        assert episode.traces_match(tg, to, RTOL), (seed, tg, to)
        # the native loop (sipp_episode_run, what ensembles are driven with) leaves exactly the Python loop's trace
        cxy, cc = episode.episode_candidates(W, H, seed, 8, 256)
        assert sipp.Context(mk(sipp.make_desc)).episode_run(lab, sipp.make_params(sipp.EVAL_RRT_STAR, **kw), cxy, cc, fov=65, init=96) == tg
        if kw['non_path_voxel_gain'] == 0.0:
            assert tg == to          # integer-valued gains: the whole trace is bit-identical


@pytest.mark.parametrize("init", ["max", "min", 75])
def test_incremental_replan_is_bit_identical_to_a_plan_from_scratch(init):
    """A sequence of patch ingests + plans: the incremental re-plan (warm start from the previous fixed point, invalidation along the
    predecessor plane where costs went up) leaves the same field, path, mask and cost as a context that plans from scratch every
    time and as the oracle. init = 75 makes labels fall on both sides of the assumed cost (lower AND higher), obstacles always raise."""
    from sipp import episode
    W, H = 384, 288
    lab = synth.make_labels(W, H, 31)
    ids = sorted(set(np.unique(lab).tolist()) - {0})
    ic = max(ids) if init == "max" else (min(ids) if init == "min" else float(init))
    mk = lambda f: f(W, H, W * synth.MPP, H * synth.MPP, (0.5, 0.5), (W * synth.MPP - 0.5, H * synth.MPP - 0.5), min(ids), max(ids), ic)
    gi, gf, o = sipp.Context(mk(sipp.make_desc)), sipp.Context(mk(sipp.make_desc)), orc.Oracle(mk(orc.make_desc))
    gf.set_incremental(False)
    rng = np.random.default_rng(77)
    used = 0
    for step in range(14):
        k = 1 if step % 5 else 3                       # sometimes several patches between two plans
        for _ in range(k):
            if step == 0:
                x0, y0, rows, cols = 0, 0, 96, 96      # the simulator's initial square
            else:
                rows = cols = int(rng.choice([1, 17, 65]))
                x0, y0 = int(rng.integers(-20, W)), int(rng.integers(-20, H))
            patch = episode.extract_patch(lab, x0, y0, rows, cols)
            for b in (gi, gf, o):
                b.map_update_patch(x0, y0, patch)
        ri, rf, ro = gi.plan(), gf.plan(), o.plan()
        info = gi.plan_info()
        used += info["incremental"]
        assert not gf.plan_info()["incremental"]
        assert ri[0] == rf[0] == ro[0] and ri[1] == rf[1] == ro[1] and ri[3] == rf[3] == ro[3], (step, ri[0], rf[0], ro[0], info)
        assert np.array_equal(ri[2], ro[2]) and np.array_equal(rf[2], ro[2])
        fi, ff = gi.field_download(), gf.field_download()
        assert np.array_equal(fi, ff), (step, info, int((fi != ff).sum()))
        assert np.array_equal(fi, o.field_download())
        assert np.array_equal(gi.map_download()[3], o.map_download()[3])
    assert used >= 10, used                            # the goal-explored switch of the init cost may force a few full plans
    # a plan right after a plan (nothing ingested) and after an init-cost change
    r2 = gi.plan()
    assert gi.plan_info() == dict(incremental=True, patch_tiles=0, relax_seed_tiles=0, invalidated_cells=0) and r2[0] == ri[0]
    for b in (gi, gf, o):
        b.update_unknown_cost(45.0)
    assert gi.plan()[0] == gf.plan()[0] == o.plan()[0] and not gi.plan_info()["incremental"]
    assert np.array_equal(gi.field_download(), o.field_download())


def test_expand_low_cost_large_labels():
    """closest-observed labels beyond 4095 (the view packs label + 1 into 15 bits next to the "counted" bit): parity with the oracle up
    to the largest representable label, an explicit refusal above it instead of aliased costs."""
    rng = np.random.default_rng(21)
    W, H = 128, 96
    mk = lambda f: f(W, H, W * 0.0625, H * 0.0625, (0.5, 0.5), (W * 0.0625 - 0.5, H * 0.0625 - 0.5), 30, 120, 30, compute_closest_observed=True)
    g, o = both(mk)
    for k in range(10):
        x0, y0 = int(rng.integers(-8, W - 10)), int(rng.integers(-8, H - 10))
        patch = rng.choice([30, 4094, 4095, 4096, 8191, 20000, 32766], size=(25, 25)).astype(np.int32)
        g.map_update_patch(x0, y0, patch)
        o.map_update_patch(x0, y0, patch)
    xy = rng.uniform(-0.2, W * 0.0625 + 0.2, (200, 2))
    xy[:, 1] = rng.uniform(-0.2, H * 0.0625 + 0.2, 200)
    for aug in (False, True):
        pg = sipp.make_params(sipp.EVAL_EXPAND_LOW_COST, half_fov=12, augment_unknown_with_mean=aug)
        po = orc.make_params(orc.EVAL_EXPAND_LOW_COST, half_fov=12, augment_unknown_with_mean=aug)
        gg, gc, gt = g.gain_batch(pg, xy)
        og, oc, ot = o.gain_batch(po, xy, variant=1)
        assert np.array_equal(gc[:, :2], oc[:, :2]) and np.allclose(gg, og, rtol=RTOL, atol=0.0, equal_nan=True)
        assert np.max(np.abs(gg - og)) <= 1e-9 * max(1.0, np.max(np.abs(og)))
    g.map_update_patch(3, 3, np.full((4, 4), 40000, np.int32))
    with pytest.raises(sipp.SippError) as e:
        g.gain_batch(sipp.make_params(sipp.EVAL_EXPAND_LOW_COST, half_fov=12), xy)
    assert e.value.code == sipp.ERR_UNSUPPORTED
    g.gain_batch(sipp.make_params(sipp.EVAL_NAIVE, half_fov=12), xy)          # the other evaluators are not affected


@pytest.mark.gpu
def test_block_resident_relaxation_is_bit_identical():
    """A11: the fixed point does not depend on the relaxation order — the block-resident kernel (SIPP_RELAX_KERNEL=1, 128 x 128 blocks in
    shared memory, many CTAs) leaves the same bits in the field as the per-tile default, on multi-block maps with ragged edges. The
    kernel choice is read once per process, so both run in processes of their own (tools/blk_debug.py)."""
    import subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "blk_debug.py"), "130", "640", "1024"], capture_output=True, text=True, timeout=600)
    text = out.stdout + out.stderr
    assert out.returncode == 0, text
    assert text.count("same_field=True") == 3 and "failed" not in text and "differs" not in text and "TIMEOUT" not in text, text
