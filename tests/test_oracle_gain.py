"""Hand-checkable footprint / gain cases derived line by line from the cited reference sources
(map_planexp.cpp:72-101, rrt_star_evaluator.cpp:36-96, average_view_cost_evaluator.cpp:31-61,
 goal_distance_evaluator.cpp:33-66). Nothing in the reference pins these ("parity unpinned"): they check that
the oracle restates the source faithfully."""
import math

import numpy as np
import pytest

from oracle import oracle as orc

VS = 0.0625


def small_world(observed_cols=(0,), **kw):
    W = H = 8
    d = orc.make_desc(W, H, 0.5, 0.5, (0.03, 0.03), (0.4375, 0.4375), 30, 120, 30, **kw)
    o = orc.Oracle(d)
    labels = np.zeros((H, W), np.int32)
    for y in range(H):
        for x in range(W):
            labels[y, x] = 30 + x + 8 * y
    obs = np.zeros((H, W), np.uint8)
    for c in observed_cols:
        obs[:, c] = 1
    o.map_upload_full(labels, obs)
    mask = np.zeros((H, W), np.uint8)
    mask[3, :] = 1
    o.set_path_mask(mask)
    return o


def seq_sum(terms):
    g = 0.0
    for t in terms:
        g += t
    return g


def test_case1_centre_counted_twice():
    o = small_world()
    p0 = orc.make_params(orc.EVAL_RRT_STAR, half_fov=1, non_path_voxel_gain=0.0)
    gain, cnt, term = o.gain_batch(p0, [[0.20, 0.21]])
    assert cnt[0].tolist() == [10, 10, 4, 0]      # 1 centre + 9 window voxels, cell (3,3) twice, 4 on the y==3 path row
    assert gain[0] == 4.0 and term[0] == 0
    p1 = orc.make_params(orc.EVAL_RRT_STAR, half_fov=1, non_path_voxel_gain=0.001)
    gain, _, _ = o.gain_batch(p1, [[0.20, 0.21]])
    # voxel order: centre (3,3)[path], then x outer / y inner: (2,2)(2,3)(2,4)(3,2)(3,3)(3,4)(4,2)(4,3)(4,4)
    assert gain[0] == seq_sum([1.0, .001, 1.0, .001, .001, 1.0, .001, .001, 1.0, .001])
    # binary mode
    pb = orc.make_params(orc.EVAL_RRT_STAR, half_fov=1, do_binary_check=True)
    assert o.gain_batch(pb, [[0.20, 0.21]])[0][0] == 1.0
    # binary + discount measures from trajectory[0] (rrt_star_evaluator.cpp:65)
    pbd = orc.make_params(orc.EVAL_RRT_STAR, half_fov=1, do_binary_check=True, use_distance_discount=True)
    g = o.gain_batch(pbd, [[0.20, 0.21]], start_xy=[[0.1, 0.1]])[0][0]
    assert g == 1.0 / math.sqrt(math.sqrt((0.4375 - 0.1) ** 2 + (0.4375 - 0.1) ** 2 + 0.0) + 0.1)
    # naive = number of unknown voxels
    assert o.gain_batch(orc.make_params(orc.EVAL_NAIVE, half_fov=1), [[0.20, 0.21]])[0][0] == 10.0


def test_case2_clipped_corner_and_average_view_cost():
    o = small_world()
    p = orc.make_params(orc.EVAL_RRT_STAR, half_fov=1, non_path_voxel_gain=0.001)
    gain, cnt, term = o.gain_batch(p, [[0.01, 0.01]])
    assert cnt[0].tolist() == [5, 2, 0, 0]
    assert gain[0] == 0.001 + 0.001 and term[0] == 0
    pa = orc.make_params(orc.EVAL_AVERAGE_VIEW_COST, half_fov=1)
    gain, cnt, _ = o.gain_batch(pa, [[0.01, 0.01]])
    assert cnt[0].tolist() == [5, 2, 3, 98]       # observed FREE: centre (0,0), (0,0), (0,1): 30 + 30 + 38
    assert gain[0] == 2 * (1.0 / (98.0 / 3))


def test_case3_goal_distance_uses_corner_coordinates():
    o = small_world()
    p = orc.make_params(orc.EVAL_RRT_STAR, half_fov=1, non_path_voxel_gain=0.001)
    gain, cnt, _ = o.gain_batch(p, [[0.03, 0.20]])
    assert cnt[0].tolist() == [7, 3, 1, 0]
    assert gain[0] == seq_sum([0.001, 1.0, 0.001])   # (1,2) (1,3)[path] (1,4)
    pg = orc.make_params(orc.EVAL_GOAL_DISTANCE, half_fov=1, max_gain=100.0)
    gain, cnt, term = o.gain_batch(pg, [[0.03, 0.20]])
    exp = seq_sum([1.0 / math.sqrt((0.4375 - 1 * VS) ** 2 + (0.4375 - y * VS) ** 2 + 0.0) for y in (2, 3, 4)])
    assert gain[0] == exp and term[0] == 0
    # discount mode: only the on-path voxel contributes the discounted term
    pd = orc.make_params(orc.EVAL_RRT_STAR, half_fov=1, non_path_voxel_gain=0.001, use_distance_discount=True)
    gain, _, _ = o.gain_batch(pd, [[0.03, 0.20]])
    disc = 1.0 / math.sqrt(math.sqrt((0.4375 - VS) ** 2 + (0.4375 - 3 * VS) ** 2 + 0.0) + 0.1)
    assert gain[0] == seq_sum([0.001, disc, 0.001])


def test_case4_terminal_when_nothing_unknown():
    o = small_world(observed_cols=(0, 1))
    p = orc.make_params(orc.EVAL_RRT_STAR, half_fov=1, non_path_voxel_gain=0.001)
    gain, cnt, term = o.gain_batch(p, [[0.03, 0.20]])
    assert cnt[0].tolist() == [7, 0, 0, 0] and gain[0] == 0.0 and term[0] == 1
    pg = orc.make_params(orc.EVAL_GOAL_DISTANCE, half_fov=1)
    gain, _, term = o.gain_batch(pg, [[0.03, 0.20]])
    assert gain[0] == 0.0 and term[0] == 0            # GoalDistance never sets gain_terminal


def test_case5_off_map_centre():
    o = small_world()
    for ev in (orc.EVAL_RRT_STAR, orc.EVAL_EXPAND_REACHABLE):
        gain, cnt, term = o.gain_batch(orc.make_params(ev, half_fov=1), [[0.50, 0.20]])
        assert cnt[0].tolist() == [0, 0, 0, 0] and gain[0] == 0.0 and term[0] == 1
    gain, _, term = o.gain_batch(orc.make_params(orc.EVAL_GOAL_DISTANCE, half_fov=1), [[0.50, 0.20]])
    assert gain[0] == 0.0 and term[0] == 0
    # AverageViewCost: 0 * (1 / mean_cost_)
    gain, _, _ = o.gain_batch(orc.make_params(orc.EVAL_AVERAGE_VIEW_COST, half_fov=1), [[0.50, 0.20]])
    m = o.map_stats()[0]
    assert (math.isnan(gain[0]) if m == 0 else gain[0] == 0.0)


def test_negative_position_quirks():
    """(int) truncation toward zero (map_planexp.cpp:134-137): x in (-vs,0) -> cell 0; x in (-2vs,-vs] -> loc -1,
    centre_loc -0.5 -> (int) 0 -> 'on map', window centred on cell 0, centre voxel at x = -0.5 vs reads as UNKNOWN
    (map_server_planexp.cpp:764-769)."""
    o = small_world()
    p = orc.make_params(orc.EVAL_RRT_STAR, half_fov=1)
    _, cnt, _ = o.gain_batch(p, [[-0.03, 0.20]])
    assert cnt[0].tolist() == [7, 3, 1, 0]
    _, cnt, _ = o.gain_batch(p, [[-0.07, 0.20]])
    # centre voxel (-0.03125, 0.21875) is off the map server's bounds -> UNKNOWN; its path lookup hits cell (0,3) -> on path
    assert cnt[0].tolist() == [7, 4, 2, 0]
    _, cnt, term = o.gain_batch(p, [[-0.13, 0.20]])
    assert cnt[0].tolist() == [0, 0, 0, 0] and term[0] == 1
    # bounding volume drops the negative centre voxel but keeps the window
    pb = orc.make_params(orc.EVAL_RRT_STAR, half_fov=1, bounding_volume=(0.0, 0.5, 0.0, 0.5))
    _, cnt, _ = o.gain_batch(pb, [[-0.07, 0.20]])
    assert cnt[0].tolist() == [6, 3, 1, 0]


def test_expand_reachable_only_start_cell_is_reachable():
    """map_server_planexp.cpp:476,830,846,853-869: flood-filled areas get ids >= 2, only the start cell is MSR_REACHABLE."""
    o = small_world()
    ids = o.reach_ids()
    assert (ids == 1).sum() == 1 and ids[0, 0] == 1      # start (0.03,0.03) -> cell (0,0)
    assert set(np.unique(ids).tolist()) <= {0, 1, 2}
    p = orc.make_params(orc.EVAL_EXPAND_REACHABLE, half_fov=1, discount_factor=0.25)
    gain, cnt, term = o.gain_batch(p, [[0.01, 0.01]])
    # unknown voxels (1,0),(1,1) both have the start cell (0,0) in their 3x3 neighbourhood
    assert cnt[0].tolist() == [5, 2, 2, 0] and gain[0] == 2.0
    gain, cnt, _ = o.gain_batch(p, [[0.20, 0.21]])
    assert cnt[0].tolist() == [10, 10, 0, 0] and gain[0] == 10 * 0.25


def test_faithful_and_linear_filters_agree():
    rng = np.random.default_rng(5)
    W = H = 96
    d = orc.make_desc(W, H, W * VS, H * VS, (0.5, 0.5), (5.0, 5.0), 30, 120, 30, compute_closest_observed=True)
    o = orc.Oracle(d)
    labels = rng.integers(0, 121, (H, W)).astype(np.int32)
    for _ in range(6):
        x0, y0 = rng.integers(-10, W - 10, 2)
        o.map_update_patch(int(x0), int(y0), labels[max(y0, 0):y0 + 33, max(x0, 0):x0 + 33] if False else
                           np.ascontiguousarray(rng.integers(0, 121, (33, 33)).astype(np.int32)))
    o.plan()
    xy = rng.uniform(-0.2, W * VS + 0.2, (200, 2))
    for ev in range(6):
        for kw in ({}, {"use_distance_discount": True}, {"do_binary_check": True}):
            p = orc.make_params(ev, half_fov=7, non_path_voxel_gain=0.001, **kw)
            a = o.gain_batch(p, xy, start_xy=xy[::-1].copy(), variant=0)
            b = o.gain_batch(p, xy, start_xy=xy[::-1].copy(), variant=1, threads=3)
            assert np.array_equal(a[0], b[0], equal_nan=True) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])


def test_closest_observed_and_expand_low_cost_hand_case():
    """setClosestObservedMaps (map_server_planexp.cpp:803-823) + ExpandLowCostAreaEvaluator (expand_low_cost_area_evaluator.cpp:33-72),
    derived by hand: a 4x4 patch at cells [20,24)x[20,24) of an all-unknown 48x48 map. The reference clamps to the CLOSED box
    [20..24]x[20..24] (bottom_right = top_left + size, one past the patch), so cells right of / below the patch take the cost of
    the unobserved column / row 24 (0.0), cells left / above it the patch's first column / row."""
    patch = (40 + np.arange(16).reshape(4, 4)).astype(np.int32)       # label(row r, col c) = 40 + 4 r + c
    # unit quirk: the initial distance is sqrt(width^2 + height^2) in METRES (map_server_planexp.cpp:470-471) but it is compared
    # with PIXEL distances (:817-818): on a 3 m x 3 m map (48 cells) nothing farther than 4.24 cells is ever updated
    d = orc.make_desc(48, 48, 48 * VS, 48 * VS, (0.5, 0.5), (2.5, 2.5), 30, 120, 30, compute_closest_observed=True)
    o = orc.Oracle(d)
    o.map_update_patch(20, 20, patch)
    cc, dc = o.closest_download()
    assert dc[22, 16] == 4.0 and cc[22, 16] == patch[2, 0] and dc[22, 15] == np.sqrt(18.0) and cc[22, 15] == 0.0
    W = H = 256                                                        # 16 m x 16 m: initial distance 22.6 > the 15 / 10 cell margins
    d = orc.make_desc(W, H, W * VS, H * VS, (0.5, 0.5), (2.5, 2.5), 30, 120, 30, compute_closest_observed=True)
    o = orc.Oracle(d)
    o.map_update_patch(20, 20, patch)
    cc, dc = o.closest_download()
    d0 = np.sqrt((W * VS) ** 2 + (H * VS) ** 2)
    assert cc[22, 18] == patch[2, 0] and dc[22, 18] == 2.0            # left of the patch: first patch column, distance 2
    assert cc[17, 21] == patch[0, 1] and dc[17, 21] == 3.0            # above: first patch row
    assert cc[22, 30] == 0.0 and dc[22, 30] == 6.0                    # right: column 24 is outside the patch and unobserved
    assert cc[18, 17] == patch[0, 0] and dc[18, 17] == np.sqrt(13.0)  # corner
    assert dc[22, 4] == d0 and cc[22, 4] == 0.0                       # 16 columns away: outside the 15-column margin
    assert dc[9, 22] == d0 and dc[10, 22] == 10.0                     # 10-row margin
    assert dc[21, 21] == d0                                           # observed cells are never touched
    # a second, nearer patch replaces only where it is strictly closer (`<=` keeps the old value on ties)
    o.map_update_patch(12, 20, (90 * np.ones((4, 4))).astype(np.int32))   # cells [12,16)x[20,24), closed box reaches column 16
    cc2, dc2 = o.closest_download()
    assert cc2[22, 18] == patch[2, 0] and dc2[22, 18] == 2.0          # tie (distance 2 to column 16, which is unobserved anyway)
    assert cc2[22, 17] == 0.0 and dc2[22, 17] == 1.0                  # column 16 (unobserved, cost 0) is now the closest
    # evaluator: window of half_fov 1 around cell (18, 22): the nine cells are all unknown; the centre voxel counts twice
    p = orc.make_params(orc.EVAL_EXPAND_LOW_COST, half_fov=1)
    gain, cnt, term = o.gain_batch(p, [[(18 + 0.3) * VS, (22 + 0.4) * VS]])
    cells = [(x, y) for x in (17, 18, 19) for y in (21, 22, 23)] + [(18, 22)]
    want = sum(1.0 / cc2[y, x] for x, y in cells if cc2[y, x] > 0)
    assert cnt[0, 0] == 10 and cnt[0, 1] == 10 and term[0] == 0
    assert abs(gain[0] - want) <= 1e-12 * want
    # augment_unknown_with_mean: the cells without a closest cost add 1 / mean_cost_ each
    p = orc.make_params(orc.EVAL_EXPAND_LOW_COST, half_fov=1, augment_unknown_with_mean=True)
    gain2, _, _ = o.gain_batch(p, [[(18 + 0.3) * VS, (22 + 0.4) * VS]])
    n0 = sum(1 for x, y in cells if not cc2[y, x] > 0)
    assert n0 > 0 and abs(gain2[0] - (want + n0 / o.map_stats()[0])) <= 1e-9 * gain2[0]
