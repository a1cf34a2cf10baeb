"""GPU parity of segments viewed from several sampled viewpoints (row A2: sensor_model/sampling_time > 0,
sensor_model_planexp.cpp:18-90): sipp_gain_batch_views / sipp_update_tree_views against the vectors the reference's OWN sources
produced (tests/golden/ref_views.npz, made by tests/golden/make_ref_views_golden.py) and against the oracle."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import sipp
from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_ref_golden as G  # noqa: E402
import make_ref_views_golden as GV  # noqa: E402

RTOL = 1e-5


def exact_gain(kw):        # integer-valued gains are compared bit for bit, sums of non-integers within the north_star tolerance
    return (kw["evaluator"] == 0 and not kw.get("use_distance_discount") and not kw.get("non_path_voxel_gain")) or kw["evaluator"] == 5


@pytest.fixture(scope="module")
def pair():
    g, o = sipp.Context(G.make_desc(sipp.make_desc)), orc.Oracle(G.make_desc(orc.make_desc))
    for x0, y0, patch in G.patches():
        g.map_update_patch(x0, y0, patch)
        o.map_update_patch(x0, y0, patch)
    rg, ro = g.plan(), o.plan()
    assert rg[0] == ro[0]
    return g, o


def views_of(traj, times, sampling_time):
    idx = orc.sample_viewpoints(times, sampling_time)
    off = np.arange(0, (len(traj) + 1) * len(idx), len(idx), dtype=np.int32)
    return off, np.ascontiguousarray(traj[:, idx, :]).reshape(-1, 2)


def test_cuda_matches_the_reference_vectors_for_sampled_viewpoints(pair):
    g, _ = pair
    vec = np.load(os.path.join(ROOT, "tests", "golden", "ref_views.npz"))
    traj, times = GV.trajectories()
    k = 0
    for st in GV.SAMPLING:
        off, vxy = views_of(traj, times, st)
        for kw in G.EVAL_SETS:
            for hf in GV.HALF_FOVS:
                for bv in GV.BVS:
                    gg, gc, gt = g.gain_batch_views(sipp.make_params(half_fov=hf, bounding_volume=bv, **kw), off, vxy, traj[:, 0, :])
                    assert np.array_equal(gc[:, 1], vec["unk_%03d" % k]) and np.array_equal(gt, vec["term_%03d" % k]), (st, kw, hf, bv)
                    if exact_gain(kw):
                        assert np.array_equal(gg, vec["gain_%03d" % k], equal_nan=True), (st, kw, hf, bv)
                    else:
                        assert np.allclose(gg, vec["gain_%03d" % k], rtol=RTOL, atol=0.0, equal_nan=True), (st, kw, hf, bv)
                    k += 1


def test_one_view_per_segment_is_the_single_view_call(pair):
    g, _ = pair
    xy, sxy = G.poses()
    off = np.arange(len(xy) + 1, dtype=np.int32)
    for kw in G.EVAL_SETS:
        p = sipp.make_params(half_fov=10, **kw)
        a = g.gain_batch(p, xy, sxy)
        b = g.gain_batch_views(p, off, xy, sxy)
        assert np.array_equal(a[0], b[0], equal_nan=True) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]), kw


def test_ragged_and_empty_view_lists_against_the_oracle(pair):
    g, o = pair
    rng = np.random.default_rng(9)
    n = 64
    nv = rng.integers(0, 5, n)
    nv[:3] = 0                                       # segments without a view: an empty voxel list
    off = np.concatenate([[0], np.cumsum(nv)]).astype(np.int32)
    vxy = np.stack([rng.uniform(-0.3, G.W * 0.0625 + 0.3, off[-1]), rng.uniform(-0.3, G.H * 0.0625 + 0.3, off[-1])], 1)
    sxy = np.stack([rng.uniform(0, 10, n), rng.uniform(0, 8, n)], 1)
    for kw in G.EVAL_SETS:
        pg, po = sipp.make_params(half_fov=16, **kw), orc.make_params(half_fov=16, **kw)
        gg, gc, gt = g.gain_batch_views(pg, off, vxy, sxy)
        og, oc, ot = o.gain_batch_views(po, off, vxy, sxy)
        assert np.array_equal(gc, oc) and np.array_equal(gt, ot), kw
        assert np.allclose(gg, og, rtol=RTOL, atol=0.0, equal_nan=True), kw
    with pytest.raises(sipp.SippError):
        g.gain_batch_views(sipp.make_params(half_fov=16, evaluator=0), np.array([1, 2], np.int32), vxy[:2])        # offsets must start at 0


@pytest.mark.parametrize("upd,fol", [(0, 0), (0, 1), (1, 0), (2, 1)])
def test_update_tree_with_sampled_viewpoints_matches_the_oracle_pass(pair, upd, fol):
    g, o = pair
    par, end, start, tcost, gain0, term0, value0 = G.tree()
    n = len(par)
    rng = np.random.default_rng(13)
    nv = rng.integers(1, 4, n)
    off = np.concatenate([[0], np.cumsum(nv)]).astype(np.int32)
    vxy = np.empty((off[-1], 2))
    for i in range(n):                               # views along the segment start -> end, the last one at the end point
        s = np.linspace(1.0 / nv[i], 1.0, nv[i])[:, None]
        vxy[off[i]:off[i + 1]] = start[i] * (1 - s) + end[i] * s
    for kw in (G.EVAL_SETS[0], G.EVAL_SETS[3], G.EVAL_SETS[5], G.EVAL_SETS[6], G.EVAL_SETS[7]):
        for pc in (False, True):
            ug, uo = sipp.make_updater(upd, fol, 0.001, 4.0, True, False, True), orc.make_updater(upd, fol, 0.001, 4.0, True, False, True)
            gg, gv, gt, _ = g.update_tree(sipp.make_params(half_fov=10, **kw), ug, par, end, gain0, tcost, value0, term0, (3.1, 2.7), pc, start_xy=start,
                                          view_off=off, view_xy=vxy)
            og, ov, ot = o.update_tree(orc.make_params(half_fov=10, **kw), uo, par, end, gain0, tcost, value0, term0, (3.1, 2.7), pc, start_xy=start,
                                       view_off=off, view_xy=vxy)
            assert np.array_equal(gt, ot), (kw, pc)
            assert np.allclose(gg, og, rtol=RTOL, atol=0.0, equal_nan=True) and np.allclose(gv, ov, rtol=RTOL, atol=0.0, equal_nan=True), (kw, pc)
