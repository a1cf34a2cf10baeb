"""GPU parity of the batched ensemble path (sipp_batch_*, SURVEY.md 8(f)3 / BASELINE config 5): m maps advanced by one launch per
kernel must leave exactly what m independent contexts leave — fields, masks, plan results, gains and selected candidates bit for
bit — and match the CPU oracle's episode trace."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import sipp
from sipp import episode, synth
from oracle import oracle as orc

W, H = 640, 480


def make_map(seed, mk):
    lab = synth.make_labels(W, H, seed)
    ids = sorted(set(np.unique(lab).tolist()) - {0})
    desc = mk(W, H, W * synth.MPP, H * synth.MPP, (0.5, 0.5), (W * synth.MPP - 0.5, H * synth.MPP - 0.5), min(ids), max(ids), max(ids))
    return lab, desc


@pytest.mark.parametrize("roam", ["2", "1", "0"])
def test_batch_steps_match_independent_contexts(roam, monkeypatch):
    """ingest / plan / cycle through the batch against the same calls made map by map: from scratch, then incrementally; with the
    batch's shared relaxation queue (2: weights read from global memory, the default; 1: staged in shared memory) and with per-map relaxation workers (0)"""
    monkeypatch.setenv("SIPP_BATCH_ROAM", roam)
    m, T = 5, 777
    params = sipp.make_params(sipp.EVAL_RRT_STAR, half_fov=32, non_path_voxel_gain=0.001)
    labs, single, batched = [], [], []
    for k in range(m):
        lab, desc = make_map(3100 + k, sipp.make_desc)
        labs.append(lab)
        single.append(sipp.Context(desc))
        batched.append(sipp.Context(desc))
    b = sipp.Batch(batched)
    rng = np.random.default_rng(5)
    pos = [(0.5, 0.5)] * m
    # the simulator's initial square, then three cycles with a patch each
    x0 = np.zeros(m, np.int32)
    patches = [lab[0:320, 0:320] for lab in labs]
    b.ingest(x0, x0, patches)
    for k in range(m):
        single[k].map_update_patch(0, 0, patches[k])
    for cyc in range(4):
        xs, ys, ps = [], [], []
        for k in range(m):
            px, py = single[k].patch_origin(pos[k][0], pos[k][1], 65, 65)
            xs.append(px); ys.append(py)
            ps.append(episode.extract_patch(labs[k], px, py, 65, 65))
            single[k].map_update_patch(px, py, ps[-1])
        b.ingest(xs, ys, ps)
        plans = b.plan()
        xy = [np.ascontiguousarray(rng.uniform([0.2, 0.2], [W * synth.MPP - 0.2, H * synth.MPP - 0.2], (T, 2))) for _ in range(m)]
        cost = [rng.uniform(0.5, 3.0, T) for _ in range(m)]
        bi, bv, gains = b.cycle(params, xy, cost, want_gain=True)
        for k in range(m):
            c, st, _, n = single[k].plan(path_cap=1)
            assert plans[k] == (c, st, n), (cyc, k, plans[k], (c, st, n))
            assert single[k].plan_info()["incremental"] == batched[k].plan_info()["incremental"] == (cyc > 0)       # same start mode: scratch, then incremental
            assert np.array_equal(single[k].field_download(), batched[k].field_download()), (cyc, k)
            if cyc == 0:      # the relaxation statistics stay per map
                assert batched[k].plan_stats()["tile_activations"] >= (W // 32) * (H // 32) and batched[k].plan_stats()["aborted"] == 0
            ms, mb = single[k].map_download(), batched[k].map_download()
            for a_, b_ in zip(ms, mb):
                assert np.array_equal(a_, b_), (cyc, k)
            g, _, v, i, val = single[k].cycle(params, xy[k], cost[k])
            assert np.array_equal(g, gains[k]) and i == bi[k] and val == bv[k], (cyc, k, i, bi[k])
            pos[k] = (float(xy[k][i, 0]), float(xy[k][i, 1]))
    assert b.launch_count() > 0
    b.close()


def test_batch_episode_matches_native_episode_and_oracle():
    m, cycles, T = 6, 8, 512
    params = sipp.make_params(sipp.EVAL_RRT_STAR, half_fov=32, non_path_voxel_gain=0.001)
    seeds = [3200 + k for k in range(m)]
    labs, ctxs, alone = [], [], []
    for sd in seeds:
        lab, desc = make_map(sd, sipp.make_desc)
        labs.append(lab)
        ctxs.append(sipp.Context(desc))
        alone.append(sipp.Context(desc))
    cands = [episode.episode_candidates(W, H, sd, cycles, T) for sd in seeds]
    b = sipp.Batch(ctxs)
    traces = b.episode_run(labs, params, [c[0] for c in cands], [c[1] for c in cands])
    for k in range(m):
        assert traces[k] == alone[k].episode_run(labs[k], params, cands[k][0], cands[k][1]), k
        assert np.array_equal(alone[k].field_download(), ctxs[k].field_download())
    po = orc.make_params(orc.EVAL_RRT_STAR, half_fov=32, non_path_voxel_gain=0.001)
    for k in range(2):
        lab, desc = make_map(seeds[k], orc.make_desc)
        want = episode.run_episode(orc.Oracle(desc), lab, po, seeds[k], cycles=4, T=T)
        assert episode.traces_match(traces[k][:4], want), (k, traces[k][:4], want)
    b.close()


@pytest.mark.parametrize("kw", [dict(evaluator=sipp.EVAL_GOAL_DISTANCE, max_gain=100.0), dict(evaluator=sipp.EVAL_AVERAGE_VIEW_COST),
                                dict(evaluator=sipp.EVAL_RRT_STAR, use_distance_discount=True, non_path_voxel_gain=0.005),
                                dict(evaluator=sipp.EVAL_EXPAND_REACHABLE, discount_factor=0.1), dict(evaluator=sipp.EVAL_NAIVE),
                                dict(evaluator=sipp.EVAL_RRT_STAR, do_binary_check=True),
                                dict(evaluator=sipp.EVAL_GOAL_DISTANCE, max_gain=100.0, bounding_volume=(1.0, 30.0, 2.0, 25.0))])
def test_batch_cycle_serves_every_evaluator_family(kw):
    """the tiled floating-point evaluators (tensor maps per map in device memory) and the count-type ones through one batched launch:
    gains bit-identical to the same evaluator called map by map, the same selected candidate"""
    m, T = 4, 1500
    params = sipp.make_params(half_fov=32, **kw)
    labs, single, batched = [], [], []
    for k in range(m):
        lab, desc = make_map(3400 + k, sipp.make_desc)
        obs = synth.make_observed(W, H, 3400 + k, 10 + 2 * k)
        for group in (single, batched):
            c = sipp.Context(desc)
            c.map_upload_full(lab, obs)
            group.append(c)
    b = sipp.Batch(batched)
    plans = b.plan()
    rng = np.random.default_rng(11)
    xy = [np.ascontiguousarray(rng.uniform([-0.5, -0.5], [W * synth.MPP + 0.5, H * synth.MPP + 0.5], (T, 2))) for _ in range(m)]     # some views hang over the map edge
    cost = [rng.uniform(0.5, 3.0, T) for _ in range(m)]
    bi, bv, gains = b.cycle(params, xy, cost, want_gain=True)
    for k in range(m):
        c, st, _, n = single[k].plan(path_cap=1)
        assert plans[k] == (c, st, n)
        g, _, v, i, val = single[k].cycle(params, xy[k], cost[k])
        assert np.array_equal(g, gains[k], equal_nan=True), (kw, k, np.flatnonzero(g != gains[k])[:5])
        assert i == bi[k] and (val == bv[k] or (np.isnan(val) and np.isnan(bv[k]))), (kw, k, i, bi[k], val, bv[k])
    b.close()


@pytest.mark.parametrize("w,h,m", [(150, 100, 3), (33, 70, 2), (257, 129, 1)])
def test_batch_on_ragged_map_sizes(w, h, m):
    """map sizes that are no multiple of the 32-cell tile (partial tiles on two sides), a batch of one, a map narrower than a footprint:
    ingest / plan (from scratch, then incremental) / cycle through the batch against the same calls map by map"""
    params = sipp.make_params(sipp.EVAL_RRT_STAR, half_fov=32, non_path_voxel_gain=0.001)
    rng = np.random.default_rng(w * 1000 + h)
    labs, single, batched = [], [], []
    for k in range(m):
        lab = synth.make_labels(w, h, 3500 + k)
        ids = sorted(set(np.unique(lab).tolist()) - {0})
        desc = sipp.make_desc(w, h, w * synth.MPP, h * synth.MPP, (0.5, 0.5), (w * synth.MPP - 0.5, h * synth.MPP - 0.5), min(ids), max(ids), max(ids))
        labs.append(lab)
        single.append(sipp.Context(desc))
        batched.append(sipp.Context(desc))
    b = sipp.Batch(batched)
    for cyc in range(3):
        sz = 41
        xs = [int(rng.integers(-10, w)) for _ in range(m)]
        ys = [int(rng.integers(-10, h)) for _ in range(m)]
        if cyc == 0:
            xs, ys = [0] * m, [0] * m
        ps = [episode.extract_patch(labs[k], xs[k], ys[k], sz, sz) for k in range(m)]
        b.ingest(xs, ys, ps)
        plans = b.plan()
        T = 300
        xy = [np.ascontiguousarray(rng.uniform([-0.3, -0.3], [w * synth.MPP + 0.3, h * synth.MPP + 0.3], (T, 2))) for _ in range(m)]
        cost = [rng.uniform(0.5, 3.0, T) for _ in range(m)]
        bi, bv, gains = b.cycle(params, xy, cost, want_gain=True)
        for k in range(m):
            single[k].map_update_patch(xs[k], ys[k], ps[k])
            c, st, _, n = single[k].plan(path_cap=1)
            assert plans[k] == (c, st, n), (cyc, k)
            assert np.array_equal(single[k].field_download(), batched[k].field_download(), equal_nan=True), (cyc, k)
            g, _, v, i, val = single[k].cycle(params, xy[k], cost[k])
            assert np.array_equal(g, gains[k]) and i == bi[k] and val == bv[k], (cyc, k)
    b.close()


def test_batch_refuses_what_it_does_not_serve():
    lab, desc = make_map(3300, sipp.make_desc)
    a, c = sipp.Context(desc), sipp.Context(desc)
    with pytest.raises(sipp.SippError):
        sipp.Batch([a, a])                                      # one context twice
    with pytest.raises(sipp.SippError):
        sipp.Batch([a, sipp.Context(synth.world(320, 256, 7, 4, sipp.make_desc)[2])])      # another map size
    b = sipp.Batch([a, c])
    b.ingest([0, 0], [0, 0], [lab[:320, :320], lab[:320, :320]])
    b.plan()
    xy, cost = synth.make_candidates(W, H, 1, 64)
    with pytest.raises(sipp.SippError) as e:                    # its closest-observed planes are not maintained by the batched ingest
        b.cycle(sipp.make_params(sipp.EVAL_EXPAND_LOW_COST, half_fov=32), [xy, xy], [cost, cost])
    assert e.value.code == sipp.ERR_UNSUPPORTED
    bi, bv = b.cycle(sipp.make_params(sipp.EVAL_RRT_STAR, half_fov=32, non_path_voxel_gain=0.001), [xy, xy], [cost, cost])
    assert bi[0] == bi[1] and bv[0] == bv[1]
    b.close()
