"""Host-side pieces of the episode / ensemble path that need no GPU: patch extraction, candidate streams, trace comparison, and the
episode loop itself driven through the CPU oracle (the same loop drives libsipp in tests/test_gpu_parity.py)."""
import numpy as np

from sipp import episode, synth
from oracle import oracle as orc


def test_extract_patch_clips_to_the_image_and_zero_fills():
    truth = np.arange(1, 7 * 5 + 1, dtype=np.int32).reshape(5, 7)
    p = episode.extract_patch(truth, -2, -1, 4, 5)
    assert p.shape == (4, 5) and p.dtype == np.int32
    assert np.array_equal(p[0], np.zeros(5)) and np.array_equal(p[:, :2], np.zeros((4, 2)))
    assert np.array_equal(p[1:, 2:], truth[0:3, 0:3])
    assert np.array_equal(episode.extract_patch(truth, 5, 3, 3, 4), [[27, 28, 0, 0], [34, 35, 0, 0], [0, 0, 0, 0]])
    assert not episode.extract_patch(truth, 100, 100, 3, 3).any()


def test_candidate_streams_depend_on_seed_and_cycle_only():
    a = episode.cycle_candidates(64, 48, 7, 3, 16)
    b = episode.cycle_candidates(64, 48, 7, 3, 16)
    c = episode.cycle_candidates(64, 48, 7, 4, 16)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and not np.array_equal(a[0], c[0])
    xy, cost = episode.episode_candidates(64, 48, 7, 5, 16)
    assert xy.shape == (5, 16, 2) and cost.shape == (5, 16) and np.array_equal(xy[3], a[0]) and np.array_equal(cost[4], c[1])


def test_traces_match_rule():
    t = [(10.5, 0, 40, 3, 2.0), (9.0, 0, 38, 7, 1.0)]
    assert episode.traces_match(t, [(10.5, 0, 40, 3, 2.0 * (1 + 1e-9)), (9.0, 0, 38, 7, 1.0)])
    assert not episode.traces_match(t, [(10.5, 0, 40, 3, 2.1), (9.0, 0, 38, 7, 1.0)])       # value outside the tolerance
    assert not episode.traces_match(t, [(10.5, 0, 40, 4, 2.0), (9.0, 0, 38, 7, 1.0)])       # another candidate
    assert not episode.traces_match(t, t[:1])


def test_episode_loop_on_the_oracle_is_deterministic_and_explores():
    W, H = 160, 120
    lab = synth.make_labels(W, H, 2005)
    ids = sorted(set(np.unique(lab).tolist()) - {0})
    mk = lambda: orc.Oracle(orc.make_desc(W, H, W * synth.MPP, H * synth.MPP, (0.5, 0.5), (W * synth.MPP - 0.5, H * synth.MPP - 0.5), min(ids), max(ids), max(ids)))
    p = orc.make_params(orc.EVAL_RRT_STAR, half_fov=16)
    o1 = mk()
    t1 = episode.run_episode(o1, lab, p, 2005, cycles=5, T=64, fov=33, init=48)
    t2 = episode.run_episode(mk(), lab, p, 2005, cycles=5, T=64, fov=33, init=48)
    assert t1 == t2 and len(t1) == 5
    assert all(0 <= bi < 64 for _, _, _, bi, _ in t1)
    state = o1.map_download()[0]
    assert (state != 2).sum() > 48 * 48                  # more than the initial square has been observed
