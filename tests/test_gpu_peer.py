"""GPU parity of the multi-GPU step: candidates sharded over several contexts, shard winners exchanged
through peer memory by the argmax kernel itself (sipp_cycle_shard_dev). On the one-GPU test box the "ranks" are contexts of
one process on cuda:0 (same protocol, mailboxes addressed directly instead of through CUDA IPC); bench.py --gpus N runs it
across processes / GPUs. The result must be the candidate ONE context scanning every candidate selects — and that is
compared with the CPU oracle."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import sipp
from sipp import shard, synth
from oracle import oracle as orc


def _attach_all(ctxs):
    handles = [c.peer_export() for c in ctxs]
    for r, c in enumerate(ctxs):
        c.peer_attach(r, handles)


def _rec(t):
    h = t.cpu().numpy()
    return float(h[:1].view(np.float64)[0]), int(h[1])


@pytest.mark.parametrize("world", [1, 2, 3, 5])
def test_sharded_cycle_through_peer_memory_matches_single_context_and_oracle(world):
    import torch
    W, H, n_total = 320, 240, 1500
    lab, obs, dg = synth.world(W, H, 42, 8, sipp.make_desc, init_cost="max")
    _, _, do = synth.world(W, H, 42, 8, orc.make_desc, init_cost="max")
    o = orc.Oracle(do)
    o.map_upload_full(lab, obs)
    o.plan()
    ctxs = [sipp.Context(dg) for _ in range(world)]
    for c in ctxs:                       # map replicated on every rank; every rank recomputes the same field / mask
        c.map_upload_full(lab, obs)
        c.plan()
    _attach_all(ctxs)
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(5)
    for step in range(5):                # several steps: the two mailbox slots alternate
        xy, cost = synth.make_candidates(W, H, 100 + step, n_total)
        if step == 1:                    # exact cross-shard tie: the lowest global index must win
            j = int(np.argmax(o.cycle(orc.make_params(orc.EVAL_RRT_STAR, half_fov=10), xy, cost, variant=1)[2]))
            xy[n_total - 1], cost[n_total - 1] = xy[j], cost[j]
        kw = dict(half_fov=10, non_path_voxel_gain=0.0 if step % 2 else 0.001)
        pg, po = sipp.make_params(sipp.EVAL_RRT_STAR, **kw), orc.make_params(orc.EVAL_RRT_STAR, **kw)
        # every buffer is allocated and filled BEFORE the first launch: a device-wide synchronisation (torch's allocator, a
        # pageable copy) while one rank already waits in its exchange would wait for the ranks that are not launched yet
        bufs = []
        for r in range(world):
            lo, hi = shard.shard_range(n_total, r, world)
            bufs.append((lo, hi, torch.from_numpy(np.ascontiguousarray(xy[lo:hi])).to(dev), torch.from_numpy(np.ascontiguousarray(cost[lo:hi])).to(dev),
                         torch.empty(hi - lo, dtype=torch.float64, device=dev), torch.empty(hi - lo, dtype=torch.float64, device=dev),
                         torch.empty(hi - lo, dtype=torch.uint8, device=dev), torch.zeros(2, dtype=torch.int64, device=dev)))
        torch.cuda.synchronize()
        outs = []
        for r in rng.permutation(world):   # the ranks reach the exchange in any order
            lo, hi, d_xy, d_cost, d_gain, d_val, d_term, d_best = bufs[int(r)]
            ctxs[r].cycle_shard_dev(pg, hi - lo, d_xy.data_ptr(), None, d_cost.data_ptr(), d_gain.data_ptr(), d_term.data_ptr(),
                                    d_val.data_ptr(), lo, d_best.data_ptr())
            outs.append((int(r), lo, hi, d_gain, d_val, d_best))
        for c in ctxs:
            assert c.peer_status() == (0, step + 1)          # synchronises the context's stream
        og, ot, ov, oi, obv = o.cycle(po, xy, cost, variant=1)
        for r, lo, hi, d_gain, d_val, d_best in outs:
            bv, bi = _rec(d_best)
            assert (bi, bv) == (oi, obv), (step, r, bi, bv, oi, obv)
            assert np.allclose(d_gain.cpu().numpy(), og[lo:hi], rtol=1e-5, atol=0)
            assert np.allclose(d_val.cpu().numpy(), ov[lo:hi], rtol=1e-5, atol=0)
    for c in ctxs:
        c.peer_detach()
        c.close()


def test_pipelined_sharded_cycles_return_the_previous_winner_and_flush_the_last():
    """sipp_cycle_shard_pipelined_dev: call k publishes step k and returns the global winner of step k-1; the flush returns the
    last one. Same winners as the oracle scanning every candidate of the respective step."""
    import torch
    world, W, H, n_total, steps = 3, 320, 240, 900, 7
    lab, obs, dg = synth.world(W, H, 42, 8, sipp.make_desc, init_cost="max")
    _, _, do = synth.world(W, H, 42, 8, orc.make_desc, init_cost="max")
    o = orc.Oracle(do)
    o.map_upload_full(lab, obs)
    o.plan()
    ctxs = [sipp.Context(dg) for _ in range(world)]
    for c in ctxs:
        c.map_upload_full(lab, obs)
        c.plan()
    _attach_all(ctxs)
    dev = torch.device("cuda", 0)
    pg, po = sipp.make_params(sipp.EVAL_RRT_STAR, half_fov=10), orc.make_params(orc.EVAL_RRT_STAR, half_fov=10)
    want, bufs = [], []
    for step in range(steps):
        xy, cost = synth.make_candidates(W, H, 300 + step, n_total)
        _, _, _, oi, obv = o.cycle(po, xy, cost, variant=1)
        want.append((obv, oi))
        row = []
        for r in range(world):
            lo, hi = shard.shard_range(n_total, r, world)
            row.append((lo, hi, torch.from_numpy(np.ascontiguousarray(xy[lo:hi])).to(dev), torch.from_numpy(np.ascontiguousarray(cost[lo:hi])).to(dev),
                        torch.empty(hi - lo, dtype=torch.float64, device=dev), torch.empty(hi - lo, dtype=torch.float64, device=dev),
                        torch.empty(hi - lo, dtype=torch.uint8, device=dev), torch.zeros(2, dtype=torch.int64, device=dev)))
        bufs.append(row)
    last = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
    torch.cuda.synchronize()
    for step in range(steps):              # everything is queued without a host synchronisation in between
        for r in range(world):
            lo, hi, d_xy, d_cost, d_gain, d_val, d_term, d_prev = bufs[step][r]
            ctxs[r].cycle_shard_pipelined_dev(pg, hi - lo, d_xy.data_ptr(), None, d_cost.data_ptr(), d_gain.data_ptr(), d_term.data_ptr(),
                                              d_val.data_ptr(), lo, d_prev.data_ptr())
    for r in range(world):
        ctxs[r].peer_flush_dev(last[r].data_ptr())
    for c in ctxs:
        assert c.peer_status() == (0, steps)
    for r in range(world):
        assert _rec(bufs[0][r][7]) == (0.0, -1)                       # nothing to collect on the first call
        for step in range(1, steps):
            assert _rec(bufs[step][r][7]) == want[step - 1], (r, step)
        assert _rec(last[r]) == want[steps - 1]
    for c in ctxs:
        c.close()


def test_peer_exchange_of_records_matches_sequential_scan():
    """sipp_peer_exchange_dev on arbitrary records (ties, NaN, empty shards) against the sequential-scan rule of sipp.h"""
    import torch
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(11)
    for world in (2, 4, 8, 16):
        ctxs = [sipp.Context(sipp.make_desc(64, 64, 4.0, 4.0, (0.5, 0.5), (3.5, 3.5), 30, 120, 30)) for _ in range(world)]
        _attach_all(ctxs)
        for trial in range(12):
            vals = rng.choice([0.0, 1.5, 1.5, 2.25, -3.0, np.nan, np.inf], size=world)
            idx = np.array([int(rng.integers(0, 1000)) + 1000 * r if rng.random() > 0.25 else -1 for r in range(world)], np.int64)
            best, have = (vals[0], idx[0]), bool(idx[0] >= 0 and not np.isnan(vals[0]))
            for r in range(1, world):
                if idx[r] < 0 or np.isnan(vals[r]):
                    continue
                if not have or vals[r] > best[0] or (vals[r] == best[0] and idx[r] < best[1]):
                    best, have = (vals[r], idx[r]), True
            d_in = [torch.from_numpy(np.array([vals[r:r + 1].view(np.int64)[0], idx[r]], np.int64)).to(dev) for r in range(world)]
            outs = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
            torch.cuda.synchronize()     # nothing below synchronises the device until every rank has launched
            for r in rng.permutation(world):
                ctxs[r].peer_exchange_dev(d_in[r].data_ptr(), outs[r].data_ptr())
            for c in ctxs:
                assert c.peer_status()[0] == 0
            for d_out in outs:
                bv, bi = _rec(d_out)
                assert bi == int(best[1]) and (bv == best[0] or (np.isnan(bv) and np.isnan(best[0]))), (world, trial, vals, idx, bv, bi, best)
        for c in ctxs:
            c.close()


def test_peer_exchange_times_out_instead_of_hanging():
    """a rank whose peer never shows up gets index -3 and a non-zero status after SIPP_PEER_TIMEOUT_MS — never a hung GPU"""
    import torch
    dev = torch.device("cuda", 0)
    os.environ["SIPP_PEER_TIMEOUT_MS"] = "150"
    try:
        ctxs = [sipp.Context(sipp.make_desc(64, 64, 4.0, 4.0, (0.5, 0.5), (3.5, 3.5), 30, 120, 30)) for _ in range(2)]
        _attach_all(ctxs)
    finally:
        del os.environ["SIPP_PEER_TIMEOUT_MS"]
    d_in = torch.tensor([np.float64(1.0).view(np.int64), 5], dtype=torch.int64, device=dev)
    d_out = torch.zeros(2, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    ctxs[0].peer_exchange_dev(d_in.data_ptr(), d_out.data_ptr())      # rank 1 never calls
    status, steps = ctxs[0].peer_status()
    bv, bi = _rec(d_out)
    assert status == 1 and bi == -3 and np.isnan(bv)
    for c in ctxs:
        c.close()


def test_peer_argument_errors():
    g = sipp.Context(sipp.make_desc(64, 64, 4.0, 4.0, (0.5, 0.5), (3.5, 3.5), 30, 120, 30))
    with pytest.raises(sipp.SippError):
        g.peer_attach(0, [b"\0" * sipp.PEER_HANDLE_BYTES])            # attach before export
    h = g.peer_export()
    with pytest.raises(sipp.SippError):
        g.peer_attach(0, [b"\0" * sipp.PEER_HANDLE_BYTES])            # not a handle
    with pytest.raises(sipp.SippError):
        g.peer_attach(1, [h])                                         # rank out of range
    g.peer_attach(0, [h])
    assert g.peer_status() == (0, 0)
    g.peer_detach()
    with pytest.raises(sipp.SippError):
        g.peer_exchange_dev(1, 1)                                     # detached
    g.close()


def test_empty_shards_take_part_in_the_exchange():
    """fewer candidates than ranks: the ranks whose shard is empty publish {0, -1} ("never wins") so that the step counter advances
    everywhere and nobody waits for a record that never comes; every rank ends with the winner of the non-empty shards."""
    import torch
    world, W, H, n_total = 4, 160, 128, 2
    lab, obs, dg = synth.world(W, H, 42, 4, sipp.make_desc, init_cost="max")
    _, _, do = synth.world(W, H, 42, 4, orc.make_desc, init_cost="max")
    o = orc.Oracle(do)
    o.map_upload_full(lab, obs)
    o.plan()
    ctxs = [sipp.Context(dg) for _ in range(world)]
    for c in ctxs:
        c.map_upload_full(lab, obs)
        c.plan()
    _attach_all(ctxs)
    dev = torch.device("cuda", 0)
    xy, cost = synth.make_candidates(W, H, 9, n_total)
    pg, po = sipp.make_params(sipp.EVAL_RRT_STAR, half_fov=10), orc.make_params(orc.EVAL_RRT_STAR, half_fov=10)
    bufs = []
    for r in range(world):
        lo, hi = shard.shard_range(n_total, r, world)
        m = max(hi - lo, 1)
        bufs.append((lo, hi, torch.from_numpy(np.ascontiguousarray(xy[lo:lo + m])).to(dev), torch.from_numpy(np.ascontiguousarray(cost[lo:lo + m])).to(dev),
                     torch.empty(m, dtype=torch.float64, device=dev), torch.empty(m, dtype=torch.float64, device=dev),
                     torch.empty(m, dtype=torch.uint8, device=dev), torch.zeros(2, dtype=torch.int64, device=dev)))
    assert sum(1 for lo, hi, *_ in bufs if hi == lo) >= 2          # at least two ranks have nothing to score
    torch.cuda.synchronize()
    for r in range(world):
        lo, hi, d_xy, d_cost, d_gain, d_val, d_term, d_best = bufs[r]
        ctxs[r].cycle_shard_dev(pg, hi - lo, d_xy.data_ptr(), None, d_cost.data_ptr(), d_gain.data_ptr(), d_term.data_ptr(), d_val.data_ptr(), lo, d_best.data_ptr())
    for c in ctxs:
        assert c.peer_status() == (0, 1)
    _, _, _, oi, obv = o.cycle(po, xy, cost, variant=1)
    for r in range(world):
        assert _rec(bufs[r][7]) == (obv, oi), (r, _rec(bufs[r][7]), oi, obv)
    for c in ctxs:
        c.peer_detach()
        c.close()
