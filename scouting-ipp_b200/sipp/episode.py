"""Host loop of one exploration episode in the style of the reference's run_experiment.sh runs:
every planning cycle the scout's simulator publishes the fov x fov label patch around the scout (mav_simulator.cpp:438-479),
the map server ingests it (mapDataCallback) and re-plans the follower path, the evaluator scores the candidate views, the
best one becomes the scout's next position. Only host plumbing lives here; `backend` is anything with the map / plan / cycle
calls of sipp.Context (the tests drive their CPU checker through the same loop to compare traces)."""
import numpy as np

from . import synth


def extract_patch(truth, x0, y0, rows, cols):
    """rows x cols window of the ground-truth label image with its top-left at cell (x0, y0); cells outside the image are 0
    (the ingest skips off-map cells, map_server_planexp.cpp:238-243)."""
    H, W = truth.shape
    out = np.zeros((rows, cols), np.int32)
    ya, yb = max(0, y0), min(H, y0 + rows)
    xa, xb = max(0, x0), min(W, x0 + cols)
    if ya < yb and xa < xb:
        out[ya - y0:yb - y0, xa - x0:xb - x0] = truth[ya:yb, xa:xb]
    return out


def cycle_candidates(W, H, seed, cycle, T):
    """the T candidate views of one cycle (uniform over the map, SegmentTime costs): fixed by (seed, cycle) only"""
    return synth.make_candidates(W, H, seed * 1000 + cycle, T)


def run_episode(backend, truth, params, seed, cycles=40, T=1024, fov=65, init=320):
    """returns the trace [(path cost, plan status, path nodes, selected candidate, its value)] of `cycles` planning cycles"""
    H, W = truth.shape
    backend.map_update_patch(0, 0, truth[0:min(H, init), 0:min(W, init)])      # the simulator's initial square at the start
    pos = (0.5, 0.5)
    trace = []
    for c in range(cycles):
        x0, y0 = backend.patch_origin(pos[0], pos[1], fov, fov)
        backend.map_update_patch(x0, y0, extract_patch(truth, x0, y0, fov, fov))
        cost, status, _, n = backend.plan(path_cap=1)
        xy, ccost = cycle_candidates(W, H, seed, c, T)
        _, _, _, bi, bv = backend.cycle(params, xy, ccost)
        trace.append((cost, status, n, bi, bv))
        if bi >= 0:
            pos = (float(xy[bi, 0]), float(xy[bi, 1]))
    return trace


def episode_candidates(W, H, seed, cycles, T):
    """all candidate views of an episode as arrays [cycles][T][2] / [cycles][T] (the inputs of sipp_episode_run)"""
    xy = np.empty((cycles, T, 2), np.float64)
    cost = np.empty((cycles, T), np.float64)
    for c in range(cycles):
        xy[c], cost[c] = cycle_candidates(W, H, seed, c, T)
    return xy, cost


def traces_match(a, b, rtol=1e-5):
    """bit-exact path cost, status, path length and selected candidate; the winner's value within the floating-point gain tolerance"""
    if len(a) != len(b):
        return False
    for (c0, s0, n0, i0, v0), (c1, s1, n1, i1, v1) in zip(a, b):
        if (c0, s0, n0, i0) != (c1, s1, n1, i1) or abs(v0 - v1) > rtol * max(abs(v0), abs(v1)):
            return False
    return True
