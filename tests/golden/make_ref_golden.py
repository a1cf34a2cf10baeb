"""Generates tests/golden/ref_evaluators.npz: input/output vectors of the REFERENCE'S OWN evaluator / updater / cost sources
(advanced_planner_modules/src/..., compiled unmodified into oracle/_ref/libsipp_ref.so by `make -C oracle _ref`).

Run in the build container (needs /root/reference):   python tests/golden/make_ref_golden.py
The fixture pins the restated oracle (tests/test_ref_pin.py::test_oracle_matches_committed_reference_vectors) and the CUDA path
(tests/test_gpu_parity.py::test_gain_matches_committed_reference_vectors) wherever the reference tree itself is not available.

Scenario: a 160 x 128 map (10 m x 8 m) that receives 14 seeded 33 x 33 label patches (20 % obstacles) through the map-update path;
the planes the reference modules read (state, cost, reachability, closest-observed cost, path mask, mean / max cost) are the
oracle's after those patches and one plan; they are stored in the fixture together with the patches that produce them."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scouting-ipp_b200"))
from oracle import oracle as orc  # noqa: E402
from oracle import ref  # noqa: E402

W, H = 160, 128
DESC = dict(W=W, H=H, width=W * 0.0625, height=H * 0.0625, start=(0.5, 0.5), goal=(W * 0.0625 - 0.5, H * 0.0625 - 0.5), min_cost=30, max_cost=120,
            init_cost=120, compute_closest_observed=True)
EVAL_SETS = [
    dict(evaluator=0), dict(evaluator=0, non_path_voxel_gain=0.001), dict(evaluator=0, do_binary_check=True),
    dict(evaluator=0, do_binary_check=True, use_distance_discount=True), dict(evaluator=0, use_distance_discount=True, non_path_voxel_gain=0.005),
    dict(evaluator=1, max_gain=100.0), dict(evaluator=2, discount_factor=0.1), dict(evaluator=3), dict(evaluator=4),
    dict(evaluator=4, augment_unknown_with_mean=True),
]
HALF_FOVS = (1, 10, 32)
BVS = (None, (1.0, 7.5, 0.7, 6.1))
UPDATER_SETS = [(0, 0), (0, 1), (1, 0), (2, 1), (2, 0)]


def make_desc(mk):
    return mk(DESC["W"], DESC["H"], DESC["width"], DESC["height"], DESC["start"], DESC["goal"], DESC["min_cost"], DESC["max_cost"], DESC["init_cost"],
              compute_closest_observed=DESC["compute_closest_observed"])


def patches():
    rng = np.random.default_rng(3)
    out = []
    for _ in range(14):
        x0, y0 = int(rng.integers(-12, W - 10)), int(rng.integers(-12, H - 10))
        patch = rng.integers(0, 121, (33, 33)).astype(np.int32)
        patch[rng.random((33, 33)) < 0.2] = 0
        out.append((x0, y0, patch))
    return out


def poses(n=300):
    rng = np.random.default_rng(4)
    xy = rng.uniform(-0.3, W * 0.0625 + 0.3, (n, 2))
    xy[:, 1] = rng.uniform(-0.3, H * 0.0625 + 0.3, n)
    return xy, xy[::-1].copy()


def tree(n=300):
    rng = np.random.default_rng(5)
    par = np.full(n, -1, np.int32)
    kids = [[] for _ in range(n)]
    for i in range(1, n):
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
    cost = rng.uniform(0.2, 5.0, n)
    cost[0] = 0.0
    gain0 = np.where(rng.random(n) < 0.3, 0.0, rng.uniform(0.0, 50.0, n))
    gain0[0] = 0.0
    term0 = (rng.random(n) < 0.15).astype(np.uint8)
    term0[0] = 0
    value0 = rng.uniform(0.0, 10.0, n)
    return par, end, start, cost, gain0, term0, value0


def segments(m=120):
    rng = np.random.default_rng(6)
    npts = rng.integers(0, 6, m)
    off = np.concatenate([[0], np.cumsum(npts)]).astype(np.int32)
    pts = np.zeros((off[-1], 3))
    pts[:, 0] = rng.uniform(-0.5, W * 0.0625 + 0.5, off[-1])
    pts[:, 1] = rng.uniform(-0.5, H * 0.0625 + 0.5, off[-1])
    pts[1::7] = pts[0::7][: len(pts[1::7])] + 1e-3          # hops below the 5 mm branch, and a non-zero z
    return off, pts, rng.uniform(0, 10, m)


def main():
    o = orc.Oracle(make_desc(orc.make_desc))
    for x0, y0, patch in patches():
        o.map_update_patch(x0, y0, patch)
    o.plan()
    r = ref.Reference(make_desc(orc.make_desc)).snapshot(o)
    state, cost, reach, path = o.map_download()
    closest, _ = o.closest_download()
    mean, mx, _ = o.map_stats()
    out = dict(state=state, cost=cost, reach=reach, path=path, closest=closest, mean_max=np.array([mean, mx]))
    xy, sxy = poses()
    k = 0
    for kw in EVAL_SETS:
        for hf in HALF_FOVS:
            for bv in BVS:
                g, u, t = r.gain_batch(orc.make_params(half_fov=hf, bounding_volume=bv, **kw), xy, sxy)
                out["gain_%03d" % k], out["unk_%03d" % k], out["term_%03d" % k] = g, u, t
                k += 1
    par, end, start, tcost, gain0, term0, value0 = tree()
    k = 0
    for upd, fol in UPDATER_SETS:
        for kw in (EVAL_SETS[0], EVAL_SETS[3], EVAL_SETS[5], EVAL_SETS[6], EVAL_SETS[7]):
            for pc in (False, True):
                for flags in ((True, False, True), (False, False, True), (True, False, False)):
                    u = orc.make_updater(upd, fol, 0.001, 4.0, *flags)
                    g, v, t, _ = r.update_tree(orc.make_params(half_fov=10, **kw), u, par, end, gain0, tcost, value0, term0, (3.1, 2.7), pc, start_xy=start)
                    out["tg_%03d" % k], out["tv_%03d" % k], out["tt_%03d" % k] = g, v, t
                    k += 1
    # update_cost with the map-integral cost computer: the pass re-integrates every followed segment
    k = 0
    for upd, fol in UPDATER_SETS:
        u = orc.make_updater(upd, fol, 0.001, 4.0, True, True, True)
        g, v, t, c = r.update_tree(orc.make_params(half_fov=10, **EVAL_SETS[5]), u, par, end, gain0, tcost, value0, term0, (3.1, 2.7), True, start_xy=start,
                                   cost_params=orc.make_cost_params(False, 0.05))
        out["cg_%03d" % k], out["cv_%03d" % k], out["ct_%03d" % k], out["cc_%03d" % k] = g, v, t, c
        k += 1
    off, pts, pcost = segments()
    k = 0
    for acc in (False, True):
        for msd in (0.05, 0.3):
            c, vd = r.cost_integral(orc.make_cost_params(acc, msd), off, pts, pcost if acc else None)
            out["ci_%d" % k], out["civ_%d" % k] = c, vd
            k += 1
    path_out = os.path.join(ROOT, "tests", "golden", "ref_evaluators.npz")
    np.savez_compressed(path_out, **out)
    print("wrote %s (%d arrays, %.1f KB)" % (path_out, len(out), os.path.getsize(path_out) / 1024.0))


if __name__ == "__main__":
    main()
