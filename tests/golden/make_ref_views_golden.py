"""Generates tests/golden/ref_views.npz: segments viewed from SEVERAL sampled viewpoints (sensor_model/sampling_time > 0,
sensor_model_planexp.cpp:18-90) scored by the REFERENCE'S OWN sources (oracle/_ref/libsipp_ref.so: its sampleViewpoints picks the
trajectory points, its evaluators count the concatenated voxel lists).

Run in the build container (needs /root/reference):   python tests/golden/make_ref_views_golden.py
Scenario = the one of make_ref_golden.py. Trajectories: M segments of NPTS points each with common time stamps; every sampling time of
SAMPLING is scored with every evaluator set, two footprints, with and without a bounding volume."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_ref_golden as G  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from oracle import ref  # noqa: E402

M, NPTS = 90, 9
SAMPLING = (0.0, 0.5, 1.0, 1.7, 10.0)      # 0: last point only; 10 s: longer than the segment -> last point only
HALF_FOVS = (3, 16)
BVS = (None, (1.0, 7.5, 0.7, 6.1))


def trajectories():
    rng = np.random.default_rng(8)
    W, H = G.W * 0.0625, G.H * 0.0625
    a = np.stack([rng.uniform(-0.3, W + 0.3, M), rng.uniform(-0.3, H + 0.3, M)], 1)
    b = a + rng.uniform(-2.5, 2.5, (M, 2))
    bend = rng.uniform(-0.6, 0.6, (M, 2))
    s = np.linspace(0.0, 1.0, NPTS)[None, :, None]
    traj = a[:, None, :] * (1 - s) + b[:, None, :] * s + bend[:, None, :] * np.sin(np.pi * s)
    traj[:5, :, :] = traj[:5, :1, :]                       # the same point nine times: identical views back to back
    traj[5:10, 1:, :] = traj[5:10, 1:2, :] + np.arange(NPTS - 1)[None, :, None] * 0.0625 * np.array([1.0, 1.0])      # one cell per step along the diagonal
    times = (np.arange(NPTS) * 0.4e9).astype(np.int64)     # 0, 0.4, ..., 3.2 s
    return np.ascontiguousarray(traj), times


def main():
    o = orc.Oracle(G.make_desc(orc.make_desc))
    for x0, y0, patch in G.patches():
        o.map_update_patch(x0, y0, patch)
    o.plan()
    r = ref.Reference(G.make_desc(orc.make_desc)).snapshot(o)
    traj, times = trajectories()
    out = dict(traj=traj, times=times)
    k = 0
    for st in SAMPLING:
        for kw in G.EVAL_SETS:
            for hf in HALF_FOVS:
                for bv in BVS:
                    g, u, t = r.gain_batch(orc.make_params(half_fov=hf, bounding_volume=bv, **kw), traj[:, -1, :], traj[:, 0, :], sampling_time=st,
                                           traj_xy=traj, times_ns=times)
                    out["gain_%03d" % k], out["unk_%03d" % k], out["term_%03d" % k] = g, u, t
                    k += 1
    path = os.path.join(ROOT, "tests", "golden", "ref_views.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, k, "result sets,", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
