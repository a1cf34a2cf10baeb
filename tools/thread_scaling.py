"""Which ABI call stops scaling when many contexts are driven from many host threads? Each thread loops one call on its own
640x480 context; prints calls/s for 1, 4, 16 threads."""
import os, sys, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scouting-ipp_b200"))
import numpy as np
import sipp
from sipp import synth, episode
W, H, T = 640, 480, 1024
NT = 16
ctxs, labs = [], []
for i in range(NT):
    lab = synth.make_labels(W, H, 2000 + i); ids = sorted(set(np.unique(lab).tolist()) - {0})
    c = sipp.Context(sipp.make_desc(W, H, W * synth.MPP, H * synth.MPP, (0.5, 0.5), (W * synth.MPP - 0.5, H * synth.MPP - 0.5), min(ids), max(ids), max(ids)))
    c.map_update_patch(0, 0, lab[:320, :320]); c.plan(path_cap=1)
    ctxs.append(c); labs.append(lab)
p = sipp.make_params(sipp.EVAL_RRT_STAR, half_fov=32, non_path_voxel_gain=0.001)
xy, cost = synth.make_candidates(W, H, 1, T)
gain = np.zeros(T); term = np.zeros(T, np.uint8); val = np.zeros(T)
patch = np.ascontiguousarray(labs[0][100:165, 100:165])
calls = {
    "plan": lambda c: c.plan(path_cap=1),
    "cycle": lambda c: c.cycle_host_ptr(p, T, xy.ctypes.data, None, cost.ctypes.data, gain.ctypes.data, term.ctypes.data, val.ctypes.data),
    "update_patch": lambda c: c.map_update_patch(100, 100, patch),
}
for name, fn in calls.items():
    for nt in (1, 4, 16):
        n = 60
        def work(c):
            for _ in range(n): fn(c)
        ths = [threading.Thread(target=work, args=(ctxs[i],)) for i in range(nt)]
        t0 = time.perf_counter()
        for t in ths: t.start()
        for t in ths: t.join()
        dt = time.perf_counter() - t0
        print("%-14s threads %2d: %8.0f calls/s  (%.3f ms per call per thread)" % (name, nt, nt * n / dt, dt / n * 1e3), flush=True)
