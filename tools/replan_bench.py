"""Re-plan after one 65x65 patch ingest (what every planning cycle does) — incremental (warm start + invalidation) against a plan
from scratch on the same map state. python tools/replan_bench.py [W] [patches]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scouting-ipp_b200"))
import numpy as np
import sipp
from sipp import synth, episode
W = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
NP = int(sys.argv[2]) if len(sys.argv) > 2 else 12
K = {640: 20, 1024: 100, 2048: 400, 4096: 1500, 8192: 6000}.get(W, 200)
H = 480 if W == 640 else W
lab = synth.make_labels(W, H, 1003); obs = synth.make_observed(W, H, 1003, K)
ids = sorted(set(np.unique(lab).tolist()) - {0})
mk = lambda: sipp.Context(sipp.make_desc(W, H, W * synth.MPP, H * synth.MPP, (0.5, 0.5), (W * synth.MPP - 0.5, H * synth.MPP - 0.5), min(ids), max(ids), max(ids)))
gi, gf = mk(), mk()
gf.set_incremental(False)
for g in (gi, gf):
    g.map_upload_full(lab, obs); g.plan(); g.plan()
rng = np.random.default_rng(5)
ti, tf, infos = [], [], []
# patches at the rim of the observed region: unobserved cells next to observed ones
ys, xs = np.nonzero((obs == 0) & (np.roll(obs, 40, 0) | np.roll(obs, 40, 1) | np.roll(obs, -40, 0) | np.roll(obs, -40, 1)).astype(bool))
for k in range(NP):
    j = int(rng.integers(0, len(xs)))
    x0, y0 = int(xs[j]) - 32, int(ys[j]) - 32
    patch = episode.extract_patch(lab, x0, y0, 65, 65)
    for g, ts in ((gi, ti), (gf, tf)):
        g.map_update_patch(x0, y0, patch)
        t0 = time.perf_counter(); r = g.plan(path_cap=1); ts.append((time.perf_counter() - t0) * 1e3)
        if g is gi: ri = r; infos.append(g.plan_info())
    assert ri[0] == r[0] and ri[3] == r[3], (ri, r)
    print("patch %2d at (%5d,%5d): incremental %.3f ms  scratch %.3f ms  cost %.3f  %s  relax %s" % (k, x0, y0, ti[-1], tf[-1], r[0], infos[-1], gi.plan_stats()), flush=True)
assert np.array_equal(gi.field_download(), gf.field_download())
print("W %d: re-plan after a 65x65 patch: incremental median %.3f ms (min %.3f, max %.3f), from scratch median %.3f ms; fields identical" % (
    W, float(np.median(ti)), min(ti), max(ti), float(np.median(tf))))
