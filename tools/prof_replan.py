"""Short fixed workload for ncu: one plan from scratch, then patch ingests + incremental re-plans. python tools/prof_replan.py [W] [patches]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scouting-ipp_b200"))
import numpy as np
import sipp
from sipp import synth, episode
W = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
NP = int(sys.argv[2]) if len(sys.argv) > 2 else 4
lab, obs, desc = synth.world(W, W, 1003, {4096: 1500}.get(W, 200), sipp.make_desc, init_cost="max")
g = sipp.Context(desc)
g.map_upload_full(lab, obs)
print("plan", g.plan(path_cap=1)[0], flush=True)
rng = np.random.default_rng(5)
ys, xs = np.nonzero((obs == 0) & (np.roll(obs, 40, 0) | np.roll(obs, 40, 1) | np.roll(obs, -40, 0) | np.roll(obs, -40, 1)).astype(bool))
for k in range(NP):
    j = int(rng.integers(0, len(xs)))
    x0, y0 = int(xs[j]) - 32, int(ys[j]) - 32
    g.map_update_patch(x0, y0, episode.extract_patch(lab, x0, y0, 65, 65))
    print("replan", g.plan(path_cap=1)[0], g.plan_info(), flush=True)
