"""Short fixed workload for ncu on the plan kernels: python tools/prof_plan.py [W] [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scouting-ipp_b200"))
import sipp
from sipp import synth
W = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
K = {1024: 100, 2048: 400, 4096: 1500, 8192: 6000}.get(W, 200)
lab, obs, desc = synth.world(W, W, 1003, K, sipp.make_desc, init_cost="max")
ctx = sipp.Context(desc)
ctx.map_upload_full(lab, obs)
for _ in range(reps):
    print("plan", ctx.plan()[0], ctx.plan_stats(), flush=True)
