"""Short fixed workload for ncu: python tools/prof_run.py [W] [T] [reps]  (one plan + reps evaluation cycles)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scouting-ipp_b200"))
import numpy as np
import torch
import sipp
from sipp import synth
W = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
K = {1024: 100, 2048: 400, 4096: 1500, 8192: 6000}.get(W, 200)
lab, obs, desc = synth.world(W, W, 1003, K, sipp.make_desc, init_cost="max")
ctx = sipp.Context(desc)
ctx.map_upload_full(lab, obs)
print("plan", ctx.plan()[0], ctx.plan_stats(), flush=True)
xy, cst = synth.make_candidates(W, W, 1003, T)
dev = torch.device("cuda:0")
d_xy = torch.from_numpy(xy).to(dev); d_cost = torch.from_numpy(cst).to(dev)
d_gain = torch.empty(T, dtype=torch.float64, device=dev); d_val = torch.empty(T, dtype=torch.float64, device=dev)
d_term = torch.empty(T, dtype=torch.uint8, device=dev); d_best = torch.empty(2, dtype=torch.int64, device=dev)
p = sipp.make_params(sipp.EVAL_RRT_STAR, non_path_voxel_gain=0.001)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(reps):
        ctx.cycle_dev(p, T, d_xy.data_ptr(), None, d_cost.data_ptr(), d_gain.data_ptr(), d_term.data_ptr(), d_val.data_ptr(), d_best.data_ptr(), s.cuda_stream)
    s.synchronize()
print("best", d_best.tolist(), flush=True)
