"""Quick GPU timing probe (development aid, not the benchmark): python tools/gpu_probe.py [W] [T]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scouting-ipp_b200"))
import numpy as np
import torch
import sipp
from sipp import synth

W = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
K = {1024: 100, 4096: 1500, 8192: 6000}.get(W, 200)
t0 = time.time()
lab, obs, desc = synth.world(W, W, 1003, K, sipp.make_desc, init_cost="max")
print("synth %.1fs observed frac %.3f" % (time.time() - t0, obs.mean()), flush=True)
ctx = sipp.Context(desc)
t0 = time.time(); ctx.map_upload_full(lab, obs); print("upload %.3fs" % (time.time() - t0), flush=True)
for rep in range(3):
    t0 = time.time(); cost, status, path, n = ctx.plan(); dt = time.time() - t0
    print("plan %.2f ms cost %.3f status %d len %d stats %s" % (dt * 1e3, cost, status, n, ctx.plan_stats()), flush=True)
xy, cst = synth.make_candidates(W, W, 1003, T)
dev = torch.device("cuda:0")
d_xy = torch.from_numpy(xy).to(dev); d_cost = torch.from_numpy(cst).to(dev)
d_gain = torch.empty(T, dtype=torch.float64, device=dev); d_val = torch.empty(T, dtype=torch.float64, device=dev)
d_cnt = torch.empty((T, 4), dtype=torch.int32, device=dev); d_term = torch.empty(T, dtype=torch.uint8, device=dev)
d_best = torch.empty(2, dtype=torch.int64, device=dev)
stream = torch.cuda.Stream()
for name, kw in [("rrt_count", dict(evaluator=sipp.EVAL_RRT_STAR, non_path_voxel_gain=0.001)),
                 ("goal_dist", dict(evaluator=sipp.EVAL_GOAL_DISTANCE)),
                 ("avg_view", dict(evaluator=sipp.EVAL_AVERAGE_VIEW_COST)),
                 ("rrt_discount", dict(evaluator=sipp.EVAL_RRT_STAR, use_distance_discount=True))]:
    p = sipp.make_params(**kw)
    with torch.cuda.stream(stream):
        for _ in range(3):
            ctx.cycle_dev(p, T, d_xy.data_ptr(), None, d_cost.data_ptr(), d_gain.data_ptr(), d_term.data_ptr(), d_val.data_ptr(), d_best.data_ptr(), stream.cuda_stream)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        R = 20
        for _ in range(R):
            ctx.cycle_dev(p, T, d_xy.data_ptr(), None, d_cost.data_ptr(), d_gain.data_ptr(), d_term.data_ptr(), d_val.data_ptr(), d_best.data_ptr(), stream.cuda_stream)
        e1.record(stream); stream.synchronize()
    ms = e0.elapsed_time(e1) / R
    print("%-12s %.3f ms/cycle  %.1f M evals/s  alg %.1f GB/s  best %s" % (name, ms, T / ms / 1e3, T * 4257 / ms / 1e6, d_best.tolist()), flush=True)
t0 = time.time(); g = ctx.cycle(sipp.make_params(sipp.EVAL_RRT_STAR, non_path_voxel_gain=0.001), xy, cst); print("host cycle %.3f ms best %d" % ((time.time() - t0) * 1e3, g[3]))
