"""Tuning sweep for k_gain's pipeline shape (warps per CTA x tile slots per warp). Usage: python tools/gain_sweep.py [W] [T]
Each configuration runs in a fresh context (the knobs are read at sipp_create); kernel time by CUDA events, L2 flushed."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scouting-ipp_b200"))
import numpy as np
import torch
import sipp
from sipp import synth
W = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
K = {1024: 100, 2048: 400, 4096: 1500, 8192: 6000}.get(W, 200)
lab, obs, desc = synth.world(W, W, 1003, K, sipp.make_desc, init_cost="max")
xy, cst = synth.make_candidates(W, W, 1003, T)
dev = torch.device("cuda:0")
d_xy = torch.from_numpy(xy).to(dev); d_cost = torch.from_numpy(cst).to(dev)
d_gain = torch.empty(T, dtype=torch.float64, device=dev); d_val = torch.empty(T, dtype=torch.float64, device=dev)
d_term = torch.empty(T, dtype=torch.uint8, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
p = sipp.make_params(sipp.EVAL_RRT_STAR, non_path_voxel_gain=0.001)
mask = (np.random.default_rng(1).random((W, W)) < 0.001).astype(np.uint8)
ref = None
configs = [(16, 2), (14, 3), (13, 3), (10, 4), (8, 5), (7, 6), (6, 7), (5, 8), (12, 3), (8, 4), (16, 1)]
if len(sys.argv) > 3:
    configs = [tuple(int(v) for v in c.split("x")) for c in sys.argv[3:]]
for warps, stages in configs:
    os.environ["SIPP_GAIN_WARPS"] = str(warps); os.environ["SIPP_GAIN_STAGES"] = str(stages)
    ctx = sipp.Context(desc)
    ctx.map_upload_full(lab, obs)
    ctx.set_path_mask(mask)
    s = torch.cuda.Stream()
    times = []
    with torch.cuda.stream(s):
        for it in range(34):
            flush.fill_(it)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(s)
            ctx.gain_batch_dev(p, T, d_xy.data_ptr(), None, d_gain.data_ptr(), None, d_term.data_ptr(), s.cuda_stream)
            e1.record(s)
            s.synchronize()
            times.append(e0.elapsed_time(e1))
    g = d_gain.cpu().numpy()
    if ref is None: ref = g
    t = float(np.mean(times[4:])) * 1e-3   # event timestamps tick in ~4 us steps on this part: average, don't take the median
    print("warps %2d stages %d: %7.2f us  %6.1f GB/s algorithmic  frac %.3f  same=%s" % (warps, stages, t * 1e6, 4257 * T / t / 1e9, 4257 * T / t / 1e9 / 6549.4, np.array_equal(g, ref)), flush=True)
    del ctx
