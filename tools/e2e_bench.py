"""End-to-end timing of sipp_cycle (host pointers, pinned) for different chunk counts: python tools/e2e_bench.py [chunks ...]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scouting-ipp_b200"))
import numpy as np, torch
import sipp
from sipp import synth
W, T = 4096, 65536
lab, obs, desc = synth.world(W, W, 1003, 1500, sipp.make_desc, init_cost="max")
xy, cost = synth.make_candidates(W, W, 1003, T)
h_xy = torch.from_numpy(xy).pin_memory(); h_cost = torch.from_numpy(cost).pin_memory()
h_gain = torch.empty(T, dtype=torch.float64).pin_memory(); h_value = torch.empty(T, dtype=torch.float64).pin_memory()
h_term = torch.empty(T, dtype=torch.uint8).pin_memory()
p = sipp.make_params(sipp.EVAL_RRT_STAR, non_path_voxel_gain=0.001)
ref = None
for ch in (sys.argv[1:] or ["1", "2", "4", "8"]):
    os.environ["SIPP_CYCLE_CHUNKS"] = ch
    ctx = sipp.Context(desc)
    ctx.map_upload_full(lab, obs); ctx.plan()
    for _ in range(5): r = ctx.cycle_host_ptr(p, T, h_xy.data_ptr(), None, h_cost.data_ptr(), h_gain.data_ptr(), h_term.data_ptr(), h_value.data_ptr())
    t0 = time.perf_counter()
    for _ in range(50): r = ctx.cycle_host_ptr(p, T, h_xy.data_ptr(), None, h_cost.data_ptr(), h_gain.data_ptr(), h_term.data_ptr(), h_value.data_ptr())
    dt = (time.perf_counter() - t0) / 50
    g = h_gain.numpy().copy()
    if ref is None: ref = (g, r)
    print("chunks %s: %.1f us/step  %.1f M evals/s  same=%s" % (ch, dt * 1e6, T / dt / 1e6, np.array_equal(g, ref[0]) and r == ref[1]), flush=True)
    del ctx
