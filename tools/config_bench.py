"""One-GPU measurements of BASELINE.json's configs 2-4 with the inputs of SURVEY.md 8(d) (bench.py covers the headline workload).
Prints one JSON object; CUDA events on the launching stream, L2 flushed between timed launches.

    python tools/config_bench.py [2] [3] [4]
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scouting-ipp_b200"))
import numpy as np, torch
import sipp
from sipp import synth

dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
st = torch.cuda.Stream(device=dev)
ALG = 4257


def timed_us(fn, n=20, warm=5):
    with torch.cuda.stream(st):
        for _ in range(warm):
            fn()
        out = []
        for k in range(n):
            flush.fill_(k)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st); fn(); b.record(st)
            st.synchronize()
            out.append(a.elapsed_time(b) * 1e3)
    return float(np.median(out)), float(min(out))


def eval_case(ctx, W, seed, T, npg, rep=1):
    xy, cost = synth.make_candidates(W, W, seed, T)
    xy, cost = np.tile(xy, (rep, 1)), np.tile(cost, rep)
    n = T * rep
    d = dict(xy=torch.from_numpy(xy).to(dev), cost=torch.from_numpy(cost).to(dev), gain=torch.empty(n, dtype=torch.float64, device=dev),
             val=torch.empty(n, dtype=torch.float64, device=dev), term=torch.empty(n, dtype=torch.uint8, device=dev),
             best=torch.zeros(2, dtype=torch.int64, device=dev))
    p = sipp.make_params(sipp.EVAL_RRT_STAR, half_fov=32, non_path_voxel_gain=npg)
    med, mn = timed_us(lambda: ctx.cycle_dev(p, n, d["xy"].data_ptr(), None, d["cost"].data_ptr(), d["gain"].data_ptr(), d["term"].data_ptr(),
                                             d["val"].data_ptr(), d["best"].data_ptr(), st.cuda_stream))
    return {"candidates": n, "us_per_cycle_median": med, "us_min": mn, "evals_per_s": n / (med * 1e-6), "algorithmic_GBps": ALG * n / (med * 1e-6) / 1e9}


def plan_case(ctx, reps=5):
    ctx.set_incremental(False)      # from scratch
    ctx.plan()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); r = ctx.plan(path_cap=1); ts.append((time.perf_counter() - t0) * 1e3)
    return {"plan_ms_median": float(np.median(ts)), "plan_ms_min": min(ts), "cost": r[0], "status": r[1], "path_nodes": r[3], "stats": ctx.plan_stats()}


which = [int(a) for a in sys.argv[1:]] or [2, 3, 4]
res = {}
if 2 in which:      # 4096 poses on a 1024^2 label map: latency-bound; throughput with the batch replicated x256
    W = 1024
    lab, obs, desc = synth.world(W, W, 1002, 100, sipp.make_desc, init_cost="max")
    ctx = sipp.Context(desc); ctx.map_upload_full(lab, obs); ctx.plan()
    res["config2_1024_4096poses"] = {"npg_0.001": eval_case(ctx, W, 1002, 4096, 0.001), "npg_0": eval_case(ctx, W, 1002, 4096, 0.0),
                                    "x256_npg_0.001": eval_case(ctx, W, 1002, 4096, 0.001, rep=256)}
    ctx.close()
if 3 in which:      # cost-to-go start -> goal on the 4096^2 map: (a) fully known (eval-planner mode), (b) partial, init_cost min
    W = 4096
    lab = synth.make_labels(W, W, 1003); obs = synth.make_observed(W, W, 1003, 1500)
    ids = sorted(set(np.unique(lab).tolist()) - {0})
    geo = (W, W, W * synth.MPP, W * synth.MPP, (0.5, 0.5), (W * synth.MPP - 0.5, W * synth.MPP - 0.5))
    ctx = sipp.Context(sipp.make_desc(*geo, min(ids), max(ids), min(ids), obstacle_ids_rov=(0, 3599)))
    ctx.map_upload_full(lab); a = plan_case(ctx); ctx.close()
    ctx = sipp.Context(sipp.make_desc(*geo, min(ids), max(ids), min(ids)))
    ctx.map_upload_full(lab, obs); b = plan_case(ctx); ctx.close()
    res["config3_4096_plan"] = {"a_fully_known": a, "b_partial_init_min": b, "algorithmic_bytes": 10 * W * W}
if 4 in which:      # full planning cycle, 65536 candidates on the 8192^2 map
    W = 8192
    lab, obs, desc = synth.world(W, W, 1004, 6000, sipp.make_desc, init_cost="max")
    ctx = sipp.Context(desc); ctx.map_upload_full(lab, obs)
    pl = plan_case(ctx, reps=3)
    ev = eval_case(ctx, W, 1004, 65536, 0.001)
    res["config4_8192_full_cycle"] = {"plan": pl, "eval": ev, "ms_per_full_cycle": pl["plan_ms_median"] + ev["us_per_cycle_median"] * 1e-3}
    ctx.close()
print(json.dumps(res), flush=True)
