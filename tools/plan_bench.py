"""Plan-time tuning: python tools/plan_bench.py W [alpha:window ...]   (cost-to-go field + path + mask through sipp_plan)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scouting-ipp_b200"))
import numpy as np
import sipp
from sipp import synth
W = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
K = {1024: 100, 2048: 400, 4096: 1500, 8192: 6000}.get(W, 200)
lab, obs, desc = synth.world(W, W, 1003, K, sipp.make_desc, init_cost="max")
cfgs = sys.argv[2:] or ["1:4096"]
ref = None
for cfg in cfgs:
    alpha, window = cfg.split(":")
    os.environ["SIPP_RELAX_ALPHA"] = alpha; os.environ["SIPP_RELAX_WINDOW"] = window
    ctx = sipp.Context(desc)
    ctx.set_incremental(False)      # every plan below starts from scratch
    ctx.map_upload_full(lab, obs)
    ts = []
    for i in range(6):
        t0 = time.perf_counter()
        try:
            r = ctx.plan()
        except Exception as e:
            print("plan %d failed: %s stats %s" % (i, e, ctx.plan_stats()), flush=True)
            break
        ts.append((time.perf_counter() - t0) * 1e3)
        if os.environ.get("PLAN_VERBOSE"): print("  plan %d: %.2f ms %s" % (i, ts[-1], ctx.plan_stats()), flush=True)
    if len(ts) < 2: continue
    f = ctx.field_download()
    if ref is None: ref = f
    print("alpha %s window %s: plan %.2f ms (min %.2f)  cost %.6f nodes %d stats %s same_field=%s" % (
        alpha, window, float(np.median(ts[1:])), min(ts[1:]), r[0], r[3], ctx.plan_stats(), np.array_equal(f, ref)), flush=True)
    del ctx
