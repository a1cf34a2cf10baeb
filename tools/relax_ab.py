"""A/B of the cost-to-go relaxation kernels: python tools/relax_ab.py W [kernel ...]   (kernel 0 = per-tile k_relax, 1 = block-resident
k_relax_blk). Every configuration runs in a process of its own (the switches are read once per process); the fields are compared
through a checksum + a full comparison against the first configuration's field saved to /tmp."""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scouting-ipp_b200"))


def child(W, tag):
    import numpy as np
    import sipp
    from sipp import synth
    K = {640: 20, 1024: 100, 2048: 400, 4096: 1500, 8192: 6000}.get(W, 200)
    H = 480 if W == 640 else W
    seed = {4096: 1003, 8192: 1004, 1024: 1002}.get(W, 1003)
    lab, obs, desc = synth.world(W, H, seed, K, sipp.make_desc, init_cost="max")
    ctx = sipp.Context(desc)
    ctx.set_incremental(False)
    ctx.map_upload_full(lab, obs)
    ts = []
    for i in range(7):
        t0 = time.perf_counter()
        r = ctx.plan(path_cap=1)
        ts.append((time.perf_counter() - t0) * 1e3)
    f = ctx.field_download()
    ref = "/tmp/relax_ab_ref_%d.npy" % W
    if os.path.exists(ref):
        same = bool(np.array_equal(np.load(ref), f))
    else:
        np.save(ref, f)
        same = None
    print("W %d %s: plan median %.3f ms (min %.3f)  cost %.6f nodes %d  stats[pushes, activations, sweeps, abort] %s  same_field=%s" % (
        W, tag, float(np.median(ts[1:])), min(ts[1:]), r[0], r[3], ctx.plan_stats(), same), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        child(int(sys.argv[2]), sys.argv[3])
        sys.exit(0)
    W = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    kernels = sys.argv[2:] or ["0", "1"]
    ref = "/tmp/relax_ab_ref_%d.npy" % W
    if os.path.exists(ref) and not os.environ.get("AB_KEEP_REF"):      # AB_KEEP_REF=1: compare with the field a previous invocation saved
        os.remove(ref)
    for k in kernels:
        env = dict(os.environ)
        for kv in k.split(","):          # "1" or "1,SIPP_RELAX_GRID=74"
            if "=" in kv:
                a, b = kv.split("=")
                env[a] = b
            else:
                env["SIPP_RELAX_KERNEL"] = kv
        env["SIPP_PLAN_TIMING"] = "1" if os.environ.get("AB_TIMING") else ""
        if not env["SIPP_PLAN_TIMING"]:
            del env["SIPP_PLAN_TIMING"]
        subprocess.run([sys.executable, os.path.abspath(__file__), "--child", str(W), "kernel=" + k], env=env, timeout=900)
