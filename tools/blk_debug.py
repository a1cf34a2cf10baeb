"""Debug aid for the block-resident relaxation kernel: plan small multi-block maps with both kernels in bounded subprocesses and
compare the fields. python tools/blk_debug.py [W ...]"""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scouting-ipp_b200"))


def child(W, H):
    import numpy as np
    import sipp
    from sipp import synth
    lab, obs, desc = synth.world(W, H, 1003, max(4, W // 64), sipp.make_desc, init_cost="max")
    ctx = sipp.Context(desc)
    ctx.set_incremental(False)
    ctx.map_upload_full(lab, obs)
    t0 = time.perf_counter()
    try:
        r = ctx.plan(path_cap=1)
    except Exception as e:
        print("  plan failed after %.2f s: %s  stats %s" % (time.perf_counter() - t0, e, ctx.plan_stats()), flush=True)
        return
    dt = time.perf_counter() - t0
    f = ctx.field_download()
    ref = "/tmp/blk_debug_ref_%d_%d.npy" % (W, H)
    same = None
    if os.path.exists(ref):
        g = np.load(ref)
        same = bool(np.array_equal(g, f))
        if not same:
            bad = np.argwhere(g != f)
            print("  field differs in %d cells, first %s: %r vs %r" % (len(bad), bad[0], f[tuple(bad[0])], g[tuple(bad[0])]), flush=True)
    else:
        np.save(ref, f)
    print("  %dx%d kernel=%s: first plan %.3f ms cost %.6f nodes %d stats %s same_field=%s" % (
        W, H, os.environ.get("SIPP_RELAX_KERNEL"), dt * 1e3, r[0], r[3], ctx.plan_stats(), same), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 3 and sys.argv[1] == "--child":
        child(int(sys.argv[2]), int(sys.argv[3]))
        sys.exit(0)
    sizes = [int(a) for a in sys.argv[1:]] or [128, 256, 640]
    for W in sizes:
        H = 480 if W == 640 else W
        ref = "/tmp/blk_debug_ref_%d_%d.npy" % (W, H)
        if os.path.exists(ref):
            os.remove(ref)
        for k in ("0", "1"):
            env = dict(os.environ, SIPP_RELAX_KERNEL=k)
            try:
                subprocess.run([sys.executable, os.path.abspath(__file__), "--child", str(W), str(H)], env=env, timeout=90)
            except subprocess.TimeoutExpired:
                print("  %dx%d kernel=%s: TIMEOUT (90 s)" % (W, H, k), flush=True)
