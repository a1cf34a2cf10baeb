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
            print("    zeros new %s ref %s" % (np.argwhere(f == 0).tolist()[:8], np.argwhere(g == 0).tolist()[:8]), flush=True)
            lows = np.argwhere(f < g)
            # the lowest-valued wrong cell: where the bad information entered
            k = lows[np.argmin(f[lows[:, 0], lows[:, 1]])]
            y, x = int(k[0]), int(k[1])
            print("    smallest wrong value at (row %d, col %d): new %r ref %r; neighbourhood new:" % (y, x, f[y, x], g[y, x]), flush=True)
            with np.printoptions(precision=3, linewidth=250, suppress=True):
                print(f[max(0, y - 3):y + 4, max(0, x - 3):x + 4], flush=True)
                print("    ref:", flush=True)
                print(g[max(0, y - 3):y + 4, max(0, x - 3):x + 4], flush=True)
            for by in range(0, H, 128):
                for bx in range(0, W, 128):
                    fb, gb = f[by:by + 128, bx:bx + 128], g[by:by + 128, bx:bx + 128]
                    with np.errstate(invalid="ignore"):
                        lo, hi = int(np.sum(fb < gb)), int(np.sum(fb > gb))
                    if lo or hi:
                        w = np.argwhere(fb != gb)
                        print("    block (%d,%d): %d lower %d higher; rows %d..%d cols %d..%d; inf new %d ref %d" % (
                            bx // 128, by // 128, lo, hi, w[:, 0].min(), w[:, 0].max(), w[:, 1].min(), w[:, 1].max(),
                            int(np.isinf(fb).sum()), int(np.isinf(gb).sum())), flush=True)
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
