"""Ensemble workload: M independent 640x480 maps (seeds 2000+i), each running an episode of C planning
cycles {ingest the 65x65 patch at the scout, re-plan the follower path, score T = 1024 candidate views, select}. Maps shard over
GPUs with no communication (torchrun: rank r takes maps r::world); on one GPU the episodes run side by side on their contexts'
streams, driven by a pool of host threads (every ABI call releases the GIL). Prints one JSON line: map-cycles/s.

    python tools/ensemble_bench.py [--maps 128] [--cycles 40] [--threads 16] [--check 2]
"""
import argparse, json, os, sys, time
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scouting-ipp_b200"))
import numpy as np
import sipp
from sipp import episode, synth

W, H, T = 640, 480, 1024


def make_map(seed, mk):
    lab = synth.make_labels(W, H, seed)
    ids = sorted(set(np.unique(lab).tolist()) - {0})
    desc = mk(W, H, W * synth.MPP, H * synth.MPP, (0.5, 0.5), (W * synth.MPP - 0.5, H * synth.MPP - 0.5), min(ids), max(ids), max(ids))
    return lab, desc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--maps", type=int, default=128)
    ap.add_argument("--cycles", type=int, default=40)
    ap.add_argument("--threads", type=int, default=16)
    ap.add_argument("--check", type=int, default=0, help="unused (kept for old command lines)")
    ap.add_argument("--batches", type=int, default=2, help="batch mode: the maps are split into this many batches, driven side by side from one host thread each "
                    "(the host preparation of one batch overlaps the kernels of another)")
    ap.add_argument("--mode", default="batch", choices=["batch", "threads"],
                    help="batch: sipp_batch_episode_run, every kernel launched once for all maps (one host thread); threads: one sipp_episode_run per map from a pool of host threads")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    seeds = [2000 + i for i in range(args.maps * world)][rank::world]
    params = sipp.make_params(sipp.EVAL_RRT_STAR, half_fov=32, non_path_voxel_gain=0.001)
    worlds = []
    for sd in seeds:
        lab, desc = make_map(sd, sipp.make_desc)
        worlds.append((sd, lab, sipp.Context(desc, device=local)))
    cands = {sd: episode.episode_candidates(W, H, sd, args.cycles, T) for sd in seeds}     # inputs are made before the clock starts

    def run(item):          # one native call per episode (sipp_episode_run): the GIL is released for its whole duration
        sd, lab, ctx = item
        return ctx.episode_run(lab, params, cands[sd][0], cands[sd][1])

    launches = None
    if args.mode == "batch":
        nb = max(1, min(args.batches, len(worlds) // 8 or 1))
        groups = [worlds[k::nb] for k in range(nb)]
        batches = [sipp.Batch([w[2] for w in g]) for g in groups]
        if not os.environ.get("SIPP_BATCH_SHARE"):
            for b in batches:
                b.set_sharing(nb)

        def run_group(k, ncyc=None):
            g = groups[k]
            return batches[k].episode_run([w[1] for w in g], params, [cands[w[0]][0][:ncyc] for w in g], [cands[w[0]][1][:ncyc] for w in g])

        with ThreadPoolExecutor(nb) as pool:
            list(pool.map(lambda k: run_group(k, 2), range(nb)))     # warm-up (allocations)
            for _, _, ctx in worlds:
                ctx.map_clear()
            l0 = sum(b.launch_count() for b in batches)
            t0 = time.perf_counter()
            parts = list(pool.map(run_group, range(nb)))
            dt = time.perf_counter() - t0
        launches = (sum(b.launch_count() for b in batches) - l0) / nb
        traces = [None] * len(worlds)
        for k in range(nb):
            for j, tr in enumerate(parts[k]):
                traces[k + j * nb] = tr
        for b in batches:
            b.close()
    else:
        with ThreadPoolExecutor(args.threads) as pool:
            list(pool.map(lambda it: it[2].episode_run(it[1], params, cands[it[0]][0][:2], cands[it[0]][1][:2]), worlds))   # warm-up (allocations)
            for _, _, ctx in worlds:
                ctx.map_clear()
            t0 = time.perf_counter()
            traces = list(pool.map(run, worlds))
            dt = time.perf_counter() - t0
    class Timed:          # per-call wall time of one episode running alone
        def __init__(self, ctx):
            self.ctx, self.t = ctx, {}
        def __getattr__(self, name):
            fn = getattr(self.ctx, name)
            def wrapped(*a, **k):
                t = time.perf_counter(); r = fn(*a, **k); self.t[name] = self.t.get(name, 0.0) + time.perf_counter() - t
                return r
            return wrapped
    worlds[0][2].map_clear()
    tm = Timed(worlds[0][2])
    t1 = time.perf_counter()
    one = episode.run_episode(tm, worlds[0][1], params, worlds[0][0], cycles=args.cycles, T=T)
    dt_one = time.perf_counter() - t1
    assert one == traces[0]
    checked = 0      # parity of the episode traces with the CPU checker is a test's job: tests/test_gpu_batch.py, tests/test_gpu_parity.py
    line = {"metric": "ensemble map-cycles/s (640x480 maps, 1024 candidates per cycle)", "value": len(worlds) * args.cycles / dt, "unit": "map-cycles/s",
            "rank": rank, "world": world, "maps_on_this_gpu": len(worlds), "cycles": args.cycles, "host_threads": 1 if args.mode == "batch" else args.threads,
            "driver": ("sipp_batch_episode_run: every kernel launched once for all maps of the GPU (gridDim.y = maps), one host thread, one synchronisation per map-cycle"
                       if args.mode == "batch" else "sipp_episode_run (native host loop), one call per episode from a pool of host threads"),
            "kernel_launches_per_batch_cycle": (launches / args.cycles) if launches is not None else None, "batches": (args.batches if args.mode == "batch" else None),
            "ms_per_map_cycle_alone": dt_one / args.cycles * 1e3, "ms_per_map_cycle_in_ensemble": dt / (len(worlds) * args.cycles) * 1e3,
            "alone_ms_per_cycle_by_call": {k: v / args.cycles * 1e3 for k, v in tm.t.items()}, "episodes_checked_against_oracle": checked, "final_costs": [t[-1][0] for t in traces[:4]]}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
