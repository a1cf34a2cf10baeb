#!/usr/bin/env python
"""Benchmark of the scouting-ipp evaluation hot path on B200 (libsipp.so) — see DESIGN.md "Measurement".

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl graft|reference]

One "step" = one pass of the per-cycle evaluation over one candidate batch: footprint gain for every candidate
(k_gain, TMA-tiled window scan) + value = gain/cost + argmax (fused into the kernel's tail); with N > 1 the candidate set is sharded
(weak scaling: 65536 candidates per GPU, map replicated) and each step ends with the exchange of the 16-byte shard
winners through peer memory over NVLink, fused into the same kernel (an NCCL all_gather form is timed beside it). Metric: candidate-view evals/s on the 4096^2 map (BASELINE.json). The follower
cost-to-go / path / mask step runs once per planning cycle; it is timed separately and reported under "cycle".

--impl reference times the reference's own CPU algorithm (the oracle port of the C++ evaluators, faithful
variant incl. the quadratic erase loop) with all host threads on a bounded sample of the same workload.
"""
import argparse
import json
import os
import sys
import threading
import time

# The end-to-end leg is a tight host loop around the CUDA driver. torchrun (N > 1) exports OMP_NUM_THREADS=1; a plain `python bench.py`
# (N = 1) would start numpy's / torch's OpenMP pools with one spinning worker per vCPU, which on the 16-vCPU GPU VMs competes with the
# thread that feeds the GPU (measured: 175-192 us per call at N = 1 against 113-126 us for the same call under torchrun). Nothing in
# this benchmark uses those pools (the CPU baseline runs its own std::threads), so both launch modes get the same setting.
os.environ.setdefault("OMP_NUM_THREADS", "1")
os.environ.setdefault("MKL_NUM_THREADS", "1")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scouting-ipp_b200"))

import numpy as np

os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL's version banner / warnings go to stderr: stdout carries exactly one JSON line

W = H = 4096
SEED = 1003
K_OBS = 1500
T_PER_GPU = 65536
SCALING = "weak"
if os.environ.get("SIPP_BENCH_WORKLOAD") == "config4":
    # BASELINE config 4 (not the default line): full planning cycle with 65536 candidates IN TOTAL on the 8192^2 map at 1/2/4/8 GPUs —
    # strong scaling of the evaluation (the candidates are split), the follower plan is replicated on every GPU (one dependent front)
    W = H = 8192
    SEED = 1004
    K_OBS = 6000
    T_PER_GPU = 65536 // max(1, int(os.environ.get("WORLD_SIZE", "1")))
    SCALING = "strong"
HALF_FOV = 32
NPG = 0.001
ALG_BYTES_PER_EVAL = 4257   # (2*32+1)^2 one-byte cell records + 16 B pose in + 16 B result out
WORKLOAD = ("planning-cycle evaluation: %d candidate views per GPU (half_fov %d = 4226 voxels each), RRTStarEvaluator count mode "
            "(non_path_voxel_gain %g) + GlobalNormalizedGain value + argmax, %dx%d synthetic map (create_artificial_maps.py style, "
            "seed %d, %d observed footprints, unknown_init_cost max)" % (T_PER_GPU, HALF_FOV, NPG, W, H, SEED, K_OBS))


def bench_config(n_gpus):
    """the same dictionary in both arms (the driver compares them key by key)"""
    return {"workload": WORKLOAD, "candidates_per_gpu": T_PER_GPU, "map": "%dx%d" % (W, H), "half_fov": HALF_FOV,
            "evaluator": "RRTStarEvaluator (count mode, non_path_voxel_gain %g)" % NPG, "n_gpus": n_gpus,
            "l2": "flushed between timed steps (256 MiB write, untimed); the scanned planes (4 + 16 MiB) are smaller than L2",
            "sharding": ("single GPU" if n_gpus == 1 else "candidates sharded over %d GPUs (%s scaling: %d each), map replicated per GPU; the only "
                         "cross-GPU traffic is each shard's 16-byte winner" % (n_gpus, SCALING, T_PER_GPU))}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons with NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)       # the timed region of the default run is ~20 ms: a sample every few ms sees the GPU under load

    def result(self):
        self.stop_flag = True
        self.join(timeout=1.0)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def bind_to_gpu_numa_node(index):
    """Host placement: run this process (and allocate its pinned buffers) on the NUMA node the GPU hangs off. The end-to-end path
    is PCIe bound; from the far socket every copy crosses the inter-socket link. Returns a note for the JSON line."""
    if os.environ.get("SIPP_BENCH_NO_NUMA"):
        return "not bound (SIPP_BENCH_NO_NUMA)"
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]                      # NVML prints an 8-digit domain, sysfs a 4-digit one
        base = "/sys/bus/pci/devices/%s" % bus
        node = int(open(base + "/numa_node").read())
        if node < 0:
            return "numa node unknown"
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return "numa node %d has no allowed cpu" % node
        os.sched_setaffinity(0, cpus)
        return "bound to numa node %d (%d cpus)" % (node, len(cpus))
    except Exception as exc:
        return "not bound (%s)" % type(exc).__name__


def build_inputs(n_gpus):
    from sipp import synth
    lab = synth.make_labels(W, H, SEED)
    obs = synth.make_observed(W, H, SEED, K_OBS)
    xy, cost = synth.make_candidates(W, H, SEED, T_PER_GPU * n_gpus)
    ids = sorted(set(np.unique(lab).tolist()) - {0})
    return lab, obs, xy, cost, ids


def cpu_reference(lab, obs, ids):
    """The reference's CPU implementation of the path for the bench map: the reference's OWN evaluator sources (oracle/_ref, built from
    /root/reference in the build container and shipped prebuilt) over the oracle's map state when that library is here, else the
    oracle's restatement. Returns (cycle(params, xy, cost, threads) -> (gain, terminal, value, best_index, best_value), kind, note,
    oracle context, seconds the oracle's follower plan took)."""
    from oracle import oracle as orc
    desc = orc.make_desc(W, H, W * 0.0625, H * 0.0625, (0.5, 0.5), (W * 0.0625 - 0.5, H * 0.0625 - 0.5), min(ids), max(ids), max(ids))
    o = orc.Oracle(desc)
    o.map_upload_full(lab, obs)
    t0 = time.perf_counter()
    o.plan()
    plan_s = time.perf_counter() - t0
    try:
        from oracle import ref
        if ref.available():
            r = ref.Reference(desc).snapshot(o)
            return (lambda p, xy, cost, threads: r.cycle(p, xy, cost, threads=threads)), "reference", (
                "the reference's own sources (advanced_planner_modules: MapPlanExp::getVoxelArea, CameraModelPlanExp, RRTStarEvaluator incl. its "
                "quadratic erase loop) compiled unmodified into oracle/_ref against stand-in headers; map planes from the oracle's map server "
                "restatement"), o, plan_s
    except Exception:
        pass
    return (lambda p, xy, cost, threads: o.cycle(p, xy, cost, variant=0, threads=threads)), "port", (
        "oracle port of the reference evaluator path (voxel vector + quadratic erase loop); oracle/_ref is not available here"), o, plan_s


def run_reference(args):
    """Reference arm: the reference's CPU evaluators on the box's host cores, all threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle as orc
    lab, obs, xy, cost, ids = build_inputs(1)
    cycle, kind, note, o, plan_s = cpu_reference(lab, obs, ids)
    p = orc.make_params(orc.EVAL_RRT_STAR, half_fov=HALF_FOV, non_path_voxel_gain=NPG)
    cores = os.cpu_count() or 1
    # bounded sample per step, sized from a probe so the whole run stays within a couple of minutes
    probe = 64 * cores
    t0 = time.perf_counter()
    cycle(p, xy[:probe], cost[:probe], cores)
    rate = probe / (time.perf_counter() - t0)
    budget_s = 90.0 / max(1, args.steps + args.warmup)
    sample = int(min(T_PER_GPU, max(8 * cores, rate * budget_s)))
    for _ in range(args.warmup):
        cycle(p, xy[:sample], cost[:sample], cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cycle(p, xy[:sample], cost[:sample], cores)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": "candidate-view evals/sec on 4096^2 cost map", "value": value, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": SCALING, "vs_baseline": None, "dtype": "u8 counts + f64 gain", "data": "synthetic",
        "config": bench_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": cores, "kind": kind,
                         "sample": "%d of the %d candidates per step, %d worker threads (the reference itself is single-threaded: one module "
                                   "instance per thread); %s" % (sample, T_PER_GPU, cores, note),
                         "plan_s": plan_s},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def config1_cycle():
    """BASELINE config 1 — example.launch defaults (launch/example.launch:33,38-39) on the reference's own map_212: GoalDistanceEvaluator
    (max_gain 100, half_fov 32), prm_star follower, one planning cycle = ingest of a 65 x 65 patch + follower plan + the updater pass over a
    1024-node tree (what updateSegment(root) does) + next-best selection, timed through the ABI; beside it the CPU checker's literal
    emulation of the reference pass on one host core (the reference's planner loop is single-threaded)."""
    import sipp
    from sipp import synth
    res = {}
    # map ingest of the observed region, follower plan, updater pass over a 1024-node tree (what updateSegment(root) does), selection;
    # the CPU datapoint is the oracle's literal emulation of the reference pass on one host core (the reference is single-threaded)
    from sipp import episode
    from oracle import oracle as orc
    maps = np.load(os.path.join(ROOT, "tests", "golden", "ref_maps.npz"))
    cfg = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_maps.json")))["map_212"]
    lab = maps["map_212"].astype(np.int32)
    W, H = cfg["W"], cfg["H"]
    ids = [i for i in cfg["present_ids"] if i not in cfg["obstacle_ids_rov"]]
    mk = lambda f: f(W, H, cfg["width"], cfg["height"], cfg["start"], cfg["goal"], min(ids), max(ids), min(ids), obstacle_ids_mav=cfg["obstacle_ids_mav"],
                     obstacle_ids_rov=cfg["obstacle_ids_rov"], unknown_id=cfg["unknown_id"])
    g, o = sipp.Context(mk(sipp.make_desc)), orc.Oracle(mk(orc.make_desc))
    obs = synth.make_observed(W, H, 1001, 40)
    ys, xs = np.nonzero(obs)
    rng = np.random.default_rng(1001)
    patches = [(0, 0, 320)] + [(int(xs[j]) - 32, int(ys[j]) - 32, 65) for j in rng.integers(0, len(xs), 60)]
    for x0, y0, sz in patches:
        pt = episode.extract_patch(lab, x0, y0, sz, sz)
        g.map_update_patch(x0, y0, pt); o.map_update_patch(x0, y0, pt)
    T = 1024
    parent, end, start, cost = synth.make_tree(W, H, 1001, T)
    kw = dict(evaluator=sipp.EVAL_GOAL_DISTANCE, max_gain=100.0)
    pg, po = sipp.make_params(half_fov=32, **kw), orc.make_params(half_fov=32, **kw)
    ug = sipp.make_updater(sipp.UPD_RRT_STAR, sipp.FOLLOW_UPDATE_ALL, -1e9, 1e9, True, False, True)
    uo = orc.make_updater(0, 0, -1e9, 1e9, True, False, True)
    z, zt = np.zeros(T), np.zeros(T, np.uint8)

    def gpu_cycle():
        x0, y0, sz = patches[-1]
        g.map_update_patch(x0, y0, episode.extract_patch(lab, x0, y0, sz, sz))
        r = g.plan(path_cap=1)
        gg, gv, gt, _ = g.update_tree(pg, ug, parent, end, z, cost, z, zt, (0.5, 0.5), True, start_xy=start)
        return r, g.value_select(gg, cost, parent)

    def cpu_cycle():
        x0, y0, sz = patches[-1]
        o.map_update_patch(x0, y0, episode.extract_patch(lab, x0, y0, sz, sz))
        r = o.plan(path_cap=1)
        og, ov, ot = o.update_tree(po, uo, parent, end, z, cost, z, zt, (0.5, 0.5), True, start_xy=start)
        return r, orc.value_select(og, cost, parent)

    def wall(fn, reps):
        out = []
        for _ in range(reps):
            t0 = time.perf_counter(); r = fn(); out.append((time.perf_counter() - t0) * 1e3)
        return float(np.median(out)), min(out), r

    gpu_cycle(); g_med, g_min, rg = wall(gpu_cycle, 15)
    c_med, c_min, rc = wall(cpu_cycle, 3)
    tg = wall(lambda: g.update_tree(pg, ug, parent, end, z, cost, z, zt, (0.5, 0.5), True, start_xy=start), 15)
    tc = wall(lambda: o.update_tree(po, uo, parent, end, z, cost, z, zt, (0.5, 0.5), True, start_xy=start), 3)
    res["config1_example_launch_map_212"] = {
        "tree_nodes": T, "gpu_cycle_ms_median": g_med, "gpu_cycle_ms_min": g_min, "cpu_cycle_ms_median": c_med, "cpu_cores": 1, "cycle_speedup": c_med / g_med,
        "gpu_update_tree_ms": tg[0], "cpu_update_tree_ms": tc[0], "same_plan": bool(rg[0][0] == rc[0][0] and rg[0][3] == rc[0][3]),
        "same_selected_child": bool(rg[1][1] == rc[1][1]),
        "note": "cycle = ingest of one 65x65 patch + follower plan (incremental on the GPU; the reference re-plans from scratch) + RRTStarUpdater/UpdateAll pass "
                "over the tree + SubsequentBest selection, host-timed through the ABI with pageable host arrays; cpu = oracle emulation of the reference pass, 1 core"}
    g.close()
    return res["config1_example_launch_map_212"]


def ensemble_workload(args):
    """SIPP_BENCH_WORKLOAD=ensemble / --workload ensemble (BASELINE config 5, not the default line): 128 independent 640x480 maps per GPU,
    each running an episode of planning cycles {ingest the 65x65 patch at the scout, re-plan the follower path incrementally, score 1024
    candidate views, select, move}; the maps advance together through sipp_batch_episode_run (every kernel launched once per batch,
    one relaxation work queue for the batch), two batches side by side per GPU (SIPP_BENCH_ENSEMBLE_BATCHES). Maps shard over the GPUs with no data-path
    communication (weak scaling). A step = one planning cycle of every map; host wall clock between device synchronisations, max over
    ranks — the loop is host-driven (one synchronisation per batch-cycle), so this IS the end-to-end number."""
    from concurrent.futures import ThreadPoolExecutor
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    import sipp
    from sipp import episode, synth
    rank, local_rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        raise SystemExit("WORLD_SIZE (%d) != --gpus (%d): launch with torchrun --nproc-per-node %d" % (world, args.gpus, args.gpus))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    EW, EH, ET, maps, nb = 640, 480, 1024, int(os.environ.get("SIPP_BENCH_ENSEMBLE_MAPS", "128")), int(os.environ.get("SIPP_BENCH_ENSEMBLE_BATCHES", "2"))
    cycles = 40 if args.steps == 200 else max(1, args.steps)          # run_experiment.sh scale unless --steps says otherwise
    params = sipp.make_params(sipp.EVAL_RRT_STAR, half_fov=32, non_path_voxel_gain=0.001)
    worlds = []
    from sipp import shard
    for sd in shard.ensemble_shard([2000 + i for i in range(maps * world)], rank, world):
        lab = synth.make_labels(EW, EH, sd)
        ids_ = sorted(set(np.unique(lab).tolist()) - {0})
        desc = sipp.make_desc(EW, EH, EW * synth.MPP, EH * synth.MPP, (0.5, 0.5), (EW * synth.MPP - 0.5, EH * synth.MPP - 0.5), min(ids_), max(ids_), max(ids_))
        worlds.append((sd, lab, sipp.Context(desc, device=local_rank)))
    cands = {sd: episode.episode_candidates(EW, EH, sd, cycles, ET) for sd, _, _ in worlds}
    groups = [worlds[k::nb] for k in range(nb)]
    batches = [sipp.Batch([w[2] for w in g]) for g in groups]
    for b in batches:
        b.set_sharing(nb)

    def run_group(k, ncyc=None):
        g = groups[k]
        return batches[k].episode_run([w[1] for w in g], params, [cands[w[0]][0][:ncyc] for w in g], [cands[w[0]][1][:ncyc] for w in g])

    with ThreadPoolExecutor(nb) as pool:
        list(pool.map(lambda k: run_group(k, max(3, args.warmup if args.warmup < cycles else 3)), range(nb)))     # warm-up cycles (allocations, clocks)
        for _, _, ctx in worlds:
            ctx.map_clear()
        l0 = sum(b.launch_count() for b in batches)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        t0 = time.perf_counter()
        parts = list(pool.map(run_group, range(nb)))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        clocks = sampler.result()
    value, dt_max = shard.job_throughput(maps * cycles, dt, world)      # all ranks' map-cycles over the slowest rank's time
    launches = sum(b.launch_count() for b in batches) - l0
    h2d = maps * (65 * 65 * 4 + ET * 24)        # per cycle and GPU: every map's patch + candidates (argument blocks not counted)
    d2h = maps * (40 * 4 + 16)                  # plan words + the winner record
    line = {"metric": "ensemble map-cycles/s (640x480 maps, 1024 candidates per cycle)", "value": value, "unit": "map-cycles/s", "n_gpus": world,
            "steps": cycles, "warmup": max(3, args.warmup if args.warmup < cycles else 3), "ms_per_step": dt_max / cycles * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8 counts + f64 gain / f64 cost-to-go", "data": "synthetic",
            "config": {"workload": "BASELINE config 5: %d maps per GPU (640x480, seeds 2000+), %d planning cycles each: patch ingest + incremental follower re-plan + "
                                   "%d candidate views (RRTStarEvaluator count mode) + selection, sipp_batch_episode_run, %d batches side by side per GPU" % (maps, cycles, ET, nb),
                       "maps_per_gpu": maps, "n_gpus": world, "l2": "working set (%d maps x ~6 MB of planes and fields) larger than L2" % maps,
                       "sharding": "maps sharded over the GPUs, no data-path communication"},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": value, "unit": "map-cycles/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "note": "the timed loop is the public call with host buffers (patches and candidates in, plan results and winners out every cycle): value is already end to end"},
            "final_costs_rank0": [tr[-1][0] for tr in parts[0][:4]]}
    for b in batches:
        b.close()
    if rank == 0:
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)      # 200 x ~41 us: long enough for the clock sampler to see the GPU under load
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="graft", choices=["graft", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default=os.environ.get("SIPP_BENCH_WORKLOAD", "evaluation"), choices=["evaluation", "config4", "ensemble"],
                    help="evaluation (default): the BASELINE metric's line; config4: the same line on the 8192^2 map with 65536 candidates in total "
                         "(set SIPP_BENCH_WORKLOAD=config4 instead: the module constants depend on it); ensemble: BASELINE config 5, map-cycles/s")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: how the 16-byte shard winners travel. p2p (default): stored into the peers' mailboxes over NVLink by the argmax "
                         "kernel itself (sipp_cycle_shard); nccl: all_gather + merge kernel. The other one is timed too and reported beside it.")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "ensemble":
        return ensemble_workload(args)

    # stdout carries exactly ONE line (the JSON): everything libraries print while the benchmark runs (NCCL's version banner,
    # torchrun notices) is sent to stderr by pointing fd 1 at fd 2 until the line is written
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    import sipp

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        raise SystemExit("WORLD_SIZE (%d) != --gpus (%d): launch with torchrun --nproc-per-node %d" % (world, args.gpus, args.gpus))
    numa_note = bind_to_gpu_numa_node(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    lab, obs, xy_all, cost_all, ids = build_inputs(world)
    desc = sipp.make_desc(W, H, W * 0.0625, H * 0.0625, (0.5, 0.5), (W * 0.0625 - 0.5, H * 0.0625 - 0.5), min(ids), max(ids), max(ids))
    ctx = sipp.Context(desc, device=local_rank)
    ctx.map_upload_full(lab, obs)
    # ---- follower cost-to-go + path + mask: once per planning cycle (replicated per GPU: the wavefront does not shard)
    ctx.set_incremental(False)          # "plan" below = from scratch; the incremental re-plan is timed at the end (rank 0)
    ctx.plan()
    plan_ms, relax_ms = [], []
    for _ in range(3):
        t0 = time.perf_counter()
        pcost, pstatus, _, plen = ctx.plan()
        plan_ms.append((time.perf_counter() - t0) * 1e3)
        relax_ms.append(ctx.plan_phases()[0])      # the relaxation kernel alone, CUDA events on the plan's stream
    pstats = ctx.plan_stats()

    T = T_PER_GPU
    lo = rank * T
    xy = np.ascontiguousarray(xy_all[lo:lo + T])
    cost = np.ascontiguousarray(cost_all[lo:lo + T])
    params = sipp.make_params(sipp.EVAL_RRT_STAR, half_fov=HALF_FOV, non_path_voxel_gain=NPG)
    d_xy = torch.from_numpy(xy).to(dev)
    d_cost = torch.from_numpy(cost).to(dev)
    d_gain = torch.empty(T, dtype=torch.float64, device=dev)
    d_value = torch.empty(T, dtype=torch.float64, device=dev)
    d_term = torch.empty(T, dtype=torch.uint8, device=dev)
    d_counts = torch.empty((T, 4), dtype=torch.int32, device=dev)
    d_best = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(2)]     # {f64 value bits, i64 index}, double buffered
    d_all = [torch.zeros(2 * world, dtype=torch.int64, device=dev) for _ in range(2)]
    d_glob = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(3)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    stream = torch.cuda.Stream(device=dev)
    cstream = torch.cuda.Stream(device=dev)                          # the collective's stream
    sh = stream.cuda_stream
    e_done = [torch.cuda.Event() for _ in range(2)]
    g_done = [torch.cuda.Event() for _ in range(2)]

    # ---- N > 1: map the peers' mailboxes (CUDA IPC handles moved with torch.distributed; plumbing only)
    p2p_ok, p2p_note = False, None
    if world > 1:
        try:
            handles = [None] * world
            dist.all_gather_object(handles, ctx.peer_export())
            ctx.peer_attach(rank, handles)
            p2p_ok = True
        except Exception as exc:     # e.g. a container without CUDA IPC: the NCCL exchange below still runs on the GPU
            p2p_note = str(exc)
        flag = torch.tensor([1 if p2p_ok else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        p2p_ok = bool(flag.item())
    exchange = "p2p" if (world > 1 and p2p_ok and args.exchange == "p2p") else ("nccl" if world > 1 else "none")

    def step_p2p(k):
        """gain + value + argmax of this rank's shard in ONE launch: the kernel's last CTA stores the shard's 16-byte winner into
        every peer's mailbox over NVLink and collects the records of the PREVIOUS step (already there: the exchange latency hides
        under this step's scan), merges them -> d_glob. No collective library, no second stream; the last step's records are
        collected by a flush that is timed in full (tail)."""
        b = k & 1
        ctx.cycle_shard_pipelined_dev(params, T, d_xy.data_ptr(), None, d_cost.data_ptr(), d_gain.data_ptr(), d_term.data_ptr(),
                                      d_value.data_ptr(), lo, d_glob[b].data_ptr(), sh)
        return d_glob[b]

    def step_p2p_sync(k):
        """same, but every step waits for its own records (the latency a planner that needs the winner before it goes on sees)"""
        b = k & 1
        ctx.cycle_shard_dev(params, T, d_xy.data_ptr(), None, d_cost.data_ptr(), d_gain.data_ptr(), d_term.data_ptr(), d_value.data_ptr(),
                            lo, d_glob[b].data_ptr(), sh)
        return d_glob[b]

    def step_nccl(k):
        """gain + value + argmax of this rank's shard; with N > 1 the only cross-GPU traffic: an all_gather of every shard's
        16-byte (value, index) record + one merge kernel (same argmax on every rank). The gather of step k runs on its own
        stream under the gain kernel of step k+1 (records double buffered); a step ends when the PREVIOUS gather is done."""
        b = k & 1
        ctx.cycle_dev(params, T, d_xy.data_ptr(), None, d_cost.data_ptr(), d_gain.data_ptr(), d_term.data_ptr(), d_value.data_ptr(),
                      d_best[b].data_ptr(), sh)
        if world == 1:
            return d_best[b]
        e_done[b].record(stream)
        with torch.cuda.stream(cstream):
            cstream.wait_event(e_done[b])
            dist.all_gather_into_tensor(d_all[b], d_best[b])
            ctx.merge_best_dev(d_all[b].data_ptr(), world, T, d_glob[b].data_ptr(), cstream.cuda_stream)
            g_done[b].record(cstream)
        if k > 0:
            stream.wait_event(g_done[1 - b])
        return d_glob[b]

    d_align = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(2)]
    d_align[0][1] = -1                          # an "empty shard" record: never wins, only lines the ranks up

    def timed(step, overlapped_gather, flush_peers=False, align=False):
        """W warm-up steps, then K timed ones: barrier + synchronize on both sides, CUDA events around every step on the
        launching stream (the L2 flush between them is untimed), max over ranks of the summed step times.
        align (N > 1, synchronous exchange): an untimed on-device exchange of an empty record sits between the flush and the start
        event of every step. It completes on a rank only when every rank has reached it, so the GPUs enter each timed step within
        the exchange's own latency of each other — without it the per-rank jitter of the untimed 256 MiB flush is charged to
        whichever rank has to wait for the slowest one inside the timed exchange."""
        with torch.cuda.stream(stream):
            for k in range(args.warmup):
                flush.fill_(1)
                best_ = step(k)
            sampler = ClockSampler(local_rank)       # NVML init happens here, before the barrier
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
            ev_tail = torch.cuda.Event(enable_timing=True)
            stream.synchronize()
            cstream.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            sampler.start()
            if world > 1:
                # one more untimed step: the ranks leave the host barrier hundreds of microseconds apart, and the first exchange
                # after it would charge that host skew to step 0. Its exchange lines the GPUs up; the timed steps are queued behind it.
                flush.fill_(2)
                best_ = step(args.warmup)
                if flush_peers:                # pipelined form: collecting this step's records is what waits for every rank
                    ctx.peer_flush_dev(d_glob[2].data_ptr(), sh)
            launches0 = ctx.launch_count()
            wall0 = time.perf_counter()
            k0 = args.warmup + 1 if world > 1 else 0        # keeps the double-buffer parity of the overlapped NCCL variant going
            for k in range(args.steps):
                flush.fill_(k & 0xFF)          # L2 flush between timed iterations (untimed)
                if align:
                    ctx.peer_exchange_dev(d_align[0].data_ptr(), d_align[1].data_ptr(), sh)
                ev[k][0].record(stream)
                best_ = step(k0 + k)
                ev[k][1].record(stream)
            if overlapped_gather:              # the last gather has nothing to hide under: its full latency is added to the total
                stream.wait_event(g_done[(k0 + args.steps - 1) & 1])
            if flush_peers:                    # pipelined peer exchange: the last step's records are collected here, timed in full
                best_ = d_glob[(k0 + args.steps) & 1]
                ctx.peer_flush_dev(best_.data_ptr(), sh)
            ev_tail.record(stream)
            stream.synchronize()
            cstream.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            wall_ = time.perf_counter() - wall0
            launches_ = ctx.launch_count() - launches0 - (args.steps if align else 0)   # the aligning exchanges sit between the timed steps, not in them
            clocks_ = sampler.result()
        tail_ms = ev[-1][1].elapsed_time(ev_tail)
        total = float(sum(a.elapsed_time(b) for a, b in ev)) + float(tail_ms)
        if os.environ.get("SIPP_BENCH_DEBUG"):
            print("[rank %d] %s step ms: %s tail %.4f gaps %s" % (rank, step.__name__, " ".join("%.3f" % a.elapsed_time(b) for a, b in ev), tail_ms,
                                                                  " ".join("%.3f" % ev[i][1].elapsed_time(ev[i + 1][0]) for i in range(len(ev) - 1))), file=sys.stderr)
        t = torch.tensor([total], dtype=torch.float64, device=dev)
        timed.per_rank = [total / args.steps]
        if world > 1:
            allt = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
            dist.all_gather(allt, t)
            timed.per_rank = [float(x.item()) / args.steps for x in allt]
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), best_, wall_, launches_, clocks_

    other = None
    if exchange == "p2p":
        total_ms, best, wall, launches, clocks = timed(step_p2p_sync, False, align=True)   # every step ends with its own global winner
        sync_per_rank = list(timed.per_rank)
        u_ms, u_best, _, _, _ = timed(step_p2p_sync, False)                     # the same without the alignment: what the flush jitter costs
        # the same kernel without any exchange, per rank: step time minus this = what a rank spends waiting for the slowest shard
        l_ms, _, _, _, _ = timed(lambda k: (ctx.cycle_dev(params, T, d_xy.data_ptr(), None, d_cost.data_ptr(), d_gain.data_ptr(), d_term.data_ptr(),
                                                         d_value.data_ptr(), d_best[k & 1].data_ptr(), sh), d_best[k & 1])[1], False)
        local_per_rank = list(timed.per_rank)
        s_ms, s_best, _, _, _ = timed(step_p2p, False, True)
        o_ms, o_best, _, _, _ = timed(step_nccl, True)      # the collective-library form of the same step, for comparison
        assert torch.equal(o_best.cpu(), best.cpu()) and torch.equal(s_best.cpu(), best.cpu())
        other = [{"exchange": "p2p, synchronous, ranks NOT aligned before the timed steps (the untimed L2 flush's per-rank jitter lands in the timed exchange)",
                  "ms_per_step": u_ms / args.steps, "value": world * T * args.steps / (u_ms * 1e-3)},
                 {"exchange": "none (every rank scores its shard and selects locally): the kernel alone, per rank", "ms_per_step_per_rank": local_per_rank,
                  "exchange_wait_us_per_rank": [1e3 * (a - b) for a, b in zip(sync_per_rank, local_per_rank)],
                  "note": "step time with the exchange minus the same rank's step time without it = time spent waiting for the slowest shard + one NVLink store / poll"},
                 {"exchange": "p2p, pipelined (step k collects the records of step k-1; the ranks may drift up to one step apart and the final flush, "
                              "timed in full, waits for the slowest)", "ms_per_step": s_ms / args.steps, "value": world * T * args.steps / (s_ms * 1e-3)},
                 {"exchange": "nccl all_gather + merge kernel (gather of step k under the gain kernel of step k+1)", "ms_per_step": o_ms / args.steps,
                  "value": world * T * args.steps / (o_ms * 1e-3)}]
    else:
        total_ms, best, wall, launches, clocks = timed(step_nccl, world > 1)
    value = world * T * args.steps / (total_ms * 1e-3)
    best_host = best.cpu().numpy()
    best_value = float(best_host[:1].view(np.float64)[0])
    best_index = int(best_host[1])

    # ---- roofline of the dominant kernel (k_gain alone, CUDA events on its stream, L2 flushed before each launch)
    peak, peak_src = measured_peaks()
    gain_ms = []
    with torch.cuda.stream(stream):
        for k in range(max(5, args.steps // 2)):
            flush.fill_(k)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            ctx.gain_batch_dev(params, T, d_xy.data_ptr(), None, d_gain.data_ptr(), d_counts.data_ptr(), d_term.data_ptr(), sh)
            e1.record(stream)
            stream.synchronize()
            gain_ms.append(e0.elapsed_time(e1))
    g_ms = float(np.mean(gain_ms))
    achieved = ALG_BYTES_PER_EVAL * T / (g_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "gain_traffic.json")     # written from the committed ncu capture (bytes per launch)
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "k_gain<MODE_COUNT, packed>", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "kernel_ms": g_ms, "algorithmic_bytes_per_launch": ALG_BYTES_PER_EVAL * T,
                "note": "algorithmic bytes = one byte per window cell + pose + result (SURVEY 8(d)); the kernel reads a 2-bit-per-cell plane (4 MiB, L2 resident "
                        "by nature: 126 MB L2), so frac > 1 of the HBM copy peak is expected; see traffic. It is bound by instruction issue and L2 latency, not by bytes"}

    # ---- end to end through the host-pointer C ABI (sipp_cycle): pinned host buffers, H2D + D2H inside the timed region
    h_xy = torch.from_numpy(xy).pin_memory()
    h_cost = torch.from_numpy(cost).pin_memory()
    h_gain = torch.empty(T, dtype=torch.float64).pin_memory()
    h_value = torch.empty(T, dtype=torch.float64).pin_memory()
    h_term = torch.empty(T, dtype=torch.uint8).pin_memory()
    h_rec = torch.zeros(2, dtype=torch.int64).pin_memory()
    h_win = torch.zeros(2, dtype=torch.int64).pin_memory()
    h_rec_np, h_win_np = h_rec.numpy(), h_win.numpy()      # same pinned memory, without the tensor-indexing overhead

    def e2e_cycle():
        """the call a user makes: host buffers in, host results out"""
        return ctx.cycle_host_ptr(params, T, h_xy.data_ptr(), None, h_cost.data_ptr(), h_gain.data_ptr(), h_term.data_ptr(), h_value.data_ptr())

    def e2e_exchange(bi_, bv_):
        """N > 1: exchange of the shard winners, enqueued on the collective's stream (it completes under the next cycle call)"""
        h_rec_np[:1].view(np.float64)[0] = bv_
        h_rec_np[1] = bi_
        with torch.cuda.stream(cstream):
            d_best[0].copy_(h_rec, non_blocking=True)
            dist.all_gather_into_tensor(d_all[0], d_best[0])
            ctx.merge_best_dev(d_all[0].data_ptr(), world, T, d_glob[0].data_ptr(), cstream.cuda_stream)
            h_win.copy_(d_glob[0], non_blocking=True)

    def e2e_run(steps):
        bi_, bv_ = -1, 0.0
        if exchange == "p2p":              # one call per cycle and rank: host buffers in, host results + the GLOBAL winner out
            for _ in range(steps):
                bi_, bv_ = ctx.cycle_shard_host_ptr(params, T, h_xy.data_ptr(), None, h_cost.data_ptr(), lo, h_gain.data_ptr(), h_term.data_ptr(),
                                                    h_value.data_ptr())
            return bi_, bv_
        for _ in range(steps):
            if world > 1:
                cstream.synchronize()          # the previous exchange (h_rec / h_win are reused)
            bi_, bv_ = e2e_cycle()
            if world > 1:
                e2e_exchange(bi_, bv_)
        if world > 1:
            cstream.synchronize()
            bi_, bv_ = int(h_win_np[1]), float(h_win_np[:1].view(np.float64)[0])
        return bi_, bv_

    e2e_run(3)
    if world > 1:
        dist.barrier()
    e2e_rep = []
    for _ in range(5):                     # host-timed: median of 5 runs of K steps each (wall clock jitters more than CUDA events)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        bi, bv = e2e_run(args.steps)
        e2e_rep.append(time.perf_counter() - t0)
    e2e_s = float(np.median(e2e_rep))
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    e2e = {"value": world * T * args.steps / e2e_s, "unit": "evals/s", "h2d_bytes_per_step": T * 24, "d2h_bytes_per_step": T * 17 + 16,
           "ms_per_step": e2e_s / args.steps * 1e3, "host_placement": numa_note,
           "api": ("sipp_cycle_shard (host pointers, pinned; copy/compute pipeline + argmax with the fused peer-memory exchange, replayed as a CUDA graph)"
                   if exchange == "p2p" else
                   "sipp_cycle (host pointers, pinned; 2-chunk copy/compute pipeline replayed as a CUDA graph)"
                   + (" + all_gather of the 16-byte shard winners + sipp_merge_best_dev, enqueued after each call and completing under the next one" if world > 1 else ""))}
    assert bi == best_index and bv == best_value, (bi, best_index, bv, best_value)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- the planning cycle as the planner runs it: ingest the scout's 65x65 patch, re-plan INCREMENTALLY, evaluate
    from sipp import episode
    ctx.set_incremental(True)
    ctx.plan()
    rng = np.random.default_rng(5)
    rim = (obs == 0) & (np.roll(obs, 40, 0) | np.roll(obs, 40, 1) | np.roll(obs, -40, 0) | np.roll(obs, -40, 1)).astype(bool)
    ys, xs = np.nonzero(rim)
    replan_ms, ingest_ms, inval = [], [], []
    for _ in range(12):
        j = int(rng.integers(0, len(xs)))
        x0, y0 = int(xs[j]) - 32, int(ys[j]) - 32
        patch = episode.extract_patch(lab, x0, y0, 65, 65)
        t0 = time.perf_counter()
        ctx.map_update_patch(x0, y0, patch)
        t1 = time.perf_counter()
        ctx.plan(path_cap=1)
        t2 = time.perf_counter()
        ingest_ms.append((t1 - t0) * 1e3)
        replan_ms.append((t2 - t1) * 1e3)
        inval.append(ctx.plan_info()["invalidated_cells"])
    incremental = {"replan_ms_median": float(np.median(replan_ms)), "replan_ms_min": float(min(replan_ms)), "replan_ms_max": float(max(replan_ms)),
                   "patch_ingest_ms_median": float(np.median(ingest_ms)), "invalidated_cells_median": float(np.median(inval)),
                   "note": "12 65x65 patches at the rim of the observed region, one per cycle; sipp_plan reuses the previous fixed point "
                           "(bit-identical to a plan from scratch: tests/test_gpu_parity.py)"}

    # ---- CPU baseline beside it: the oracle port of the reference evaluators on this box's host cores (bounded sample)
    cpu = None
    config1 = None
    if not args.no_cpu_baseline and world == 1:      # the CPU baseline belongs to the N = 1 line (the reference arm times it at every N)
        from oracle import oracle as orc
        odesc = orc.make_desc(W, H, W * 0.0625, H * 0.0625, (0.5, 0.5), (W * 0.0625 - 0.5, H * 0.0625 - 0.5), min(ids), max(ids), max(ids))
        o = orc.Oracle(odesc)
        o.map_upload_full(lab, obs)
        t0 = time.perf_counter()
        ocost, ostatus, _, olen = o.plan()
        oplan_s = time.perf_counter() - t0
        po = orc.make_params(orc.EVAL_RRT_STAR, half_fov=HALF_FOV, non_path_voxel_gain=NPG)
        n1 = 256
        t0 = time.perf_counter()
        og, ot, ov, oi, obv = o.cycle(po, xy[:n1], cost[:n1], variant=0, threads=1)
        r1 = n1 / (time.perf_counter() - t0)
        ns = int(min(T, max(n1, r1 * 12.0)))
        t0 = time.perf_counter()
        og, ot, ov, oi, obv = o.cycle(po, xy[:ns], cost[:ns], variant=0, threads=1)
        b1 = ns / (time.perf_counter() - t0)
        t0 = time.perf_counter()
        o.cycle(po, xy[:ns], cost[:ns], variant=1, threads=1)
        b2 = ns / (time.perf_counter() - t0)
        cores = os.cpu_count() or 1
        nb3 = min(T, ns * cores)
        t0 = time.perf_counter()
        o.cycle(po, xy[:nb3], cost[:nb3], variant=1, threads=cores)
        b3 = nb3 / (time.perf_counter() - t0)
        # parity spot-check of the timed GPU result against the oracle on the sample (the full check lives in tests/)
        hg = h_gain.numpy()[:ns]
        hv = h_value.numpy()[:ns]
        gi = int(np.argmax(hv))                # first maximum = the candidate the GPU's argmax selects among the sample
        parity = bool(np.allclose(hg, og, rtol=1e-5, atol=0) and (pcost == ocost) and (plen == olen) and gi == oi)
        cpu = {"value": b1, "unit": "evals/s", "cores": 1, "kind": "port",
               "sample": "first %d of the %d candidates, 1 thread, faithful port of the reference evaluator path (voxel vector + quadratic erase loop)" % (ns, T),
               "linear_filter_1core": b2, "linear_filter_all_cores": b3, "host_cores": cores,
               "plan_s": oplan_s, "parity_on_sample": parity, "selected_on_sample": {"gpu": gi, "cpu": int(oi)}}
        config1 = config1_cycle()

    line = {
        "metric": "candidate-view evals/sec on 4096^2 cost map", "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": SCALING, "vs_baseline": None,
        "dtype": "u8 counts + f64 gain", "data": "synthetic",
        "config": bench_config(world),
        "exchange": {"kind": exchange, "note": p2p_note, "other": other,
                     "how": ("single GPU" if world == 1 else
                             "the gain kernel's last CTA stores each shard's 16-byte winner into every peer's mailbox over NVLink (peer memory), waits for the "
                             "peers' records of the same step and merges them: every step ends with the global winner on every rank, one launch per step, "
                             "no collective call on the data path" if exchange == "p2p" else
                             "NCCL all_gather of one 16-byte record per rank per step + one merge kernel; the gather of step k overlaps the gain kernel of "
                             "step k+1, the last gather is timed in full")},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        "cycle": {"plan_ms": float(np.median(plan_ms)), "plan_cost": pcost, "plan_path_nodes": plen, "plan_stats": pstats,
                  "relax_roofline": {"bound": "hbm", "kernel": "k_relax", "kernel_ms": float(np.median(relax_ms)),
                                     "achieved": 10.0 * W * H / (float(np.median(relax_ms)) * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                     "frac": 10.0 * W * H / (float(np.median(relax_ms)) * 1e-3) / 1e9 / peak,
                                     "algorithmic_bytes_per_launch": 10 * W * H,
                                     "note": "10 B per cell settled (u16 code read + f64 written); the kernel is bound by its chain of dependent tile "
                                             "activations, not by bytes (DESIGN.md 4.2)"},
                  "eval_ms": total_ms / args.steps, "ms_per_full_cycle": float(np.median(plan_ms)) + total_ms / args.steps,
                  "incremental": incremental,
                  "ms_per_cycle_incremental": incremental["patch_ingest_ms_median"] + incremental["replan_ms_median"] + total_ms / args.steps,
                  "note": "plan = fp64 cost-to-go field + canonical path + mask through sipp_plan FROM SCRATCH (host call, wall clock); "
                          "incremental = the same call after a patch ingest, warm-started from the previous field"},
        "selected": {"index": best_index, "value": best_value}, "wall_s_timed_region": wall,
        "config1": config1,
    }
    sys.stdout.flush()
    os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
