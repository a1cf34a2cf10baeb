"""Probe: sipp_cycle_dev reading poses / costs from, and writing results to, PINNED HOST memory directly (zero copy over PCIe)
versus the staged 2-chunk pipeline of sipp_cycle. Prints ms per call for both."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scouting-ipp_b200"))
import numpy as np, torch
import sipp
from sipp import synth

W = H = 4096; T = 65536
lab = synth.make_labels(W, H, 1003); obs = synth.make_observed(W, H, 1003, 1500)
xy, cost = synth.make_candidates(W, H, 1003, T)
ids = sorted(set(np.unique(lab).tolist()) - {0})
ctx = sipp.Context(sipp.make_desc(W, H, W * 0.0625, H * 0.0625, (0.5, 0.5), (W * 0.0625 - 0.5, H * 0.0625 - 0.5), min(ids), max(ids), max(ids)))
ctx.map_upload_full(lab, obs); ctx.plan()
p = sipp.make_params(sipp.EVAL_RRT_STAR, half_fov=32, non_path_voxel_gain=0.001)
dev = torch.device("cuda", 0)
h_xy = torch.from_numpy(xy).pin_memory(); h_cost = torch.from_numpy(cost).pin_memory()
h_gain = torch.empty(T, dtype=torch.float64).pin_memory(); h_value = torch.empty(T, dtype=torch.float64).pin_memory()
h_term = torch.empty(T, dtype=torch.uint8).pin_memory(); h_best = torch.zeros(2, dtype=torch.int64).pin_memory()
d_value = torch.empty(T, dtype=torch.float64, device=dev); d_best = torch.zeros(2, dtype=torch.int64, device=dev)
d_xy = torch.from_numpy(xy).to(dev); d_cost = torch.from_numpy(cost).to(dev)
d_gain = torch.empty(T, dtype=torch.float64, device=dev); d_term = torch.empty(T, dtype=torch.uint8, device=dev)
st = torch.cuda.Stream(device=dev); sh = st.cuda_stream

def run(name, fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    print("%-50s %.1f us/call" % (name, (time.perf_counter() - t0) / n * 1e6), flush=True)

def staged():
    ctx.cycle_host_ptr(p, T, h_xy.data_ptr(), None, h_cost.data_ptr(), h_gain.data_ptr(), h_term.data_ptr(), h_value.data_ptr())
def device_only():
    ctx.cycle_dev(p, T, d_xy.data_ptr(), None, d_cost.data_ptr(), d_gain.data_ptr(), d_term.data_ptr(), d_value.data_ptr(), d_best.data_ptr(), sh); st.synchronize()
def zc_all():      # everything in pinned host memory (value is re-read by the argmax over PCIe)
    ctx.cycle_dev(p, T, h_xy.data_ptr(), None, h_cost.data_ptr(), h_gain.data_ptr(), h_term.data_ptr(), h_value.data_ptr(), h_best.data_ptr(), sh); st.synchronize()
def zc_in():       # inputs from host, outputs to the device
    ctx.cycle_dev(p, T, h_xy.data_ptr(), None, h_cost.data_ptr(), d_gain.data_ptr(), d_term.data_ptr(), d_value.data_ptr(), d_best.data_ptr(), sh); st.synchronize()
def zc_out():      # inputs on the device, gain / terminal to host, value on the device
    ctx.cycle_dev(p, T, d_xy.data_ptr(), None, d_cost.data_ptr(), h_gain.data_ptr(), h_term.data_ptr(), d_value.data_ptr(), d_best.data_ptr(), sh); st.synchronize()
def zc_mixed():    # inputs from host, gain / terminal to host, value on the device + one D2H copy of value and best
    ctx.cycle_dev(p, T, h_xy.data_ptr(), None, h_cost.data_ptr(), h_gain.data_ptr(), h_term.data_ptr(), d_value.data_ptr(), d_best.data_ptr(), sh)
    with torch.cuda.stream(st):
        h_value.copy_(d_value, non_blocking=True); h_best.copy_(d_best, non_blocking=True)
    st.synchronize()
run("sipp_cycle host pointers (ZEROCOPY=%s CHUNKS=%s)" % (os.environ.get("SIPP_CYCLE_ZEROCOPY", "1"), os.environ.get("SIPP_CYCLE_CHUNKS", "auto")), staged, 200)
if os.environ.get("PROBE_SHORT"):
    sys.exit(0)
run("device-resident sipp_cycle_dev + sync", device_only)
run("zero-copy inputs only", zc_in)
run("zero-copy outputs only (gain, terminal)", zc_out)
run("zero-copy in + gain/term out, value D2H copy", zc_mixed)
run("zero-copy everything (argmax reads host)", zc_all)
ref = h_gain.numpy().copy(); staged(); assert np.array_equal(ref, h_gain.numpy())
print("zero-copy results identical to the staged path")
