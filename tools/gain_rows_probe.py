"""Is k_gain bound by the number of TMA box rows per candidate? Gain-only launches at several half_fov, packed and byte planes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scouting-ipp_b200"))
import numpy as np, torch
import sipp
from sipp import synth
W = 4096; T = 65536
lab, obs, desc = synth.world(W, W, 1003, 1500, sipp.make_desc, init_cost="max")
ctx = sipp.Context(desc); ctx.map_upload_full(lab, obs); ctx.plan()
dev = torch.device("cuda", 0)
xy, cost = synth.make_candidates(W, W, 1003, T)
d_xy = torch.from_numpy(xy).to(dev); d_gain = torch.empty(T, dtype=torch.float64, device=dev)
d_cnt = torch.empty((T, 4), dtype=torch.int32, device=dev); d_term = torch.empty(T, dtype=torch.uint8, device=dev)
st = torch.cuda.Stream(device=dev)
for hf in (8, 16, 24, 32, 40, 48):
    p = sipp.make_params(sipp.EVAL_RRT_STAR, half_fov=hf, non_path_voxel_gain=0.001)
    ts = []
    with torch.cuda.stream(st):
        for k in range(12):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st); ctx.gain_batch_dev(p, T, d_xy.data_ptr(), None, d_gain.data_ptr(), d_cnt.data_ptr(), d_term.data_ptr(), st.cuda_stream); b.record(st)
            st.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    print("half_fov %2d (%3d rows): %.1f us  (%.2f ns per row-candidate)" % (hf, 2 * hf + 1, np.median(ts[2:]), np.median(ts[2:]) * 1e3 / (T * (2 * hf + 1))), flush=True)
