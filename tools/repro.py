import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scouting-ipp_b200"))
import numpy as np
import sipp
W = H = 256
g = sipp.Context(sipp.make_desc(W, H, 16.0, 16.0, (0.5, 0.5), (15.5, 15.5), 30, 120, 30))
lab = np.full((H, W), 30, np.int32)
obs = np.zeros((H, W), np.uint8); obs[:, :100] = 1
g.map_upload_full(lab, obs)
print("upload ok", flush=True)
s, c, r, p = g.map_download()
print("download ok", s.sum(), flush=True)
hf = int(sys.argv[1]) if len(sys.argv) > 1 else 32
print(g.gain_batch(sipp.make_params(sipp.EVAL_NAIVE, half_fov=hf), np.array([[5.0, 5.0], [6.3, 8.0]])))
