import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from multiview_stitcher_b200 import registration, synthetic
views, stage, true = synthetic.make_grid(bench.GRID, bench.TILE, bench.OVERLAP, np.float32, jitter=2, seed=1, subpixel=True)
pairs = bench.c2_pairs()
host = [v.tensor.cpu().numpy() for v in views]
hf, hm = bench.pair_crops(host, pairs)
hf = [np.ascontiguousarray(a) for a in hf]; hm = [np.ascontiguousarray(a) for a in hm]
plans = {}
for _ in range(2): registration.register_pairs(hf, hm, plans=plans)
ts = []
for _ in range(16):
    t0 = time.perf_counter(); registration.register_pairs(hf, hm, plans=plans); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
print(os.environ.get("MVS_REG_UPLOAD_EARLY"), " ".join(f"{t:.1f}" for t in ts))
