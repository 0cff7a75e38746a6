"""ncu target: one registration step on C2's 40 pairs (crops resident)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from multiview_stitcher_b200 import registration, synthetic
views, stage, true = synthetic.make_grid(bench.GRID, bench.TILE, bench.OVERLAP, np.float32, jitter=2, seed=1, subpixel=True)
pairs = bench.c2_pairs()
fixed, moving = bench.pair_crops([v.tensor for v in views], pairs)
fixed = [f.contiguous() for f in fixed]; moving = [m.contiguous() for m in moving]
plans = {}
for _ in range(3):
    registration.register_pairs(fixed, moving, plans=plans)
torch.cuda.synchronize()
t0 = time.perf_counter(); registration.register_pairs(fixed, moving, plans=plans); torch.cuda.synchronize()
print("reg step ms", (time.perf_counter() - t0) * 1e3)
