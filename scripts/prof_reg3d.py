"""ncu target: registration of C3's face pairs from the resident tiles (a 2x2x2 sub-grid)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from multiview_stitcher_b200 import pairs as pairs_mod, synthetic
grid = tuple(int(x) for x in os.environ.get("GRID", "2,2,2").split(","))
views, stage, true = synthetic.make_grid(grid, bench.C3["tile"], bench.C3["overlap"], np.uint16, jitter=2, seed=1, subpixel=True)
pairs = bench._c3_pairs(grid)
plans = {}
pplan = pairs_mod.PairPlan(views, stage, pairs, registration_binning={"z": 1, "y": 1, "x": 1})
for _ in range(2):
    pairs_mod.register_views(views, plan=pplan, pc_plans=plans)
torch.cuda.synchronize()
t0 = time.perf_counter(); pairs_mod.register_views(views, plan=pplan, pc_plans=plans); torch.cuda.synchronize()
print("pairs", len(pairs), "ms", (time.perf_counter() - t0) * 1e3)
