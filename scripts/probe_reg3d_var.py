"""Call-to-call variation of the 3-D registration from tiles (64 C3 face pairs)."""
import os, sys, time, cProfile, pstats, io
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from multiview_stitcher_b200 import pairs as pairs_mod, synthetic
grid = bench.C3["grid"]
views, stage, true = synthetic.make_grid(grid, bench.C3["tile"], bench.C3["overlap"], np.uint16, jitter=2, seed=bench.SEED, subpixel=True)
pairs = bench._c3_pairs(grid)
plans = {}
pplan = pairs_mod.PairPlan(views, stage, pairs, registration_binning={"z": 1, "y": 1, "x": 1})
ts = []
for i in range(8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    if i == 6:
        pr = cProfile.Profile(); pr.enable()
    pairs_mod.register_views(views, plan=pplan, pc_plans=plans)
    torch.cuda.synchronize()
    if i == 6:
        pr.disable()
    ts.append((time.perf_counter() - t0) * 1e3)
print("ms per call:", " ".join(f"{t:.0f}" for t in ts), "mem GB", torch.cuda.memory_reserved() / 1e9, torch.cuda.mem_get_info())
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(14); print(s.getvalue()[:3000])
