"""ncu target: C3-like 3-D uint16 grid at sub-pixel positions, a few fused launches."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiview_stitcher_b200 import fusion, geometry, synthetic
grid, tile, ov = (2, 2, 4), (256, 512, 512), (26, 51, 51)
views, stage, true = synthetic.make_grid(grid, tile, ov, np.uint16, jitter=2, seed=0, subpixel=True)
osp = geometry.union_stack_props([v.bb() for v in views], true, views[0].spacing)
plan = fusion.FusionPlan(views, true, osp)
for _ in range(4):
    plan.run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); plan.run(); e1.record(); torch.cuda.synchronize()
print("ms", e0.elapsed_time(e1), "GB/s", plan.algorithmic_bytes() / e0.elapsed_time(e1) / 1e6, "blocks", plan.blocks)
