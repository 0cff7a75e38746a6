"""ncu target: the C2 headline launch (fuse_stencil_kernel<2,float,WAVG>, sub-pixel tiles)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from multiview_stitcher_b200 import fusion, geometry, synthetic
views, stage, true = synthetic.make_grid(bench.GRID, bench.TILE, bench.OVERLAP, np.float32, jitter=2, seed=bench.SEED, subpixel=True)
osp = geometry.union_stack_props([v.bb() for v in views], true, views[0].spacing)
plan = fusion.FusionPlan(views, true, osp)
for _ in range(5):
    plan.run()
torch.cuda.synchronize()
print("bytes", plan.algorithmic_bytes())
