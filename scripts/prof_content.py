"""ncu target: content-weighted fusion of a few interior chunks of a C3-like grid."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiview_stitcher_b200 import content, fusion, geometry, synthetic
grid, tile, ov = (2, 3, 3), (256, 512, 512), (26, 51, 51)
views, stage, true = synthetic.make_grid(grid, tile, ov, np.uint16, jitter=2, seed=0, subpixel=True)
osp = geometry.union_stack_props([v.bb() for v in views], true, views[0].spacing)
n = len(geometry.chunk_grid(osp, geometry.DEFAULT_CHUNKSIZE_3D))
sub = list(range(n))[n // 2 - 3 : n // 2 + 3]
def run():
    return content.fuse_with_weights(views, true, osp, None, fusion.weighted_average_fusion, fusion.content_based, None, 1, None, chunk_subset=sub)
run(); torch.cuda.synchronize()
t0 = time.perf_counter(); run(); torch.cuda.synchronize()
print("chunks", len(sub), "of", n, "ms per chunk", (time.perf_counter() - t0) * 1e3 / len(sub))
