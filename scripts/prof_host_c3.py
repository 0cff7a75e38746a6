import cProfile, pstats, os, sys, io
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiview_stitcher_b200 import fusion, geometry, synthetic
grid, tile, ov = (2, 4, 4), (256, 512, 512), (26, 51, 51)
views, stage, true = synthetic.make_grid(grid, tile, ov, np.uint16, jitter=2, seed=0)
osp = geometry.union_stack_props([v.bb() for v in views], true, views[0].spacing)
f = lambda: fusion.fuse(views, true, output_stack_properties=osp, weights_func=fusion.content_based, output_on_backend=True)
f(); torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable(); f(); torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45); print(s.getvalue()[:9000])
