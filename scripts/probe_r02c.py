"""Round-2 probe: fused-kernel times on C2 / C3 / C5 row (default kernels), sub-pixel tiles."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, ".")
from multiview_stitcher_b200 import fusion, geometry, synthetic

def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

cfgs = {"C2": ((5, 5), (2048, 2048), (307, 307), np.float32), "C3": ((2, 4, 4), (256, 512, 512), (26, 51, 51), np.uint16),
        "C5row": ((1, 1, 8), (512, 2048, 2048), (0, 205, 205), np.uint16)}
out = {}
for name in sys.argv[1:] or ["C2", "C3", "C5row"]:
    grid, tile, ov, dt = cfgs[name]
    views, stage, true = synthetic.make_grid(grid, tile, ov, dt, jitter=2, seed=0, subpixel=True)
    bbs = [v.bb() for v in views]
    osp = geometry.union_stack_props(bbs, true, bbs[0]["spacing"])
    plan = fusion.FusionPlan(views, true, osp)
    ms = timeit(plan.run)
    b = plan.algorithmic_bytes()
    out[name] = {"ms": ms, "frac": b / ms / 1e6 / 6553.3}
    print(name, out[name], flush=True)
    plan.close(); del views, plan; torch.cuda.empty_cache()
json.dump(out, open("gpurun_out/probe_r02c.json", "w"), indent=1)
