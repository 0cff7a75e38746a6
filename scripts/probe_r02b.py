"""Round-2 probe: 3-D stencil kernels (z-marching vs MVS_STENCIL3_OLD=1) on C3 and a C5 row."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from multiview_stitcher_b200 import fusion, geometry, synthetic

def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

out = {}
which = sys.argv[1:] or ["C3", "C5row"]
cfgs = {"C3": ((2, 4, 4), (256, 512, 512), (26, 51, 51)), "C5row": ((1, 1, 8), (512, 2048, 2048), (0, 205, 205)),
        "C3int": ((2, 4, 4), (256, 512, 512), (26, 51, 51))}
for name in which:
    grid, tile, ov = cfgs[name]
    views, stage, true = synthetic.make_grid(grid, tile, ov, np.uint16, jitter=2, seed=0, subpixel=(name != "C3int"))
    bbs = [v.bb() for v in views]
    osp = geometry.union_stack_props(bbs, true, bbs[0]["spacing"])
    res = {}
    outs = {}
    for mode in ("new", "old"):
        if mode == "new": os.environ["MVS_STENCIL3"] = "1"
        else: os.environ.pop("MVS_STENCIL3", None)
        plan = fusion.FusionPlan(views, true, osp)
        ms = timeit(plan.run)
        b = plan.algorithmic_bytes()
        res[mode] = {"ms": ms, "frac": b / ms / 1e6 / 6553.3}
        outs[mode] = plan.out.clone() if name != "C5row" else plan.out[:, ::7, ::5].clone()
        plan.close(); del plan
    d = (outs["new"].to(torch.int32) - outs["old"].to(torch.int32)).abs()
    res["max_diff_new_vs_old"] = int(d.max()); res["n_diff"] = int((d > 0).sum())
    out[name] = res
    print(name, res, flush=True)
    del views, outs; torch.cuda.empty_cache()
json.dump(out, open("gpurun_out/probe_r02b.json", "w"), indent=1)
