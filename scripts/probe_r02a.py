"""Round-2 probe: sub-pixel C2 / C3 fusion times and C2 registration accuracy."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from multiview_stitcher_b200 import fusion, geometry, synthetic, registration
import bench

def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

out = {}
for name, grid, tile, ov, dt in (("C2", (5, 5), (2048, 2048), (307, 307), np.float32),
                                 ("C3", (2, 4, 4), (256, 512, 512), (26, 51, 51), np.uint16)):
    for sub in (False, True):
        t0 = time.time()
        views, stage, true = synthetic.make_grid(grid, tile, ov, dt, jitter=2, seed=0, subpixel=sub)
        torch.cuda.synchronize(); gen_s = time.time() - t0
        bbs = [v.bb() for v in views]
        osp = geometry.union_stack_props(bbs, true, bbs[0]["spacing"])
        plan = fusion.FusionPlan(views, true, osp)
        ms = timeit(plan.run)
        b = plan.algorithmic_bytes()
        out[f"{name}_sub{int(sub)}"] = {"ms": ms, "GBps": b / ms / 1e6, "frac": b / ms / 1e6 / 6553.3, "gen_s": gen_s, "vox": plan.out_voxels}
        print(name, sub, out[f"{name}_sub{int(sub)}"], flush=True)
        if name == "C2" and sub:
            pairs = bench.c2_pairs()
            fixed, moving = bench.pair_crops([v.tensor for v in views], pairs)
            fixed = [f.contiguous() for f in fixed]; moving = [m.contiguous() for m in moving]
            plans = {}
            res = registration.register_pairs(fixed, moving, plans=plans)
            t = np.array([p[:2, 2] for p in true])
            err = [np.abs(r["affine_matrix"][:2, 2] + (t[b_] - t[a_])).max() for r, (a_, b_, _) in zip(res, pairs)]
            ms_reg = timeit(lambda: registration.register_pairs(fixed, moving, plans=plans), n=5, warm=2)
            out["C2_reg_sub"] = {"ms": ms_reg, "max_err_vs_truth": float(max(err)), "mean_err": float(np.mean(err)), "q_min": float(min(r["quality"] for r in res))}
            print(out["C2_reg_sub"], flush=True)
        plan.close(); del views, plan
        torch.cuda.empty_cache()
json.dump(out, open("gpurun_out/probe_r02a.json", "w"), indent=1)
