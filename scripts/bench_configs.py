#!/usr/bin/env python
"""Ad-hoc measurements of the other BASELINE configs (not the bench.py contract):
C3 (4x4x2 grid of 256x512x512 uint16 tiles) fusion with cosine-edge blending, one
content-weighted chunk, and 3-D pair registration; C1 (2x1, 256^2 uint16)."""

import json
import sys
import time
import os

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiview_stitcher_b200 import fusion, geometry, registration, synthetic  # noqa: E402


def timed(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    out = {}
    # ---- C3 fusion ----
    grid, tile, ov = (2, 4, 4), (256, 512, 512), (26, 51, 51)
    views, stage, true = synthetic.make_grid(grid, tile, ov, np.uint16, jitter=2, seed=0)
    osp = geometry.union_stack_props([v.bb() for v in views], true, views[0].spacing)
    plan = fusion.FusionPlan(views, true, osp)
    ms = timed(plan.run)
    b = plan.algorithmic_bytes()
    out["c3_fusion"] = {"out_shape": [osp["shape"][d] for d in "zyx"], "ms": ms, "Mvoxel_per_s": plan.out_voxels / ms / 1e3,
                        "GB_per_s": b / ms / 1e6, "frac_of_6545": b / ms / 1e6 / 6545.3, "launches": plan.launches_per_run}
    plan.close()
    # ---- C3 registration: face pairs along x (crop 256 x 512 x 51) ----
    idx = list(np.ndindex(*grid))
    pairs = [(i, i + 1) for i, t in enumerate(idx) if t[2] + 1 < grid[2]][:8]
    fixed = [views[a].tensor[:, :, -ov[2]:].to(torch.float32).contiguous() for a, b_ in pairs]
    moving = [views[b_].tensor[:, :, : ov[2]].to(torch.float32).contiguous() for a, b_ in pairs]
    plans = {}
    res = registration.register_pairs(fixed, moving, plans=plans)
    t0 = time.perf_counter()
    res = registration.register_pairs(fixed, moving, plans=plans)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    true_t = np.array([t[:3, 3] for t in true])
    err = max(float(np.abs(r["affine_matrix"][:3, 3] + (true_t[b_] - true_t[a])).max()) for r, (a, b_) in zip(res, pairs))
    out["c3_registration"] = {"pairs": len(pairs), "crop": list(fixed[0].shape), "pairs_per_s": len(pairs) / dt, "max_shift_err_px": err,
                              "quality_min": min(r["quality"] for r in res)}
    # ---- C3 registration from the resident tiles: all 64 face pairs through pairs.register_views
    # (overlap boxes, crop windows, resampling onto the fixed tile's grid, registration, physical
    # transform); binning 1 and the reference's default heuristic (bins z by 2 here)
    from multiview_stitcher_b200 import pairs as pairs_mod

    all_pairs = []
    for i, t in enumerate(idx):
        for ax in range(3):
            if t[ax] + 1 < grid[ax]:
                u = list(t)
                u[ax] += 1
                all_pairs.append((i, idx.index(tuple(u))))
    for label, binning in (("c3_registration_from_tiles", {"z": 1, "y": 1, "x": 1}), ("c3_registration_from_tiles_default_binning", None)):
        t0 = time.perf_counter()
        pplan = pairs_mod.PairPlan(views, stage, all_pairs, registration_binning=binning)
        plan_s = time.perf_counter() - t0
        pcp = {}
        res = pairs_mod.register_views(views, plan=pplan, pc_plans=pcp)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = pairs_mod.register_views(views, plan=pplan, pc_plans=pcp)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        err = max(float(np.abs(r["transform"][:3, 3] + (true_t[b_] - true_t[a])).max()) for r, (a, b_) in zip(res, all_pairs))
        out[label] = {"pairs": len(all_pairs), "crops": sorted({tuple(it["shape"]) for it in pplan.items}),
                      "binning": sorted({it["binning"] for it in pplan.items}), "pairs_per_s": len(all_pairs) / dt,
                      "ms": dt * 1e3, "host_geometry_plan_s_once": plan_s, "max_shift_err_world": err,
                      "quality_min": min(float(r["quality"]) for r in res)}
        for p_ in pcp.values():
            p_.close()
        del pcp, res, pplan
        torch.cuda.empty_cache()
    # ---- C3 content-weighted fusion of one 256^3 chunk (sigma 5 / 11, halo 22) ----
    sub = {"origin": {d: osp["origin"][d] + 300 * osp["spacing"][d] for d in "zyx"}, "spacing": osp["spacing"],
           "shape": {"z": 128 + 44, "y": 256 + 44, "x": 256 + 44}}
    sub["origin"]["z"] = osp["origin"]["z"] + 100
    bbs = [v.bb() for v in views]
    sel = [i for i in range(len(views)) if not (np.any(geometry.transformed_aabb(bbs[i], true[i], list("zyx"))[1] < np.array([sub["origin"][d] for d in "zyx"])) or
                                                 np.any(geometry.transformed_aabb(bbs[i], true[i], list("zyx"))[0] > np.array([sub["origin"][d] + sub["shape"][d] for d in "zyx"])))]
    def cw():
        return fusion.fuse_np([views[i] for i in sel], [true[i] for i in sel], sub, weights_func=fusion.content_based,
                              full_view_bbs=[bbs[i] for i in sel], trim_overlap_in_pixels=22, output_on_backend=True)
    ms = timed(cw, n=2, warm=1)
    vox = 128 * 256 * 256
    out["c3_content_chunk"] = {"views": len(sel), "chunk": [128, 256, 256], "halo": 22, "ms": ms, "Mvoxel_per_s": vox / ms / 1e3}
    del views
    torch.cuda.empty_cache()
    # ---- C1: 2x1 grid of 256^2 uint16 ----
    views, stage, true = synthetic.make_grid((1, 2), (256, 256), (51, 51), np.uint16, jitter=2, seed=1)
    osp = geometry.union_stack_props([v.bb() for v in views], true, views[0].spacing)
    plan = fusion.FusionPlan(views, true, osp)
    ms = timed(plan.run, n=20)
    out["c1_fusion"] = {"ms": ms, "Mvoxel_per_s": plan.out_voxels / ms / 1e3}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
