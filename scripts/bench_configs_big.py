#!/usr/bin/env python
"""Ad-hoc measurements of the large BASELINE configs (not the bench.py contract):
  C3  4x4x2 grid of (256,512,512) uint16 tiles, content-weighted fusion of the WHOLE stack
  C4  4 views of (512,1024,1024) uint16 with preset affines, content-weighted fusion
  C5  one GPU's shard of the 8x8 grid of (512,2048,2048) uint16 tiles: 2x4 tiles, blend-only fusion
Sizes can be scaled down with MVS_BIG_SCALE=2 (halves every tile axis)."""

import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiview_stitcher_b200 import fusion, geometry, synthetic  # noqa: E402
from multiview_stitcher_b200.fusion import DeviceView  # noqa: E402

SC = int(os.environ.get("MVS_BIG_SCALE", "1"))


def wall(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = fn()
    torch.cuda.synchronize()
    return r, time.perf_counter() - t0


def main():
    out = {}
    which = set(sys.argv[1:]) or {"c3", "c4", "c5"}
    if "c3" in which:
        grid, tile, ov = (2, 4, 4), (256 // SC, 512 // SC, 512 // SC), (26 // SC, 51 // SC, 51 // SC)
        views, stage, true = synthetic.make_grid(grid, tile, ov, np.uint16, jitter=2, seed=0)
        osp = geometry.union_stack_props([v.bb() for v in views], true, views[0].spacing)
        vox = int(np.prod([osp["shape"][d] for d in "zyx"]))
        (res, _), dt = wall(lambda: fusion.fuse(views, true, output_stack_properties=osp, weights_func=fusion.content_based,
                                                 output_on_backend=True))
        (res, _), dt = wall(lambda: fusion.fuse(views, true, output_stack_properties=osp, weights_func=fusion.content_based,
                                                 output_on_backend=True))
        out["c3_content_full"] = {"out_shape": [osp["shape"][d] for d in "zyx"], "s": dt, "Mvoxel_per_s": vox / dt / 1e6,
                                  "nonzero_frac": float((res[::4, ::4, ::4].to(torch.int32) > 0).float().mean())}
        del views, res
        torch.cuda.empty_cache()
    if "c4" in which:
        # 4 views around the y axis (0 / 90 / 180 / 270 degrees) with a +-2 degree tilt and 0.5 % scale
        shape = (512 // SC, 1024 // SC, 1024 // SC)
        spacing = {"z": 2.0, "y": 1.0, "x": 1.0}
        ext = np.array([shape[0] * 2.0, shape[1] * 1.0, shape[2] * 1.0])
        centre = ext / 2
        views, params = [], []
        for k in range(4):
            t = synthetic.make_tile(shape, (0, 0, 0), np.uint16, seed=k)
            views.append(DeviceView(t, {"z": 0.0, "y": 0.0, "x": 0.0}, spacing))
            a = np.deg2rad(90.0 * k)
            tilt = np.deg2rad(2.0 if k % 2 else -2.0)
            # rotation about y (mixes z and x), tilt about x, in (z, y, x) order
            ry = np.array([[np.cos(a), 0, -np.sin(a)], [0, 1, 0], [np.sin(a), 0, np.cos(a)]])
            rx = np.array([[np.cos(tilt), np.sin(tilt), 0], [-np.sin(tilt), np.cos(tilt), 0], [0, 0, 1]])
            m = ry @ rx @ np.diag([1.0, 1.005, 0.995])
            p = np.eye(4)
            p[:3, :3] = m
            p[:3, 3] = centre - m @ centre
            params.append(p)
        bbs = [v.bb() for v in views]
        osp = geometry.union_stack_props(bbs, params, {"z": 2.0, "y": 1.0, "x": 1.0})
        vox = int(np.prod([osp["shape"][d] for d in "zyx"]))
        (res, _), dt = wall(lambda: fusion.fuse(views, params, output_stack_properties=osp, weights_func=fusion.content_based,
                                                 output_on_backend=True))
        (res, _), dt = wall(lambda: fusion.fuse(views, params, output_stack_properties=osp, weights_func=fusion.content_based,
                                                 output_on_backend=True))
        out["c4_content_full"] = {"out_shape": [osp["shape"][d] for d in "zyx"], "s": dt, "Mvoxel_per_s": vox / dt / 1e6,
                                  "nonzero_frac": float((res[::4, ::4, ::4].to(torch.int32) > 0).float().mean())}
        plan = fusion.FusionPlan(views, params, osp)
        for _ in range(2):
            plan.run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            plan.run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        out["c4_blend_full"] = {"ms": ms, "Mvoxel_per_s": plan.out_voxels / ms / 1e3, "kernel": "fuse_kernel<3,1,WAVG> (general affine)"}
        plan.close()
        del views, res
        torch.cuda.empty_cache()
    if "c5" in which:
        grid, tile, ov = (1, 2, 4), (512 // SC, 2048 // SC, 2048 // SC), (51 // SC, 205 // SC, 205 // SC)
        views, stage, true = synthetic.make_grid(grid, tile, ov, np.uint16, jitter=2, seed=0)
        osp = geometry.union_stack_props([v.bb() for v in views], true, views[0].spacing)
        plan = fusion.FusionPlan(views, true, osp)
        for _ in range(2):
            plan.run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            plan.run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        b = plan.algorithmic_bytes()
        out["c5_shard_fusion"] = {"tiles": 8, "tile": list(tile), "out_shape": [osp["shape"][d] for d in "zyx"], "ms": ms,
                                  "Mvoxel_per_s": plan.out_voxels / ms / 1e3, "GB_per_s": b / ms / 1e6,
                                  "frac_of_peak": b / ms / 1e6 / 6553.3}
        plan.close()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
