"""C4 (4 rotated views, general affine): TMA-staged brick kernel vs the gather kernel (MVS_FUSE_GATHER=1)."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, ".")
from multiview_stitcher_b200 import fusion, geometry, synthetic
from multiview_stitcher_b200.fusion import DeviceView
SC = int(os.environ.get("SC", "1"))
shape = (512 // SC, 1024 // SC, 1024 // SC)
spacing = {"z": 2.0, "y": 1.0, "x": 1.0}
ext = np.array([shape[0] * 2.0, shape[1] * 1.0, shape[2] * 1.0]); centre = ext / 2
views, params = [], []
for k in range(4):
    t = synthetic.make_tile_field(shape, (0.0, 0.0, 0.0), np.uint16, seed=10 + k, tile_id=k)
    views.append(DeviceView(t, {"z": 0.0, "y": 0.0, "x": 0.0}, spacing))
    a = np.deg2rad(90.0 * k); tilt = np.deg2rad(2.0 if k % 2 else -2.0)
    ry = np.array([[np.cos(a), 0, -np.sin(a)], [0, 1, 0], [np.sin(a), 0, np.cos(a)]])
    rx = np.array([[np.cos(tilt), np.sin(tilt), 0], [-np.sin(tilt), np.cos(tilt), 0], [0, 0, 1]])
    m = ry @ rx @ np.diag([1.0, 1.005, 0.995])
    p = np.eye(4); p[:3, :3] = m; p[:3, 3] = centre - m @ centre
    params.append(p)
bbs = [v.bb() for v in views]
osp = geometry.union_stack_props(bbs, params, spacing)
res, outs = {}, {}
for mode in ("staged", "gather"):
    if mode == "gather": os.environ["MVS_FUSE_GATHER"] = "1"
    else: os.environ.pop("MVS_FUSE_GATHER", None)
    for order in (1, 0):
        plan = fusion.FusionPlan(views, params, osp, interpolation_order=order)
        for _ in range(2): plan.run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): plan.run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        res[f"{mode}_order{order}"] = {"ms": ms, "frac": plan.algorithmic_bytes() / ms / 1e6 / 6553.3, "launches": plan.launches_per_run}
        outs[(mode, order)] = plan.out.clone()
        plan.close()
for order in (1, 0):
    d = (outs[("staged", order)].to(torch.int32) - outs[("gather", order)].to(torch.int32)).abs()
    res[f"diff_order{order}"] = {"max": int(d.max()), "n_gt1": int((d > 1).sum()), "n_diff": int((d > 0).sum()), "n": d.numel(),
                                 "nonzero_frac": float((outs[("gather", order)].to(torch.int32) > 0).float().mean())}
print(json.dumps(res, indent=1))
json.dump(res, open("gpurun_out/probe_c4.json", "w"), indent=1)
