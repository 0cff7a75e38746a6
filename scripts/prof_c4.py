"""ncu target: half-size C4 (4 rotated views), general-affine staged kernel."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiview_stitcher_b200 import fusion, geometry, synthetic
from multiview_stitcher_b200.fusion import DeviceView
shape = (256, 512, 512)
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
osp = geometry.union_stack_props([v.bb() for v in views], params, spacing)
plan = fusion.FusionPlan(views, params, osp)
for _ in range(3):
    plan.run()
torch.cuda.synchronize()
