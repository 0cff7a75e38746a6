#!/usr/bin/env python
"""Diagnostic: a C3 (3-D) pair whose recovered shift is off the synthetic ground
truth -- engine details next to the oracle's for the same crops."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiview_stitcher_b200 import registration, synthetic
from oracle import registration as oreg

grid, tile, ov = (2, 4, 4), (256, 512, 512), (26, 51, 51)
views, stage, true = synthetic.make_grid(grid, tile, ov, np.uint16, jitter=2, seed=0)
idx = list(np.ndindex(*grid))
pairs = [(i, i + 1) for i, t in enumerate(idx) if t[2] + 1 < grid[2]][:8]
fixed = [views[a].tensor[:, :, -ov[2]:].to(torch.float32).contiguous() for a, b_ in pairs]
moving = [views[b_].tensor[:, :, : ov[2]].to(torch.float32).contiguous() for a, b_ in pairs]
res = registration.register_pairs(fixed, moving, return_details=True)
true_t = np.array([t[:3, 3] for t in true])
bad = []
for k, (r, (a, b)) in enumerate(zip(res, pairs)):
    err = np.abs(r["affine_matrix"][:3, 3] + (true_t[b] - true_t[a])).max()
    print("pair", k, "t", r["affine_matrix"][:3, 3], "truth", -(true_t[b] - true_t[a]), "quality", r["quality"])
    if err > 0.25:
        bad.append(k)
print("off ground truth:", bad)
for k in bad[:1]:
    r = res[k]
    print(" engine shift cands", [list(map(float, s)) for s in r["shift_candidates"]])
    print(" engine ssim", [round(float(x), 6) for x in r["ssim"]])
    t0 = time.time()
    o = oreg.phase_correlation_registration(fixed[k].cpu().numpy(), moving[k].cpu().numpy(), return_details=True)
    print(" oracle t", o["affine_matrix"][:3, 3], "quality", o["quality"], "in", round(time.time() - t0), "s")
    print(" oracle shift cands", [list(map(float, s)) for s in o["shift_candidates"]])
    print(" oracle ssim", [round(float(x), 6) for x in o["ssim"]])
