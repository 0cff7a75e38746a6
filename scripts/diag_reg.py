#!/usr/bin/env python
"""Diagnostic: C2 pairs whose recovered shift differs from the synthetic ground
truth -- engine details next to the oracle's for the same crops."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from multiview_stitcher_b200 import registration, synthetic
from oracle import registration as oreg

views, stage, true = synthetic.make_grid(bench.GRID, bench.TILE, bench.OVERLAP, np.float32, jitter=2, seed=0)
pairs = bench.c2_pairs()
tiles = [v.tensor for v in views]
fixed, moving = bench.pair_crops(tiles, pairs)
fixed = [f.contiguous() for f in fixed]; moving = [m.contiguous() for m in moving]
res = registration.register_pairs(fixed, moving, return_details=True)
true_t = np.array([t[:2, 2] for t in true])
bad = []
for k, (r, (a, b, _)) in enumerate(zip(res, pairs)):
    err = np.abs(r["affine_matrix"][:2, 2] + (true_t[b] - true_t[a])).max()
    if err > 0.05:
        bad.append(k)
print("pairs off ground truth:", bad)
for k in bad[:3]:
    r = res[k]
    print("pair", k, pairs[k], "truth", -(true_t[pairs[k][1]] - true_t[pairs[k][0]]))
    print(" engine t", r["affine_matrix"][:2, 2], "shift cands", [list(map(float, s)) for s in r["shift_candidates"]])
    print(" engine ssim", [round(float(x), 6) for x in r["ssim"]])
    o = oreg.phase_correlation_registration(fixed[k].cpu().numpy(), moving[k].cpu().numpy(), return_details=True)
    print(" oracle t", o["affine_matrix"][:2, 2], "shift cands", [list(map(float, s)) for s in o["shift_candidates"]])
    print(" oracle ssim", [round(float(x), 6) for x in o["ssim"]])
    print(" t cands equal:", np.allclose(np.array(o["t_candidates"], dtype=float), np.array(r["t_candidates"], dtype=float)))
