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

# ---- integer peaks of both correlation surfaces vs numpy (complex128) on the same crops ----
import scipy.fft as sfft
def rescaled(a):
    a = a.astype(np.float32)
    return (a - a.min()) / np.float32(float(a.max()) - float(a.min()))
for k in sorted(set(bad[:2] + [0, 1, 2])):
    f, m = fixed[k].cpu().numpy(), moving[k].cpu().numpy()
    plan = registration.PhaseCorrPlan(f.shape, 1, 10)
    plan.load_pairs([fixed[k]], [moving[k]])
    peaks, updft = plan.correlate()
    F0, F1 = sfft.fftn(rescaled(f).astype(np.float64)), sfft.fftn(rescaled(m).astype(np.float64))
    P = F0 * F1.conj()
    Pn = P / np.maximum(np.abs(P), 100 * np.finfo(np.float32).eps)
    out = []
    for slot, prod in ((0, P), (1, Pn)):
        cc = np.abs(sfft.ifftn(prod))
        peak = np.array(np.unravel_index(np.argmax(cc), cc.shape))
        wrapped = np.where(peak > np.array(f.shape) // 2, peak - np.array(f.shape), peak)
        srt = np.sort(cc.ravel())
        out.append((slot, list(peaks[0, slot, 1:]), list(wrapped), float(srt[-1] / srt[-2])))
    print("pair", k, f.shape, "engine/numpy peaks + numpy peak ratio:", out)
    plan.close()

# ---- where does a wrong peak come from?  engine buffers vs numpy ----
import ctypes
from multiview_stitcher_b200 import _lib
lib = _lib.load(require_device=True)
for k in sorted(set(bad[:1] + [1])):
    f, m = fixed[k].cpu().numpy(), moving[k].cpu().numpy()
    plan = registration.PhaseCorrPlan(f.shape, 1, 10)
    plan.load_pairs([fixed[k]], [moving[k]])
    peaks, updft = plan.correlate()
    N = f.size
    bufs = []
    for which in (0, 1):
        h = np.zeros(2 * N, dtype=np.float32)
        _lib.check(lib.mvs_pc_debug_copy(plan._h, which, 0, h.ctypes.data_as(ctypes.c_void_p)), "debug_copy")
        bufs.append((h[0::2] + 1j * h[1::2]).reshape(f.shape))
    Pe, Qy = bufs
    F0, F1 = sfft.fftn(rescaled(f).astype(np.float64)), sfft.fftn(rescaled(m).astype(np.float64))
    P = F0 * F1.conj()
    Pn = P / np.maximum(np.abs(P), 100 * np.finfo(np.float32).eps)
    scale = 1.0 / float(N) ** 2 if os.environ.get("MVS_PC_SCALE_N2") else 1.0
    Qref = scale * P + 1j * Pn
    Qy_ref = sfft.ifft(Qref, axis=0) * f.shape[0]
    print("pair", k, "P rel err", np.abs(Pe - P).max() / np.abs(P).max(), "nonfinite", (~np.isfinite(Pe)).sum(),
          "| Q(after inverse y) rel err", np.abs(Qy - Qy_ref).max() / np.abs(Qy_ref).max(), "nonfinite", (~np.isfinite(Qy)).sum())
    cc = sfft.ifft(Qy.astype(np.complex128), axis=1)
    for slot, surf in ((0, cc.real), (1, cc.imag)):
        pk = np.array(np.unravel_index(np.argmax(np.abs(surf)), surf.shape))
        print("   slot", slot, "engine peak", list(peaks[0, slot, 1:]), "numpy-on-engine-Q peak", list(np.where(pk > np.array(f.shape) // 2, pk - np.array(f.shape), pk)))
    bad_rows = np.where(np.abs(Qy - Qy_ref).max(axis=1) > 1e-3 * np.abs(Qy_ref).max())[0]
    bad_cols = np.where(np.abs(Qy - Qy_ref).max(axis=0) > 1e-3 * np.abs(Qy_ref).max())[0]
    print("   rows/cols of Q with large error:", bad_rows[:20], len(bad_rows), bad_cols[:20], len(bad_cols))
    plan.close()
