#!/usr/bin/env python
"""Diagnostic: the engine's final correlation surfaces (MVS_PC_DEBUG_STORE=1) vs numpy."""
import os, sys, ctypes
import numpy as np
import scipy.fft as sfft
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from multiview_stitcher_b200 import registration, synthetic, _lib

views, stage, true = synthetic.make_grid(bench.GRID, bench.TILE, bench.OVERLAP, np.float32, jitter=2, seed=0)
pairs = bench.c2_pairs()
tiles = [v.tensor for v in views]
fixed, moving = bench.pair_crops(tiles, pairs)
fixed = [f.contiguous() for f in fixed]; moving = [m.contiguous() for m in moving]
lib = _lib.load(require_device=True)
def rescaled(a):
    a = a.astype(np.float32)
    return (a - a.min()) / np.float32(float(a.max()) - float(a.min()))
nbad = 0
for k in range(len(pairs)):
    f, m = fixed[k].cpu().numpy(), moving[k].cpu().numpy()
    plan = registration.PhaseCorrPlan(f.shape, 1, 10)
    plan.load_pairs([fixed[k]], [moving[k]])
    peaks, updft = plan.correlate()
    N = f.size
    h = np.zeros(2 * N, dtype=np.float32)
    _lib.check(lib.mvs_pc_debug_copy(plan._h, 1, 0, h.ctypes.data_as(ctypes.c_void_p)), "debug_copy")
    W = (h[0::2] + 1j * h[1::2]).reshape(f.shape)
    pk = []
    for surf in (W.real, W.imag):
        q = np.array(np.unravel_index(np.argmax(np.abs(surf)), surf.shape))
        pk.append(list(np.where(q > np.array(f.shape) // 2, q - np.array(f.shape), q)))
    eng = [list(map(int, peaks[0, s, 1:])) for s in (0, 1)]
    ok = eng == [list(map(int, p_)) for p_ in pk]
    if not ok or k < 2:
        F0, F1 = sfft.fftn(rescaled(f).astype(np.float64)), sfft.fftn(rescaled(m).astype(np.float64))
        P = F0 * F1.conj()
        Pn = P / np.maximum(np.abs(P), 100 * np.finfo(np.float32).eps)
        ref = sfft.ifftn(P + 1j * Pn) * N
        err = np.abs(W - ref)
        im = np.abs(W.imag)
        top = np.argsort(im.ravel())[-3:][::-1]
        print("pair", k, f.shape, "engine keys", eng, "argmax of engine surface", pk, "OK" if ok else "MISMATCH",
              "| surface rel err", err.max() / np.abs(ref).max(), "top-3 |Im|:", [(int(i // f.shape[1]), int(i % f.shape[1]), float(im.ravel()[i])) for i in top],
              "ref |Im| max", float(np.abs(ref.imag).max()))
    nbad += not ok
    plan.close()
print("mismatching pairs:", nbad)
