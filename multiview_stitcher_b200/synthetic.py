"""Seeded synthetic tile grids for tests and benchmarks.

Tiles are cut on the GPU from an integer value-noise ground truth
(``csrc/synth.cu``) at integer pixel positions ``index * (tile - overlap) +
jitter``; the *stage* transform handed to the engine omits the jitter, so the
true pairwise shift is the jitter difference.
"""

from __future__ import annotations

import ctypes

import numpy as np

from . import _lib, geometry


def grid_layout(grid, tile_shape, overlap, jitter=2, seed=0):
    """Integer tile origins (true, with jitter) and stage origins (nominal).

    grid / tile_shape / overlap are (z,)y,x tuples.  Returns
    ``(true_origins [n, ndim] int64, stage_origins [n, ndim] float64, index list)``.
    """
    ndim = len(tile_shape)
    rng = np.random.default_rng(seed)
    pitch = np.array(tile_shape) - np.array(overlap)
    idx = list(np.ndindex(*grid))
    stage = np.array([np.array(i) * pitch for i in idx], dtype=np.float64)
    jit = rng.integers(-jitter, jitter + 1, size=(len(idx), ndim)) if jitter else np.zeros((len(idx), ndim), dtype=np.int64)
    true = stage.astype(np.int64) + jit
    return true, stage, idx


def make_tile(shape, origin, dtype, seed=0, device="cuda"):
    """One tile of the synthetic ground truth as a CUDA tensor."""
    import torch

    lib = _lib.load(require_device=True)
    ndim = len(shape)
    tdt = {np.dtype(np.uint8): torch.uint8, np.dtype(np.uint16): torch.uint16, np.dtype(np.float32): torch.float32}[np.dtype(dtype)]
    t = torch.empty(tuple(int(s) for s in shape), dtype=tdt, device=device)
    shp = (ctypes.c_int32 * 3)(*([1] * (3 - ndim) + [int(s) for s in shape]))
    strd = (ctypes.c_int64 * 3)(*([0] * (3 - ndim) + [int(s) for s in t.stride()]))
    org = (ctypes.c_int64 * 3)(*([0] * (3 - ndim) + [int(o) for o in origin]))
    st = lib.mvs_synth_tile(ctypes.c_void_p(t.data_ptr()), _lib.mvs_dtype(dtype), shp, strd, org, ctypes.c_uint32(seed), _lib.current_stream_ptr())
    _lib.check(st, "mvs_synth_tile")
    return t


def make_grid(grid, tile_shape, overlap, dtype, jitter=2, seed=0, spacing=None, device="cuda"):
    """Device-resident tile grid.  Returns (views, stage_params, true_params):
    ``views`` are DeviceViews placed at their *stage* origin; ``stage_params``
    are identities (what an unregistered dataset has); ``true_params`` are the
    translations that register them exactly."""
    from .fusion import DeviceView

    ndim = len(tile_shape)
    dims = geometry.spatial_dims(ndim)
    if spacing is None:
        spacing = {d: 1.0 for d in dims}
    sp = np.array([spacing[d] for d in dims])
    true, stage, idx = grid_layout(grid, tile_shape, overlap, jitter, seed)
    views, stage_params, true_params = [], [], []
    for t_org, s_org in zip(true, stage):
        tens = make_tile(tile_shape, t_org, dtype, seed, device)
        views.append(DeviceView(tens, dict(zip(dims, s_org * sp)), spacing))
        stage_params.append(np.eye(ndim + 1))
        p = np.eye(ndim + 1)
        p[:ndim, ndim] = (t_org - s_org) * sp
        true_params.append(p)
    return views, stage_params, true_params
