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


# --- host mirror of csrc/synth.cu (numpy, integer-only: bit-identical to the kernel) ---


def _mix32(h):
    h = h.astype(np.uint32)
    h ^= h >> np.uint32(16)
    h = h * np.uint32(0x85EBCA6B)
    h ^= h >> np.uint32(13)
    h = h * np.uint32(0xC2B2AE35)
    h ^= h >> np.uint32(16)
    return h


def _hash3(seed, z, y, x):
    def lohi(v, k):
        v = v.astype(np.int64)
        lo = (v & 0xFFFFFFFF).astype(np.uint32)
        hi = ((v.view(np.uint64) >> np.uint64(32)) & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        return lo ^ (hi * np.uint32(k))

    h = _mix32(np.uint32(seed) ^ np.uint32(0x9E3779B9) + np.zeros(1, np.uint32))
    h = _mix32(h ^ lohi(z, 0x27D4EB2F))
    h = _mix32(h ^ lohi(y, 0x165667B1))
    h = _mix32(h ^ lohi(x, 0xD3A2646C))
    return h


def _lattice_octave(seed, z, y, x, shift, amp_mask):
    cell = 1 << shift
    cz, cy, cx = z >> shift, y >> shift, x >> shift
    wz, wy, wx = (z - (cz << shift)).astype(np.uint64), (y - (cy << shift)).astype(np.uint64), (x - (cx << shift)).astype(np.uint64)
    c = np.uint64(cell)
    acc = np.zeros(np.broadcast(z, y, x).shape, dtype=np.uint64)
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                h = (_hash3(seed, cz + dz, cy + dy, cx + dx) & np.uint32(amp_mask)).astype(np.uint64)
                w = ((wz if dz else c - wz) * (wy if dy else c - wy) * (wx if dx else c - wx)) & np.uint64(0xFFFFFFFF)
                acc = acc + h * w
    return (acc >> np.uint64(3 * shift)).astype(np.uint32)


def ground_truth(shape, origin, dtype, seed=0):
    """The tile ``make_tile(shape, origin, dtype, seed)`` generates, computed on the host."""
    ndim = len(shape)
    shp = [1] * (3 - ndim) + [int(s) for s in shape]
    org = [0] * (3 - ndim) + [int(o) for o in origin]
    with np.errstate(over="ignore"):
        z = (np.arange(shp[0], dtype=np.int64) + org[0])[:, None, None]
        y = (np.arange(shp[1], dtype=np.int64) + org[1])[None, :, None]
        x = (np.arange(shp[2], dtype=np.int64) + org[2])[None, None, :]
        z, y, x = np.broadcast_arrays(z, y, x)
        v = _lattice_octave(seed, z, y, x, 4, 4095)
        v = v + _lattice_octave(seed + 1, z, y, x, 2, 1023)
        v = v + (_hash3(seed + 2, z, y, x) & np.uint32(255))
    v = v.reshape([int(s) for s in shape])
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return v.astype(np.float32) * np.float32(1.0 / 8192.0)
    if dtype == np.uint16:
        return v.astype(np.uint16)
    return (v >> np.uint32(5)).astype(np.uint8)
