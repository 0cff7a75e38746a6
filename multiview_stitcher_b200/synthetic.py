"""Seeded synthetic tile grids for tests and benchmarks.

Two ground truths, both generated on the GPU (``csrc/synth.cu``) with a numpy
host mirror:

* ``subpixel=True`` (benchmarks, BASELINE-size tests, smoke): a band-limited
  analytic field -- a sum of plane waves -- evaluated at ``index * (tile -
  overlap) + jitter`` with jitter ~ U(-2, 2) px per axis, plus per-tile hash
  noise.  Tiles sit at FRACTIONAL positions, so interpolation fractions,
  sub-pixel peak refinement and the SSIM tie-breaking are all exercised.
* ``subpixel=False``: an integer value-noise lattice sampled at integer
  positions (overlapping tiles agree bit for bit; the fused stack reproduces
  the ground truth exactly).

The *stage* transform handed to the engine omits the jitter, so the true
pairwise shift is the jitter difference.
"""

from __future__ import annotations

import ctypes

import numpy as np

from . import _lib, geometry


def grid_layout(grid, tile_shape, overlap, jitter=2, seed=0, subpixel=False):
    """Tile origins (true, with jitter) and stage origins (nominal), in pixels.

    grid / tile_shape / overlap are (z,)y,x tuples.  Returns ``(true_origins
    [n, ndim], stage_origins [n, ndim] float64, index list)``; true origins are
    int64 (integer jitter) or float64 with jitter ~ U(-jitter, jitter)
    (``subpixel=True``).
    """
    ndim = len(tile_shape)
    rng = np.random.default_rng(seed)
    pitch = np.array(tile_shape) - np.array(overlap)
    idx = list(np.ndindex(*grid))
    stage = np.array([np.array(i) * pitch for i in idx], dtype=np.float64)
    if subpixel:
        jit = rng.uniform(-jitter, jitter, size=(len(idx), ndim)) if jitter else np.zeros((len(idx), ndim))
        # 1/64 px grid: offsets stay exactly representable through the 10-decimal
        # rounding of transform_sim (transformation.py:72-83)
        jit = np.round(jit * 64.0) / 64.0
        return stage + jit, stage, idx
    jit = rng.integers(-jitter, jitter + 1, size=(len(idx), ndim)) if jitter else np.zeros((len(idx), ndim), dtype=np.int64)
    true = stage.astype(np.int64) + jit
    return true, stage, idx


# --- band-limited analytic field (sub-pixel tile positions) ---------------------

FIELD_TERMS = 64
FIELD_BASE = 0.5
FIELD_SIGMA = 0.12
FIELD_WAVELENGTHS = (3.5, 28.0)  # px; broadband enough that the un-normalised correlation peak is sharp
U16_SCALE = 4095.0  # 12-bit camera range


def field_terms(seed, ndim, n_terms=FIELD_TERMS):
    """Plane waves of the analytic ground truth: ``omega [K, 3]`` (cycles / px,
    z,y,x; z = 0 in 2-D), ``amp [K]``, ``phase [K]``.  Log-uniform wavelengths,
    isotropic directions, equal amplitudes, field sigma 0.12 around 0.5.  (With
    longer waves the reference's un-normalised correlation peak is broad and the
    overlap-area taper of the crop drags it towards zero by > 0.1 px.)"""
    rng = np.random.default_rng([int(seed), 0xF1E1D, ndim])
    lam = np.exp(rng.uniform(np.log(FIELD_WAVELENGTHS[0]), np.log(FIELD_WAVELENGTHS[1]), n_terms))
    d = rng.normal(size=(n_terms, ndim))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    omega = np.zeros((n_terms, 3))
    omega[:, 3 - ndim :] = d / lam[:, None]
    amp = np.ones(n_terms)
    amp *= FIELD_SIGMA / np.sqrt(0.5 * np.sum(amp**2))
    phase = rng.uniform(0, 2 * np.pi, n_terms)
    return omega, amp, phase


def _field_tables(shape, origin, seed):
    """Per-axis complex exponentials (float64) and the complex coefficients."""
    ndim = len(shape)
    omega, amp, phase = field_terms(seed, ndim)
    shp = [1] * (3 - ndim) + [int(s) for s in shape]
    org = [0.0] * (3 - ndim) + [float(o) for o in origin]
    tabs = []
    for a in range(3):
        pos = org[a] + np.arange(shp[a], dtype=np.float64)
        tabs.append(np.exp(2j * np.pi * omega[:, a : a + 1] * pos[None, :]))
    coef = amp * np.exp(1j * phase)
    return tabs, coef, shp


def _as_f2(c):
    return np.ascontiguousarray(np.stack([c.real, c.imag], axis=-1).astype(np.float32))


def make_tile_field(shape, origin, dtype, seed=0, tile_id=0, noise=0.02, device="cuda"):
    """One tile of the analytic ground truth at a (fractional) pixel origin as a
    CUDA tensor; ``noise``: peak-to-peak amplitude of the per-tile noise."""
    import torch

    lib = _lib.load(require_device=True)
    ndim = len(shape)
    dtype = np.dtype(dtype)
    tdt = {np.dtype(np.uint8): torch.uint8, np.dtype(np.uint16): torch.uint16, np.dtype(np.float32): torch.float32}[dtype]
    tabs, coef, shp = _field_tables(shape, origin, seed)
    dev = [torch.from_numpy(_as_f2(t)).to(device) for t in tabs] + [torch.from_numpy(_as_f2(coef)).to(device)]
    t = torch.empty(tuple(int(s) for s in shape), dtype=tdt, device=device)
    shp_c = (ctypes.c_int32 * 3)(*shp)
    strd = (ctypes.c_int64 * 3)(*([0] * (3 - ndim) + [int(s) for s in t.stride()]))
    scale = 1.0 if dtype == np.float32 else (U16_SCALE if dtype == np.uint16 else 255.0)
    st = lib.mvs_synth_field(
        ctypes.c_void_p(t.data_ptr()), _lib.mvs_dtype(dtype), shp_c, strd,
        *[ctypes.c_void_p(d.data_ptr()) for d in dev], len(coef), FIELD_BASE, scale, float(noise),
        ctypes.c_uint32((int(seed) * 0x9E3779B1 + int(tile_id) * 0x85EBCA77 + 1) & 0xFFFFFFFF),
        _lib.current_stream_ptr(),
    )
    _lib.check(st, "mvs_synth_field")
    torch.cuda.current_stream().synchronize()  # the tables die with this call
    return t


def field_tile_host(shape, origin, dtype, seed=0, tile_id=0, noise=0.02):
    """Host mirror of ``make_tile_field`` (float64 sums: equal to the kernel's
    float32 result up to ~1e-6 of the value range, i.e. at most 1 LSB for integer
    dtypes).  Used by the CPU arm of the benchmark and CPU-only tests."""
    ndim = len(shape)
    dtype = np.dtype(dtype)
    tabs, coef, shp = _field_tables(shape, origin, seed)
    ez, ey, ex = tabs
    rows = (coef[:, None, None] * ez[:, :, None] * ey[:, None, :]).reshape(len(coef), -1)  # [K, z*y]
    v = FIELD_BASE + (rows.T @ ex).imag.reshape(shp)
    if noise:
        nseed = (int(seed) * 0x9E3779B1 + int(tile_id) * 0x85EBCA77 + 1) & 0xFFFFFFFF
        with np.errstate(over="ignore"):
            z = np.arange(shp[0], dtype=np.int64)[:, None, None]
            y = np.arange(shp[1], dtype=np.int64)[None, :, None]
            x = np.arange(shp[2], dtype=np.int64)[None, None, :]
            z, y, x = np.broadcast_arrays(z, y, x)
            h = _hash3(nseed, z, y, x)
        v = v + np.float32(noise) * ((h >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0) - np.float32(0.5))
    v = np.clip(v, 0.0, 0.99999994).astype(np.float32).reshape([int(s) for s in shape])
    if dtype == np.float32:
        return v
    return (v * np.float32(U16_SCALE if dtype == np.uint16 else 255.0)).astype(dtype)


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


def make_grid(grid, tile_shape, overlap, dtype, jitter=2, seed=0, spacing=None, device="cuda",
              subpixel=False, noise=0.02, only=None):
    """Device-resident tile grid.  Returns (views, stage_params, true_params):
    ``views`` are DeviceViews placed at their *stage* origin; ``stage_params``
    are identities (what an unregistered dataset has); ``true_params`` are the
    translations that register them exactly.  ``subpixel``: analytic field at
    fractional jitter (see module docstring).  ``only``: indices of the tiles to
    materialise (others get ``None`` in ``views``; sharded jobs)."""
    from .fusion import DeviceView

    ndim = len(tile_shape)
    dims = geometry.spatial_dims(ndim)
    if spacing is None:
        spacing = {d: 1.0 for d in dims}
    sp = np.array([spacing[d] for d in dims])
    true, stage, idx = grid_layout(grid, tile_shape, overlap, jitter, seed, subpixel)
    views, stage_params, true_params = [], [], []
    for k, (t_org, s_org) in enumerate(zip(true, stage)):
        stage_params.append(np.eye(ndim + 1))
        p = np.eye(ndim + 1)
        p[:ndim, ndim] = (t_org - s_org) * sp
        true_params.append(p)
        if only is not None and k not in only:
            views.append(None)
            continue
        if subpixel:
            tens = make_tile_field(tile_shape, t_org, dtype, seed, k, noise, device)
        else:
            tens = make_tile(tile_shape, t_org, dtype, seed, device)
        views.append(DeviceView(tens, dict(zip(dims, s_org * sp)), spacing))
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
