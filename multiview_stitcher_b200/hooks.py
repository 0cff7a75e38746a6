"""GPU implementations of the reference's post-resample extension hooks.

``fusion_func`` / ``weights_func`` callables with the reference's names,
signatures and semantics (docs/extension_api_fusion.md; fusion/_core.py:42-131,
weights.py:22-74, :325-345) operating on ``(V, *chunk)`` float32 stacks with NaN
outside the views.  They can be handed to the reference's ``fusion.fuse(...,
fusion_func=..., weights_func=...)`` unchanged: ``fuse_np`` then calls them per
chunk with host arrays (uploaded here) -- or they are used by this package's own
fused path, which keeps everything on the device.

Input numpy -> output numpy; input CUDA tensors -> output CUDA tensors.
"""

from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from ._lib import EngineError


def _to_stack(a):
    """-> (CUDA float32 contiguous tensor, was_numpy)"""
    import torch

    if isinstance(a, torch.Tensor):
        return a.to("cuda", dtype=torch.float32).contiguous(), False
    if isinstance(a, (list, tuple)):
        a = np.stack([np.asarray(x) for x in a])
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to("cuda"), True


def _back(t, was_numpy):
    return t.cpu().numpy() if was_numpy else t


def _shape3(shape):
    shape = tuple(int(s) for s in shape)
    return (ctypes.c_int32 * 3)(*((1,) * (3 - len(shape)) + shape))


def gaussian_kernel1d(sigma, truncate=4.0):
    """scipy.ndimage's normalised Gaussian kernel (``_gaussian_kernel1d`` for
    order 0) as one-sided weights[0..radius] in float64."""
    sd = float(sigma)
    radius = int(truncate * sd + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sd * sd) * x**2)
    phi = phi / phi.sum()
    return np.ascontiguousarray(phi[radius:], dtype=np.float64), radius


def normalize_weights(weights):
    """weights.normalize_weights (weights.py:325-345)."""
    lib = _lib.load(require_device=True)
    w, was_np = _to_stack(weights)
    w = w.clone() if not was_np else w
    V = w.shape[0]
    N = w[0].numel()
    _lib.check(lib.mvs_normalize_weights(ctypes.c_void_p(w.data_ptr()), V, N, _lib.current_stream_ptr()), "mvs_normalize_weights")
    return _back(w, was_np)


def _fuse_stack(mode, transformed_views, blending_weights=None, fusion_weights=None):
    import torch

    lib = _lib.load(require_device=True)
    tv, was_np = _to_stack(transformed_views)
    V = tv.shape[0]
    N = tv[0].numel()
    bw = fw = None
    if blending_weights is not None:
        bw, _ = _to_stack(blending_weights)
        if bw.shape != tv.shape:
            raise EngineError("blending_weights shape differs from transformed_views")
    if fusion_weights is not None:
        fw, _ = _to_stack(fusion_weights)
        if fw.shape != tv.shape:
            raise EngineError("fusion_weights shape differs from transformed_views")
    out = torch.empty(tv.shape[1:], dtype=torch.float32, device="cuda")
    _lib.check(
        lib.mvs_fuse_stack(
            ctypes.c_void_p(tv.data_ptr()),
            ctypes.c_void_p(bw.data_ptr() if bw is not None else 0),
            ctypes.c_void_p(fw.data_ptr() if fw is not None else 0),
            V, N, mode, ctypes.c_void_p(out.data_ptr()), _lib.current_stream_ptr(),
        ),
        "mvs_fuse_stack",
    )
    return _back(out, was_np)


def weighted_average_fusion(transformed_views, blending_weights, fusion_weights=None):
    """fusion.weighted_average_fusion (fusion/_core.py:61-94)."""
    return _fuse_stack(_lib.MVS_FUSE_WAVG, transformed_views, blending_weights, fusion_weights)


def max_fusion(transformed_views):
    """fusion.max_fusion (fusion/_core.py:42-58)."""
    return _fuse_stack(_lib.MVS_FUSE_MAX, transformed_views)


def simple_average_fusion(transformed_views):
    """fusion.simple_average_fusion (fusion/_core.py:97-131)."""
    return _fuse_stack(_lib.MVS_FUSE_MEAN, transformed_views)


def content_based(transformed_views, blending_weights, sigma_1=5, sigma_2=11):
    """weights.content_based (weights.py:22-74): Preibisch content-based fusion
    weights ``W = G_s2 * (I - G_s1 * I)^2`` with NaN-aware Gaussians, normalised
    over the views."""
    import torch

    lib = _lib.load(require_device=True)
    tv, was_np = _to_stack(transformed_views)
    bw, _ = _to_stack(blending_weights)
    if bw.shape != tv.shape:
        raise EngineError("blending_weights shape differs from transformed_views")
    ndim = tv.ndim - 1
    if ndim not in (2, 3):
        raise EngineError("content_based needs (V, (z,) y, x) stacks")
    w1, r1 = gaussian_kernel1d(sigma_1)
    w2, r2 = gaussian_kernel1d(sigma_2)
    out = torch.empty_like(tv)
    _lib.check(
        lib.mvs_content_based(
            ctypes.c_void_p(tv.data_ptr()), ctypes.c_void_p(bw.data_ptr()), tv.shape[0], _shape3(tv.shape[1:]), ndim,
            w1.ctypes.data_as(ctypes.c_void_p), r1, w2.ctypes.data_as(ctypes.c_void_p), r2,
            ctypes.c_void_p(out.data_ptr()), _lib.current_stream_ptr(),
        ),
        "mvs_content_based",
    )
    return _back(out, was_np)


# halo the reference's fuse() reads from the hook (misc_utils.py:69-105)
content_based.required_overlap = lambda kwargs: 2 * kwargs["sigma_2"]


def _clamp_overlap(overlap, output_chunksize):
    """weights._clamp_overlap (weights.py:514-524)."""
    sdims = sorted(output_chunksize.keys())[::-1]
    if not isinstance(overlap, dict):
        overlap = {dim: int(overlap) for dim in sdims}
    return {dim: min(overlap[dim], output_chunksize[dim]) for dim in sdims}


def content_based_dct(transformed_views, dct_size=32, exponent=1.0, otf_support_fraction=0.5, output_chunksize=None):
    """weights.content_based_dct (weights.py:77-290) on the GPU: DCT Shannon-entropy quality per
    block of ``dct_size`` voxels, normalised over the views and interpolated back to voxel
    resolution.  Same arguments and defaults as the reference; block edges up to 32."""
    import torch

    lib = _lib.load(require_device=True)
    tv, was_np = _to_stack(transformed_views)
    spatial = tuple(tv.shape[1:])
    ndim = len(spatial)
    sdims = ["z", "y", "x"][-ndim:]
    sizes = tuple(dct_size[d] for d in sdims) if isinstance(dct_size, dict) else (dct_size,) * ndim
    if output_chunksize is not None:
        sizes = tuple(int(min(ds, output_chunksize[dim], s)) for ds, dim, s in zip(sizes, sdims, spatial))
    else:
        sizes = tuple(int(min(ds, s)) for ds, s in zip(sizes, spatial))
    if max(sizes) > 32:
        raise EngineError(f"content_based_dct: dct_size {sizes} exceeds the engine's block limit of 32")
    r_o = -1.0 if otf_support_fraction is None else float(otf_support_fraction) * min(sizes)
    out = torch.empty_like(tv)
    blk = (ctypes.c_int32 * 3)(*((1,) * (3 - ndim) + sizes))
    _lib.check(
        lib.mvs_content_based_dct(ctypes.c_void_p(tv.data_ptr()), tv.shape[0], _shape3(spatial), ndim, blk,
                                  ctypes.c_float(r_o), ctypes.c_float(float(exponent)), ctypes.c_void_p(out.data_ptr()),
                                  _lib.current_stream_ptr()),
        "mvs_content_based_dct",
    )
    return _back(out, was_np)


content_based_dct.required_overlap = lambda kwargs: _clamp_overlap(kwargs["dct_size"], kwargs["output_chunksize"])


def gaussian_filter(stack, sigma):
    """scipy.ndimage.gaussian_filter(mode="reflect") of every volume of a stack."""
    import torch

    lib = _lib.load(require_device=True)
    t, was_np = _to_stack(stack)
    ndim = t.ndim - 1
    w, r = gaussian_kernel1d(sigma)
    out = torch.empty_like(t)
    _lib.check(
        lib.mvs_gaussian_filter(
            ctypes.c_void_p(t.data_ptr()), ctypes.c_void_p(out.data_ptr()), t.shape[0], _shape3(t.shape[1:]), ndim,
            w.ctypes.data_as(ctypes.c_void_p), r, _lib.current_stream_ptr(),
        ),
        "mvs_gaussian_filter",
    )
    return _back(out, was_np)
