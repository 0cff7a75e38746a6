"""Multi-view deconvolution fusion behind the reference's ``fusion_func`` hook.

``multi_view_deconvolution`` has the signature, defaults and ``required_overlap`` attribute of
``fusion.mv_deconv.multi_view_deconvolution`` (fusion/mv_deconv.py:251-527), so
``fusion.fuse(..., fusion_func=multi_view_deconvolution, fusion_func_kwargs=...)`` calls it
unchanged.  The iterative part -- two PSF-sized convolutions per view and iteration with the
Richardson-Lucy quotient / update fused into their epilogues -- runs on the GPU
(csrc/deconv.cu).  The PSFs and compound back-projection kernels are a few hundred numbers;
like the reference ("always on CPU", :173-185) they are built on the host with the same
scipy calls.
"""

from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from ._lib import EngineError

PSF_TYPES = ("EFFICIENT_BAYESIAN", "OPTIMIZATION_I", "OPTIMIZATION_II", "INDEPENDENT")


def _norm(kernel):
    kernel = kernel.astype(np.float64)
    s = kernel.sum()
    if s > 0:
        kernel = kernel / s
    return kernel.astype(np.float32)


def make_gaussian_psf(sigma, ndim=None, shape=None):
    """Normalised Gaussian PSF (mv_deconv.py:94-129): a delta filtered by
    ``scipy.ndimage.gaussian_filter``; default extent ``ceil(6 sigma) | 1``."""
    from scipy.ndimage import gaussian_filter

    sigma = np.atleast_1d(sigma)
    if sigma.size == 1 and ndim is not None:
        sigma = np.full(ndim, float(sigma[0]))
    if shape is None:
        shape = tuple(int(np.ceil(6.0 * s)) | 1 for s in sigma)
    psf = np.zeros(shape, dtype=np.float32)
    psf[tuple(s // 2 for s in shape)] = 1.0
    return _norm(gaussian_filter(psf, sigma=sigma.tolist()))


def estimate_psf(spacing, na=0.8, wavelength_um=0.5):
    """Gaussian PSF from pixel spacing and objective parameters (mv_deconv.py:132-166)."""
    lateral = 0.5 * wavelength_um / na
    axial = 2.0 * wavelength_um / (na**2)
    return make_gaussian_psf([max(0.5, (axial if d == "z" else lateral) / float(sp)) for d, sp in spacing.items()])


def compound_kernel(v_idx, psfs, psf_type):
    """Back-projection kernel of view ``v_idx`` (mv_deconv.py:173-245)."""
    from scipy.ndimage import convolve

    n_views = len(psfs)
    psf_type = getattr(psf_type, "value", str(psf_type))
    if psf_type not in PSF_TYPES:
        raise EngineError(f"unknown psf_type {psf_type!r} ({', '.join(PSF_TYPES)})")
    psf_v = psfs[v_idx].astype(np.float64)
    if n_views == 1 or psf_type == "INDEPENDENT":
        return _norm(np.flip(psf_v))
    if psf_type == "OPTIMIZATION_II":
        return _norm(np.flip(psf_v**n_views))
    flip_v = np.flip(psf_v)
    if psf_type == "OPTIMIZATION_I":
        tmp = psf_v.copy()
        for w, psf_w in enumerate(psfs):
            if w != v_idx:
                tmp = tmp * convolve(flip_v, psf_w.astype(np.float64), mode="constant", cval=0.0)
        return _norm(np.flip(tmp))
    tmp = flip_v.copy()
    for w, psf_w in enumerate(psfs):
        if w == v_idx:
            continue
        pw = psf_w.astype(np.float64)
        tmp = tmp * convolve(convolve(flip_v, pw, mode="constant", cval=0.0), np.flip(pw), mode="constant", cval=0.0)
    return _norm(tmp)


def prepare_kernels(n_views, ndim, psfs=None, psf_type="EFFICIENT_BAYESIAN", output_spacing=None, na=0.8, wavelength_um=0.5):
    """(kernels1, kernels2): the per-view PSFs padded to a common shape and their compound
    back-projection kernels (mv_deconv.py:373-415)."""
    if psfs is None:
        psf0 = estimate_psf(output_spacing, na=na, wavelength_um=wavelength_um) if output_spacing is not None else make_gaussian_psf(1.5, ndim=ndim)
        base = [psf0] * n_views
    else:
        if len(psfs) != n_views:
            raise ValueError(f"len(psfs) = {len(psfs)}, but n_views = {n_views}. Provide one PSF per view.")
        base = [_norm(np.asarray(p).astype(np.float32)) for p in psfs]
    max_shape = tuple(max(p.shape[d] for p in base) for d in range(ndim))
    padded = []
    for p in base:
        if p.shape != max_shape:
            pad = [((t - a) // 2, (t - a) - (t - a) // 2) for a, t in zip(p.shape, max_shape)]
            p = np.pad(p, pad, mode="constant")
        padded.append(_norm(p))
    return padded, [compound_kernel(v, padded, psf_type) for v in range(n_views)]


def multi_view_deconvolution(transformed_views, blending_weights, psfs=None, psf_type="EFFICIENT_BAYESIAN", n_iterations=10,
                             lambda_reg=0.0, min_value=1e-4, output_spacing=None, na=0.8, wavelength_um=0.5,
                             sample_boundary_erosion_px=0):
    """GPU ``fusion_func``: Bayesian multi-view deconvolution (fusion/mv_deconv.py:251-500).
    numpy in -> numpy out, CUDA tensors in -> CUDA tensor out."""
    import torch

    from .hooks import _back, _shape3, _to_stack

    lib = _lib.load(require_device=True)
    tv, was_np = _to_stack(transformed_views)
    bw, _ = _to_stack(blending_weights)
    if bw.shape != tv.shape:
        raise EngineError("blending_weights shape differs from transformed_views")
    n_views, spatial = tv.shape[0], tuple(tv.shape[1:])
    ndim = len(spatial)
    k1, k2 = prepare_kernels(n_views, ndim, psfs, psf_type, output_spacing, na, wavelength_um)
    kshape = k1[0].shape
    if any(s % 2 == 0 or s > 15 for s in kshape):
        raise EngineError(f"PSF extent {kshape}: the engine convolves odd extents up to 15")
    K1 = np.ascontiguousarray(np.stack(k1), dtype=np.float32)
    K2 = np.ascontiguousarray(np.stack(k2), dtype=np.float32)
    out = torch.empty(spatial, dtype=torch.float32, device="cuda")
    ksh = (ctypes.c_int32 * 3)(*((1,) * (3 - ndim) + tuple(int(s) for s in kshape)))
    _lib.check(
        lib.mvs_mv_deconvolution(
            ctypes.c_void_p(tv.data_ptr()), ctypes.c_void_p(bw.data_ptr()), n_views, _shape3(spatial), ndim,
            K1.ctypes.data_as(ctypes.c_void_p), K2.ctypes.data_as(ctypes.c_void_p), ksh, int(n_iterations),
            ctypes.c_float(float(lambda_reg)), ctypes.c_float(float(min_value)), int(sample_boundary_erosion_px),
            ctypes.c_void_p(out.data_ptr()), _lib.current_stream_ptr(),
        ),
        "mvs_mv_deconvolution",
    )
    return _back(out, was_np)


def _required_overlap_for_deconvolution(func_kwargs):
    """PSF half-width as the chunk halo (mv_deconv.py:504-524)."""
    kwargs = func_kwargs or {}
    output_spacing = kwargs.get("output_spacing", None)
    if output_spacing is not None:
        psf = estimate_psf(output_spacing, na=kwargs.get("na", 0.8), wavelength_um=kwargs.get("wavelength_um", 0.5))
        psf_size = max(psf.shape)
    else:
        psf_size = int(np.ceil(6.0 * 1.5)) | 1
    return psf_size // 2


multi_view_deconvolution.required_overlap = _required_overlap_for_deconvolution
