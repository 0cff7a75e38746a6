"""Restatement of the three scikit-image functions the reference's phase
correlation calls (registration.py:383, :424-442, :543).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

scikit-image is a third-party dependency of the reference (pinned 0.26.0 in
``/root/reference/uv.lock:5029-5030``) that is NOT installed in this image and
not vendored under ``/root/reference``; what follows restates its published
algorithms:

* ``phase_cross_correlation``: Guizar-Sicairos, Thurman & Fienup, "Efficient
  subpixel image registration algorithms", Opt. Lett. 33 (2008) -- FFT cross
  correlation, integer peak, matrix-multiply upsampled DFT around the peak.
* masked variant: Padfield, "Masked object registration in the Fourier
  domain", IEEE TIP 21 (2012).
* ``structural_similarity``: Wang et al., IEEE TIP 13 (2004), uniform 7-wide
  window, sample covariance, K1=0.01, K2=0.03, border of (win-1)/2 cropped.
* ``rescale_intensity``: clip to in_range, map linearly onto out_range.

Anchors: the reference's artificial-ground-truth test
(``_tests/test_registration.py:262-336``) and exact Fourier-shift self checks
in ``tests/test_oracle_registration.py``.  Parity of SSIM values with the real
scikit-image is otherwise unpinned.
"""

from __future__ import annotations

import numpy as np
import scipy.fft as sfft
from scipy.ndimage import uniform_filter


def rescale_intensity(image, in_range, out_range=(0, 1)):
    """skimage.exposure.rescale_intensity for float input and an explicit
    (imin, imax) in_range / (omin, omax) out_range.  Output dtype follows the
    input float dtype (float32 stays float32)."""
    imin, imax = map(float, in_range)
    omin, omax = map(float, out_range)
    out_dtype = image.dtype if image.dtype.kind == "f" else np.float64
    image = np.clip(image, imin, imax)
    if imin != imax:
        image = (image - imin) / (imax - imin)
        return np.asarray(image * (omax - omin) + omin, dtype=out_dtype)
    return np.clip(image, omin, omax).astype(out_dtype)


def _upsampled_dft(data, upsampled_region_size, upsample_factor=1, axis_offsets=None):
    """Matrix-multiply DFT of ``data`` on an ``upsampled_region_size`` window
    (per axis) of the ``upsample_factor``-times finer grid, starting at
    ``axis_offsets``; kernels cast to the data's complex dtype."""
    if not hasattr(upsampled_region_size, "__iter__"):
        upsampled_region_size = [upsampled_region_size] * data.ndim
    if axis_offsets is None:
        axis_offsets = [0] * data.ndim
    im2pi = 1j * 2 * np.pi
    dim_properties = list(zip(data.shape, upsampled_region_size, axis_offsets))
    for n_items, ups_size, ax_offset in dim_properties[::-1]:
        kernel = (np.arange(ups_size) - ax_offset)[:, None] * sfft.fftfreq(
            n_items, upsample_factor
        )
        kernel = np.exp(-im2pi * kernel)
        kernel = kernel.astype(data.dtype, copy=False)
        data = np.tensordot(kernel, data, axes=(1, -1))
    return data


def _masked_phase_cross_correlation(
    reference_image, moving_image, reference_mask, moving_mask, overlap_ratio=0.3
):
    """Padfield masked normalised cross-correlation shift (masks: True =
    valid).  NOTE the reference passes ``isnan`` masks, i.e. True = INVALID
    (registration.py:435-442); this restatement reproduces what the published
    algorithm then does: valid pixels are zeroed, NaNs survive into the FFTs,
    ``fmax(., 0)`` squashes the NaN denominators to 0, the correlation is
    identically 0, every position is a maximum and the mean position gives a
    zero shift."""
    float_dtype = np.result_type(reference_image.dtype, moving_image.dtype, np.float32)
    eps = np.finfo(float_dtype).eps
    # cross_correlate_masked(moving, reference, moving_mask, reference_mask)
    fixed_image = np.array(moving_image, dtype=float_dtype)
    fixed_mask = np.array(moving_mask, dtype=bool)
    mov_image = np.array(reference_image, dtype=float_dtype)
    mov_mask = np.array(reference_mask, dtype=bool)
    axes = tuple(range(fixed_image.ndim))
    final_shape = tuple(
        fixed_image.shape[a] + mov_image.shape[a] - 1 for a in axes
    )
    final_slice = tuple(slice(0, int(sz)) for sz in final_shape)
    fast_shape = tuple(sfft.next_fast_len(final_shape[a]) for a in axes)

    def fft(x):
        return sfft.fftn(x, s=fast_shape, axes=axes)

    def ifft(x):
        return sfft.ifftn(x, s=fast_shape, axes=axes).real

    fixed_image[~fixed_mask] = 0.0
    mov_image[~mov_mask] = 0.0
    flip = tuple(slice(None, None, -1) for _ in axes)
    rot_mov = mov_image[flip]
    rot_mask = mov_mask[flip]
    with np.errstate(all="ignore"):
        fixed_fft = fft(fixed_image)
        rot_mov_fft = fft(rot_mov)
        fixed_mask_fft = fft(fixed_mask.astype(float_dtype))
        rot_mask_fft = fft(rot_mask.astype(float_dtype))
        n_overlap = ifft(rot_mask_fft * fixed_mask_fft)
        n_overlap[:] = np.round(n_overlap)
        n_overlap[:] = np.fmax(n_overlap, eps)
        mc_fixed = ifft(rot_mask_fft * fixed_fft)
        mc_mov = ifft(fixed_mask_fft * rot_mov_fft)
        numerator = ifft(rot_mov_fft * fixed_fft)
        numerator -= mc_fixed * mc_mov / n_overlap
        fixed_denom = ifft(rot_mask_fft * fft(np.square(fixed_image)))
        fixed_denom -= np.square(mc_fixed) / n_overlap
        fixed_denom[:] = np.fmax(fixed_denom, 0.0)
        mov_denom = ifft(fixed_mask_fft * fft(np.square(rot_mov)))
        mov_denom -= np.square(mc_mov) / n_overlap
        mov_denom[:] = np.fmax(mov_denom, 0.0)
        denom = np.sqrt(fixed_denom * mov_denom)
        numerator = numerator[final_slice]
        denom = denom[final_slice]
        n_overlap = n_overlap[final_slice]
        tol = 1e3 * eps * np.max(np.abs(denom), axis=axes, keepdims=True)
        nonzero = denom > tol
        out = np.zeros_like(denom, dtype=float_dtype)
        out[nonzero] = numerator[nonzero] / denom[nonzero]
        np.clip(out, -1, 1, out=out)
        thr = overlap_ratio * np.max(n_overlap, axis=axes, keepdims=True)
        out[n_overlap < thr] = 0.0
        xcorr = out
        maxima = np.stack(np.nonzero(xcorr == xcorr.max()), axis=1)
        center = np.mean(maxima, axis=0)
    shifts = center - np.array(reference_image.shape) + 1
    size_mismatch = np.array(moving_image.shape) - np.array(reference_image.shape)
    return -shifts + (size_mismatch / 2)


def phase_cross_correlation(
    reference_image,
    moving_image,
    *,
    upsample_factor=1,
    normalization="phase",
    reference_mask=None,
    moving_mask=None,
    overlap_ratio=0.3,
    disambiguate=False,
):
    """skimage.registration.phase_cross_correlation, ``space="real"``,
    ``disambiguate=False``.  Returns ``(shift, nan, nan)`` (the reference uses
    only ``[0]``, registration.py:424-442)."""
    if reference_mask is not None or moving_mask is not None:
        shift = _masked_phase_cross_correlation(
            reference_image, moving_image, reference_mask, moving_mask, overlap_ratio
        )
        return shift, np.nan, np.nan
    if reference_image.shape != moving_image.shape:
        raise ValueError("images must be same shape")
    src_freq = sfft.fftn(reference_image)
    target_freq = sfft.fftn(moving_image)
    shape = src_freq.shape
    image_product = src_freq * target_freq.conj()
    if normalization == "phase":
        eps = np.finfo(image_product.real.dtype).eps
        image_product /= np.maximum(np.abs(image_product), 100 * eps)
    elif normalization is not None:
        raise ValueError("normalization must be either phase or None")
    cross_correlation = sfft.ifftn(image_product)
    maxima = np.unravel_index(
        np.argmax(np.abs(cross_correlation)), cross_correlation.shape
    )
    midpoint = np.array([np.fix(axis_size / 2) for axis_size in shape])
    float_dtype = image_product.real.dtype
    shift = np.stack(maxima).astype(float_dtype, copy=False)
    shift[shift > midpoint] -= np.array(shape)[shift > midpoint]
    if upsample_factor > 1:
        upsample_factor = np.array(upsample_factor, dtype=float_dtype)
        shift = np.round(shift * upsample_factor) / upsample_factor
        upsampled_region_size = np.ceil(upsample_factor * 1.5)
        dftshift = np.fix(upsampled_region_size / 2.0)
        sample_region_offset = dftshift - shift * upsample_factor
        cross_correlation = _upsampled_dft(
            image_product.conj(),
            upsampled_region_size,
            upsample_factor,
            sample_region_offset,
        ).conj()
        maxima = np.unravel_index(
            np.argmax(np.abs(cross_correlation)), cross_correlation.shape
        )
        maxima = np.stack(maxima).astype(float_dtype, copy=False)
        maxima -= dftshift
        shift += maxima / upsample_factor
    for dim in range(src_freq.ndim):
        if shape[dim] == 1:
            shift[dim] = 0
    return shift, np.nan, np.nan


def structural_similarity(im1, im2, *, win_size=7, data_range=None):
    """skimage.metrics.structural_similarity with a uniform window, sample
    covariance and default constants; mean over the interior in float64."""
    if im1.shape != im2.shape:
        raise ValueError("Input images must have the same dimensions.")
    float_type = np.float32 if im1.dtype in (np.float32, np.float16) else np.float64
    K1, K2 = 0.01, 0.03
    if np.any((np.asarray(im1.shape) - win_size) < 0):
        raise ValueError("win_size exceeds image extent.")
    if win_size % 2 != 1:
        raise ValueError("Window size must be odd.")
    ndim = im1.ndim
    im1 = im1.astype(float_type, copy=False)
    im2 = im2.astype(float_type, copy=False)
    NP = win_size**ndim
    cov_norm = NP / (NP - 1)
    ux = uniform_filter(im1, size=win_size)
    uy = uniform_filter(im2, size=win_size)
    uxx = uniform_filter(im1 * im1, size=win_size)
    uyy = uniform_filter(im2 * im2, size=win_size)
    uxy = uniform_filter(im1 * im2, size=win_size)
    vx = cov_norm * (uxx - ux * ux)
    vy = cov_norm * (uyy - uy * uy)
    vxy = cov_norm * (uxy - ux * uy)
    R = data_range
    C1 = (K1 * R) ** 2
    C2 = (K2 * R) ** 2
    A1, A2, B1, B2 = (
        2 * ux * uy + C1,
        2 * vxy + C2,
        ux**2 + uy**2 + C1,
        vx + vy + C2,
    )
    S = (A1 * A2) / (B1 * B2)
    pad = (win_size - 1) // 2
    crop = tuple(slice(pad, s - pad) for s in S.shape)
    return S[crop].mean(dtype=np.float64)
