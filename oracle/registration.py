"""numpy/scipy restatement of the reference's pairwise phase-correlation
registration (registration.py:353-565) and its caller-side guard
(registration.py:1477-1544).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  scikit-image calls go to
``oracle.skimage_restated``.
"""

from __future__ import annotations

import warnings

import numpy as np
from scipy import ndimage, stats

from . import skimage_restated as sk


def affine_from_translation(translation):
    """param_utils.py:7-14."""
    ndim = len(translation)
    M = np.concatenate([translation, [1]], axis=0)
    M = np.concatenate([np.eye(ndim + 1)[:, :ndim], M[:, None]], axis=1)
    return M


def link_quality_metric_func(im0, im1t):
    """registration.py:109-111."""
    return stats.spearmanr(im0.flatten(), im1t.flatten()).correlation


def get_bb_from_nanmask(mask):
    """registration.py:482-489."""
    bbs = []
    for idim in range(mask.ndim):
        axes = list(range(mask.ndim))
        axes.remove(idim)
        valids = np.where(np.max(mask, axis=tuple(axes)))
        bbs.append([np.min(valids), np.max(valids)])
    return bbs


def shift_candidates(im0nn, im1nn, im0, im1, im0nm, im1nm, upsample_factor):
    """registration.py:410-443: the two (three with NaNs) sub-pixel shifts."""
    cands = []
    for normalization in ["phase", None]:
        cands.append(
            sk.phase_cross_correlation(
                im0nn,
                im1nn,
                disambiguate=False,
                normalization=normalization,
                upsample_factor=upsample_factor,
            )[0]
        )
    if np.any([im0nm, im1nm]):
        cands.append(
            sk.phase_cross_correlation(
                im0,
                im1,
                reference_mask=im0nm,
                moving_mask=im1nm,
                disambiguate=False,
                upsample_factor=upsample_factor,
            )[0]
        )
    return cands


def expand_candidates(shift_cands, shape, max_shift_per_dim):
    """registration.py:461-477: sign / wrap-around alternatives per axis."""
    ndim = len(shape)
    t_candidates = []
    for sc in shift_cands:
        for s in np.ndindex(tuple(1 if sc[d] == 0 else 4 for d in range(ndim))):
            t = []
            for d in range(ndim):
                if s[d] == 0:
                    t.append(sc[d])
                elif s[d] == 1:
                    t.append(-sc[d])
                elif s[d] == 2:
                    t.append(-(sc[d] - shape[d]))
                elif s[d] == 3:
                    t.append(-sc[d] - shape[d])
            if np.max(np.abs(t)) < max_shift_per_dim:
                t_candidates.append(t)
    return t_candidates


def phase_correlation_registration(
    fixed_data,
    moving_data,
    disambiguate_region_mode=None,
    upsample_factor=None,
    return_details=False,
):
    """registration.py:353-565 on plain float arrays (NaN = outside)."""
    im0 = np.asarray(fixed_data)
    im1 = np.asarray(moving_data)
    ndim = im0.ndim

    with warnings.catch_warnings():
        warnings.simplefilter("ignore", category=RuntimeWarning)
        # :381-389
        im0, im1 = (
            sk.rescale_intensity(
                im, in_range=(np.nanmin(im), np.nanmax(im)), out_range=(0, 1)
            )
            for im in [im0, im1]
        )
    im0nm = np.isnan(im0)
    im1nm = np.isnan(im1)
    if disambiguate_region_mode is None:
        disambiguate_region_mode = (
            "intersection" if np.any([im0nm, im1nm]) else "union"
        )
    valid_pixels1 = np.sum(~im1nm)
    if np.any([im0nm, im1nm]):
        im0nn = np.nan_to_num(im0)
        im1nn = np.nan_to_num(im1)
    else:
        im0nn, im1nn = im0, im1
    if upsample_factor is None:
        upsample_factor = 10 if ndim == 2 else 2

    with warnings.catch_warnings():
        warnings.simplefilter("ignore", category=RuntimeWarning)
        shift_cands = shift_candidates(
            im0nn, im1nn, im0, im1, im0nm, im1nm, upsample_factor
        )

    max_shift_per_dim = np.max([im.shape for im in [im0, im1]])
    data_range = np.nanmax([im0, im1]) - np.nanmin([im0, im1])
    im1_min = np.nanmin(im1)

    t_candidates = expand_candidates(shift_cands, im1.shape, max_shift_per_dim)
    if not len(t_candidates):
        return [np.zeros(ndim)]

    im0_bb = get_bb_from_nanmask(~im0nm)
    disambiguate_metric_vals = []
    quality_metric_vals = []
    for t_ in t_candidates:
        im1t = ndimage.affine_transform(
            im1,
            affine_from_translation(list(t_)),
            order=1,
            mode="constant",
            cval=np.nan,
        )
        mask = ~np.isnan(im1t) * ~im0nm
        if np.all(~mask) or float(np.sum(mask)) / valid_pixels1 < 0.1:
            disambiguate_metric_val = -1
            quality_metric_val = -1
        else:
            im1t_bb = get_bb_from_nanmask(~np.isnan(im1t))
            if disambiguate_region_mode == "union":
                mask_slices = tuple(
                    slice(
                        min(im0_bb[i][0], im1t_bb[i][0]),
                        max(im0_bb[i][1], im1t_bb[i][1]) + 1,
                    )
                    for i in range(ndim)
                )
            elif disambiguate_region_mode == "intersection":
                mask_slices = tuple(
                    slice(
                        max(im0_bb[i][0], im1t_bb[i][0]),
                        min(im0_bb[i][1], im1t_bb[i][1]) + 1,
                    )
                    for i in range(ndim)
                )
            with warnings.catch_warnings():
                warnings.simplefilter("ignore", category=RuntimeWarning)
                flat = np.nanmax(im1t[mask_slices]) <= im1_min
            if flat:
                # :530-533 -- NOTE: skips the appends below (list semantics)
                continue
            min_shape = np.min(im0[mask_slices].shape)
            ssim_win_size = np.min([7, min_shape - ((min_shape - 1) % 2)])
            if ssim_win_size < 3 or np.max(im1t[mask_slices]) <= im1_min:
                disambiguate_metric_val = -1
            else:
                disambiguate_metric_val = sk.structural_similarity(
                    np.nan_to_num(im0[mask_slices]),
                    np.nan_to_num(im1t[mask_slices]),
                    data_range=data_range,
                    win_size=int(ssim_win_size),
                )
            quality_metric_val = link_quality_metric_func(im0[mask], im1t[mask] - 1)
        disambiguate_metric_vals.append(disambiguate_metric_val)
        quality_metric_vals.append(quality_metric_val)

    argmax_index = np.nanargmax(disambiguate_metric_vals)
    t = t_candidates[argmax_index]
    result = {
        "affine_matrix": affine_from_translation(t),
        "quality": quality_metric_vals[argmax_index],
    }
    if return_details:
        result["shift_candidates"] = [np.asarray(s) for s in shift_cands]
        result["t_candidates"] = t_candidates
        result["ssim"] = disambiguate_metric_vals
        result["spearman"] = quality_metric_vals
    return result


def dispatch_pairwise_reg_func(pairwise_reg_func, fixed_data, moving_data, **kwargs):
    """registration.py:1477-1544: constant-image guard, then the hook."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", category=RuntimeWarning)
        constant = any(
            np.nanmin(d) == np.nanmax(d) for d in (fixed_data, moving_data)
        )
    if constant:
        warnings.warn(
            "An overlap region between tiles/views is all zero or constant.",
            UserWarning,
            stacklevel=1,
        )
        ndim = np.asarray(fixed_data).ndim
        return {"affine_matrix": np.eye(ndim + 1), "quality": np.nan}
    return pairwise_reg_func(fixed_data, moving_data, **kwargs)
