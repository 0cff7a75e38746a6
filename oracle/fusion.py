"""numpy/scipy restatement of the reference's per-chunk fusion path.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Each function cites the
reference lines it follows; paths are relative to
``/root/reference/src/multiview_stitcher``.

Data model used here instead of xarray "sims":

* a *view* (or view slice) is a dict ``{"data": ndarray (z,)y,x,
  "origin": {dim: float}, "spacing": {dim: float}}``;
* a bounding box / stack-properties dict is the reference's own
  ``{"origin": {dim:..}, "spacing": {dim:..}, "shape": {dim:..}}``;
* an affine ``param`` is a plain ``(ndim+1, ndim+1)`` float64 array mapping
  view physical coordinates to fused-space physical coordinates (zyx order).
"""

from __future__ import annotations

import inspect
import warnings

import numpy as np
from scipy import ndimage

SPATIAL_DIMS = ["z", "y", "x"]


def sdims_of(ndim):
    return SPATIAL_DIMS[-ndim:]


# --------------------------------------------------------------------------
# transformation.py
# --------------------------------------------------------------------------


def pixel_affine(p, output_stack_properties, input_origin, input_spacing):
    """Physical affine (output -> input) to scipy's pixel matrix/offset.

    Follows transformation.py:31-83: ``M' = Sy^-1 M Sx`` (:56), origins made
    relative to the output origin (:60-65), both rounded to 10 decimals
    (:72-74) and near-integer offsets snapped (:79-83).
    """
    ndim = len(input_spacing)
    dims = sdims_of(ndim)
    p = np.asarray(p, dtype=np.float64)
    matrix = p[:ndim, :ndim]
    offset = p[:ndim, ndim]

    Sx = np.diag([output_stack_properties["spacing"][d] for d in dims])
    Sy = np.diag([input_spacing[d] for d in dims])
    Ox = np.array([output_stack_properties["origin"][d] for d in dims])
    Oy = np.array([input_origin[d] for d in dims])

    matrix_prime = np.linalg.solve(Sy, np.dot(matrix, Sx))
    local_input_origin = Oy - Ox
    local_offset = offset + np.dot(matrix - np.eye(ndim), Ox)
    offset_prime = np.linalg.solve(Sy, local_offset - local_input_origin)

    matrix_prime = np.around(matrix_prime, decimals=10)
    offset_prime = np.around(offset_prime, decimals=10)

    nearest_integer = np.round(offset_prime)
    near_integer = np.isclose(offset_prime, nearest_integer, rtol=0, atol=1e-6)
    offset_prime[near_integer] = nearest_integer[near_integer]
    return matrix_prime, offset_prime


def transform_view(
    view, p, output_stack_properties, input_spacing=None, order=1, cval=0.0
):
    """``transformation.transform_sim`` (transformation.py:15-148) on a view
    dict; returns the resampled ndarray on the output grid."""
    data = view["data"]
    ndim = data.ndim
    dims = sdims_of(ndim)
    if p is None:
        p = np.eye(ndim + 1)
    if input_spacing is None:
        input_spacing = view["spacing"]
    matrix_prime, offset_prime = pixel_affine(
        p, output_stack_properties, view["origin"], input_spacing
    )
    output_shape = tuple(int(output_stack_properties["shape"][d]) for d in dims)

    # transformation.py:102-119 -- identical sampling grid: hand data through
    is_noop = (
        output_shape == tuple(data.shape)
        and np.allclose(matrix_prime, np.eye(ndim), rtol=0, atol=1e-10)
        and np.allclose(offset_prime, 0, rtol=0, atol=1e-10)
    )
    if is_noop:
        return data
    # transformation.py:136-139
    return ndimage.affine_transform(
        data,
        matrix=matrix_prime,
        offset=offset_prime,
        output_shape=output_shape,
        mode="constant",
        cval=cval,
        order=order,
    )


# --------------------------------------------------------------------------
# weights.py
# --------------------------------------------------------------------------


def normalize_weights(weights):
    """weights.py:325-345."""
    wsum = np.nansum(weights, axis=0)
    wsum[wsum == 0] = 1
    return weights / wsum


def _shrink_source_bb(source_bb, shrink_distance):
    """weights.py:348-388."""
    dims = list(source_bb["origin"].keys())
    if isinstance(shrink_distance, (int, float)):
        shrink_distance = {d: float(shrink_distance) for d in dims}
    return {
        "origin": {
            d: source_bb["origin"][d] + shrink_distance.get(d, 0) for d in dims
        },
        "spacing": source_bb["spacing"],
        "shape": {
            d: source_bb["shape"][d]
            - 2 * shrink_distance.get(d, 0) / source_bb["spacing"][d]
            for d in dims
        },
    }


def blending_support(source_bb, blending_widths=None, shrink_distance=0):
    """The 5^ndim EDT support table and its placement, weights.py:430-470.

    Returns ``(table float32, origin dict, spacing dict)``.
    """
    if blending_widths is None:
        blending_widths = {"z": 3, "y": 10, "x": 10}
    dims = sorted(source_bb["origin"].keys())[::-1]
    if shrink_distance:
        source_bb = _shrink_source_bb(source_bb, shrink_distance)
    ndim = len(dims)
    mask = np.zeros([5] * ndim)
    mask[(slice(1, -1),) * ndim] = 1
    support_spacing = {
        d: (source_bb["shape"][d] - 1) / 4 * source_bb["spacing"][d] for d in dims
    }
    edt_support_spacing = {
        d: support_spacing[d]
        * (source_bb["shape"][d] - 1 + 2 * 1)
        / (source_bb["shape"][d] - 1)
        for d in dims
    }
    edt_support_origin = {
        d: source_bb["origin"][d] - 1 * source_bb["spacing"][d] for d in dims
    }
    edt_support = ndimage.distance_transform_edt(
        mask,
        sampling=[edt_support_spacing[d] / blending_widths[d] for d in dims],
    )
    return edt_support.astype(np.float32), edt_support_origin, edt_support_spacing


def cosine_weights(x):
    """weights.py:502-507 (in place on a float32 array, like the reference)."""
    mask = x < 1
    x[mask] = (np.cos((1 - x[mask]) * np.pi) + 1) / 2
    x = np.clip(x, 0, 1)
    return x


def get_blending_weights(
    target_bb, source_bb, affine, blending_widths=None, shrink_distance=0
):
    """weights.py:391-511."""
    table, origin, spacing = blending_support(
        source_bb, blending_widths, shrink_distance
    )
    support = {"data": table, "origin": origin, "spacing": spacing}
    target_weights = transform_view(
        support,
        p=np.linalg.inv(affine),
        output_stack_properties=target_bb,
        order=1,
        cval=0.0,
    )
    if target_weights is table:  # no-op shortcut returned the table itself
        target_weights = table.copy()
    return cosine_weights(target_weights)


def nan_gaussian_filter(ar, *args, **kwargs):
    """weights.py:293-322."""
    U = ar
    nan_mask = np.isnan(U)
    V = U.copy()
    V[nan_mask] = 0
    VV = ndimage.gaussian_filter(V, *args, **kwargs)
    W = 0 * U.copy() + 1
    W[nan_mask] = 0
    WW = ndimage.gaussian_filter(W, *args, **kwargs)
    WW[nan_mask] = 1
    Z = VV / WW
    Z[nan_mask] = np.nan
    return Z


def content_based(transformed_views, blending_weights, sigma_1=5, sigma_2=11):
    """weights.py:22-74 (Preibisch content-based weights)."""
    transformed_views = transformed_views.astype(np.float32)
    transformed_views[blending_weights < 1e-7] = np.nan
    weights = [
        nan_gaussian_filter(
            (sim_t - nan_gaussian_filter(sim_t, sigma=sigma_1, mode="reflect"))
            ** 2,
            sigma=sigma_2,
            mode="reflect",
        )
        for sim_t in transformed_views
    ]
    weights = np.stack(weights, axis=0)
    return normalize_weights(weights)


content_based.required_overlap = lambda kwargs: 2 * kwargs.get("sigma_2", 11)


# --------------------------------------------------------------------------
# fusion/_core.py
# --------------------------------------------------------------------------


def max_fusion(transformed_views):
    """fusion/_core.py:42-58."""
    return np.nanmax(transformed_views, axis=0)


def weighted_average_fusion(transformed_views, blending_weights, fusion_weights=None):
    """fusion/_core.py:61-94."""
    if fusion_weights is None:
        additive_weights = blending_weights
    else:
        additive_weights = blending_weights * fusion_weights
        additive_weights = normalize_weights(additive_weights)
    product = transformed_views * additive_weights
    return np.nansum(product, axis=0).astype(transformed_views[0].dtype)


def simple_average_fusion(transformed_views):
    """fusion/_core.py:97-131."""
    number_of_valid_views = np.zeros(transformed_views[0].shape, dtype=np.float32)
    for tv in transformed_views:
        number_of_valid_views = np.nansum(
            [number_of_valid_views, ~np.isnan(tv)], axis=0
        )
    number_of_valid_views[number_of_valid_views == 0] = np.nan
    return (np.nansum(transformed_views, axis=0) / number_of_valid_views).astype(
        transformed_views[0].dtype
    )


def _has_keyword(func, name):
    if func is None:
        return False
    try:
        return name in inspect.signature(func).parameters
    except (TypeError, ValueError):
        return False


def fuse_np(
    sims,
    params,
    output_properties,
    fusion_func=weighted_average_fusion,
    fusion_func_kwargs=None,
    weights_func=None,
    weights_func_kwargs=None,
    trim_overlap_in_pixels=0,
    interpolation_order=1,
    full_view_bbs=None,
    spacings=None,
    blending_widths=None,
    shrink_distance=0,
    return_intermediates=False,
):
    """fusion/_core.py:1513-1733 on view dicts.

    ``sims[i]["data"]`` may be any slice of view ``i`` (its ``origin`` is the
    slice origin); ``full_view_bbs[i]`` is the bounding box of the whole view.
    """
    requires_bw = _has_keyword(fusion_func, "blending_weights") or _has_keyword(
        weights_func, "blending_weights"
    )
    fusion_func_kwargs = dict(fusion_func_kwargs or {})
    weights_func_kwargs = dict(weights_func_kwargs or {})
    input_dtype = sims[0]["data"].dtype
    if spacings is None:
        spacings = (
            [bb["spacing"] for bb in full_view_bbs]
            if full_view_bbs is not None
            else [None] * len(sims)
        )

    # :1621-1632
    field_ims_t = np.stack(
        [
            transform_view(
                {**sim, "data": sim["data"].astype(np.float32)},
                np.linalg.inv(param),
                output_stack_properties=output_properties,
                input_spacing=spacing,
                order=interpolation_order,
                cval=np.nan,
            )
            for sim, param, spacing in zip(sims, params, spacings)
        ]
    )

    # :1635-1651
    if requires_bw:
        field_ws_t = np.stack(
            [
                get_blending_weights(
                    target_bb=output_properties,
                    source_bb=full_view_bbs[iview],
                    affine=params[iview],
                    blending_widths=blending_widths,
                    shrink_distance=shrink_distance,
                )
                for iview in range(len(sims))
            ]
        )
        field_ws_t = field_ws_t * ~np.isnan(field_ims_t)
        field_ws_t = normalize_weights(field_ws_t)
    else:
        field_ws_t = None

    # :1653-1662
    fusion_func_kwargs["transformed_views"] = field_ims_t
    if _has_keyword(fusion_func, "params"):
        fusion_func_kwargs["params"] = params
    if requires_bw:
        fusion_func_kwargs["blending_weights"] = field_ws_t
    if (
        _has_keyword(fusion_func, "output_spacing")
        and "output_spacing" not in fusion_func_kwargs
    ):
        fusion_func_kwargs["output_spacing"] = output_properties["spacing"]

    # :1664-1680
    fusion_weights = None
    if weights_func is not None and _has_keyword(fusion_func, "fusion_weights"):
        weights_func_kwargs["transformed_views"] = field_ims_t
        if _has_keyword(weights_func, "params"):
            weights_func_kwargs["params"] = params
        if _has_keyword(weights_func, "blending_weights"):
            weights_func_kwargs["blending_weights"] = field_ws_t
        if (
            _has_keyword(weights_func, "output_chunksize")
            and "output_chunksize" not in weights_func_kwargs
        ):
            weights_func_kwargs["output_chunksize"] = output_properties["shape"]
        fusion_weights = weights_func(**weights_func_kwargs)
        fusion_func_kwargs["fusion_weights"] = fusion_weights

    # :1682-1685 (func_ignore_nan_warning, :1504-1510)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", category=RuntimeWarning)
        fused = fusion_func(**fusion_func_kwargs)

    # :1687-1711
    dims = list(output_properties["shape"].keys())
    if not isinstance(trim_overlap_in_pixels, dict):
        trim_overlap_in_pixels = {d: trim_overlap_in_pixels for d in dims}
    if any(trim_overlap_in_pixels[d] > 0 for d in dims):
        fused = fused[
            tuple(
                slice(trim_overlap_in_pixels[d], -trim_overlap_in_pixels[d])
                if trim_overlap_in_pixels[d] > 0
                else slice(None)
                for d in dims
            )
        ]

    # :1713
    with np.errstate(invalid="ignore"):
        fused = np.nan_to_num(fused).astype(input_dtype)
    if return_intermediates:
        return fused, field_ims_t, field_ws_t, fusion_weights
    return fused


# --------------------------------------------------------------------------
# output stack geometry (fusion/_core.py:1821-1992) and chunk grid
# (mv_graph.py:934-986) -- needed to replay the reference's fuse() KATs
# --------------------------------------------------------------------------


def view_bb(view):
    dims = sdims_of(view["data"].ndim)
    return {
        "origin": {d: float(view["origin"][d]) for d in dims},
        "spacing": {d: float(view["spacing"][d]) for d in dims},
        "shape": {d: int(s) for d, s in zip(dims, view["data"].shape)},
    }


def calc_stack_properties(view_bbs, params, spacing, mode="union"):
    """fusion/_core.py:1821-1992 (pixel-centre vertices :1957-1962, shape rule
    :1982-1985)."""
    ndim = len(spacing)
    dims = sdims_of(ndim)
    sp = np.array([spacing[d] for d in dims], dtype=float)
    corners = np.array(list(np.ndindex(*([2] * ndim))), dtype=float)
    verts = []
    for bb, p in zip(view_bbs, params):
        shape = np.array([bb["shape"][d] for d in dims], dtype=float)
        vsp = np.array([bb["spacing"][d] for d in dims], dtype=float)
        org = np.array([bb["origin"][d] for d in dims], dtype=float)
        v = corners * (shape - 1) * vsp + org
        p = np.asarray(p, dtype=float)
        verts.append(np.dot(p[:ndim, :ndim], v.T).T + p[:ndim, ndim])
    verts = np.array(verts)
    if mode == "union":
        lo, hi = np.min(np.min(verts, 1), 0), np.max(np.max(verts, 1), 0)
    elif mode == "intersection":
        lo, hi = np.max(np.min(verts, 1), 0), np.min(np.max(verts, 1), 0)
    else:
        raise NotImplementedError(mode)
    shape = np.floor((hi - lo) / sp + 1e-9).astype(np.uint64) + 1
    return {
        "origin": {d: float(lo[i]) for i, d in enumerate(dims)},
        "spacing": {d: float(sp[i]) for i, d in enumerate(dims)},
        "shape": {d: int(shape[i]) for i, d in enumerate(dims)},
    }


def chunk_bbs(array_bb, chunksize):
    """mv_graph.py:934-986 for regular chunk sizes."""
    dims = sorted(array_bb["origin"].keys())[::-1]
    bounds = []
    for d in dims:
        n, c = int(array_bb["shape"][d]), int(chunksize[d])
        starts = list(range(0, n, c))
        bounds.append([(s, min(c, n - s)) for s in starts])
    out = []
    for idx in np.ndindex(*[len(b) for b in bounds]):
        out.append(
            (
                {
                    "origin": {
                        d: array_bb["origin"][d]
                        + array_bb["spacing"][d] * bounds[i][idx[i]][0]
                        for i, d in enumerate(dims)
                    },
                    "shape": {d: bounds[i][idx[i]][1] for i, d in enumerate(dims)},
                    "spacing": array_bb["spacing"],
                },
                tuple(bounds[i][idx[i]][0] for i in range(len(dims))),
            )
        )
    return out


def fuse(
    views,
    params,
    output_stack_properties=None,
    output_spacing=None,
    output_chunksize=None,
    overlap_in_pixels=None,
    **fuse_np_kwargs,
):
    """Chunked fusion of whole views: the arithmetic of ``fusion.fuse``'s lazy
    path (fusion/_core.py:1173-1462) with the planner's per-chunk view windows
    replaced by the whole views -- a superset of every window the planner
    would hand to ``fuse_np``, so results are identical wherever the planner's
    window covers the chunk (which is its contract, fusion/_core.py:462-533).

    Views that cannot touch a chunk are skipped by an AABB test like
    fusion/_core.py:582-653.
    """
    ndim = views[0]["data"].ndim
    dims = sdims_of(ndim)
    bbs = [view_bb(v) for v in views]
    if output_spacing is None:
        output_spacing = bbs[0]["spacing"]  # fusion/_core.py:316-325
    if output_stack_properties is None:
        output_stack_properties = calc_stack_properties(bbs, params, output_spacing)
    osp = output_stack_properties
    full_shape = tuple(int(osp["shape"][d]) for d in dims)
    if output_chunksize is None:
        output_chunksize = {d: s for d, s in zip(dims, full_shape)}
    weights_func = fuse_np_kwargs.get("weights_func")
    fusion_func = fuse_np_kwargs.get("fusion_func", weighted_average_fusion)
    if overlap_in_pixels is None:
        overlap_in_pixels = 0
        for f, kw in (
            (weights_func, fuse_np_kwargs.get("weights_func_kwargs")),
            (fusion_func, fuse_np_kwargs.get("fusion_func_kwargs")),
        ):
            if f is not None and hasattr(f, "required_overlap"):
                defaults = {
                    k: v.default
                    for k, v in inspect.signature(f).parameters.items()
                    if v.default is not inspect.Parameter.empty
                }
                overlap_in_pixels = max(
                    overlap_in_pixels,
                    int(np.ceil(f.required_overlap({**defaults, **(kw or {})}))),
                )
    ov = int(overlap_in_pixels)

    out = np.zeros(full_shape, dtype=views[0]["data"].dtype)
    for cbb, start in chunk_bbs(osp, output_chunksize):
        # halo'd chunk (fusion/_core.py:1225-1254)
        hbb = {
            "origin": {
                d: cbb["origin"][d] - ov * osp["spacing"][d] for d in dims
            },
            "spacing": cbb["spacing"],
            "shape": {d: cbb["shape"][d] + 2 * ov for d in dims},
        }
        sel = [
            i
            for i in range(len(views))
            if _view_touches(bbs[i], params[i], hbb, ndim)
        ]
        sl = tuple(
            slice(start[i], start[i] + cbb["shape"][d]) for i, d in enumerate(dims)
        )
        if not sel:
            continue
        out[sl] = fuse_np(
            [views[i] for i in sel],
            [params[i] for i in sel],
            hbb,
            full_view_bbs=[bbs[i] for i in sel],
            trim_overlap_in_pixels=ov,
            **fuse_np_kwargs,
        )
    return out, osp


def _view_touches(bb, param, target_bb, ndim):
    dims = sdims_of(ndim)
    corners = np.array(list(np.ndindex(*([2] * ndim))), dtype=float)
    shape = np.array([bb["shape"][d] for d in dims], dtype=float)
    v = corners * (shape - 1) * np.array([bb["spacing"][d] for d in dims]) + np.array(
        [bb["origin"][d] for d in dims]
    )
    p = np.asarray(param, dtype=float)
    vt = np.dot(p[:ndim, :ndim], v.T).T + p[:ndim, ndim]
    lo, hi = vt.min(0), vt.max(0)
    tlo = np.array([target_bb["origin"][d] for d in dims])
    thi = tlo + (np.array([target_bb["shape"][d] for d in dims]) - 1) * np.array(
        [target_bb["spacing"][d] for d in dims]
    )
    eps = 1e-6
    return bool(np.all(hi >= tlo - eps) and np.all(lo <= thi + eps))
