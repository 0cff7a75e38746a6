"""CPU oracle for the output pyramid.  TEST INFRASTRUCTURE ONLY.

Restates, on numpy, multiview-stitcher @ 629f72d: ``ngff_utils.mean_dtype`` under
``da.coarsen(..., trim_excess=True)`` (ngff_utils.py:1284-1285, :1315-1322, :1456-1463),
``msi_utils._downsample_sim`` (msi_utils.py:49-77) and
``msi_utils.calc_resolution_levels`` (msi_utils.py:279-327; pinned against the reference's
own function in tests/test_oracle_pyramid.py when /root/reference is present)."""

from __future__ import annotations

import numpy as np


def mean_dtype(arr, **kwargs):
    """ngff_utils.py:1284-1285."""
    return np.mean(arr, **kwargs).astype(arr.dtype)


def coarsen(data, factors):
    """``da.coarsen(mean_dtype, data, axes, trim_excess=True)``: reshape into windows and
    reduce over the window axes."""
    n = [s // f for s, f in zip(data.shape, factors)]
    t = data[tuple(slice(0, a * f) for a, f in zip(n, factors))]
    shp = []
    for a, f in zip(n, factors):
        shp += [a, f]
    return mean_dtype(t.reshape(shp), axis=tuple(range(1, 2 * data.ndim, 2)))


def calc_resolution_levels(spatial_shape, downscale_factors_per_spatial_dim=None, min_shape=100):
    """msi_utils.py:279-327."""
    sdims = list(spatial_shape.keys())
    if downscale_factors_per_spatial_dim is None:
        downscale_factors_per_spatial_dim = {dim: 2 for dim in sdims}
    res_shapes = [spatial_shape]
    res_rel = [{dim: 1 for dim in sdims}]
    res_abs = [{dim: 1 for dim in sdims}]
    while True:
        new_rel = {
            dim: downscale_factors_per_spatial_dim[dim]
            if res_shapes[-1][dim] // downscale_factors_per_spatial_dim[dim] > min_shape
            else 1
            for dim in sdims
        }
        new_abs = {dim: res_abs[-1][dim] * new_rel[dim] for dim in sdims}
        new_shape = {dim: res_shapes[-1][dim] // new_rel[dim] for dim in sdims}
        if not any(new_rel[dim] > 1 for dim in sdims):
            break
        res_shapes.append(new_shape)
        res_rel.append(new_rel)
        res_abs.append(new_abs)
    return res_shapes, res_rel, res_abs


def build_pyramid(view, downscale_factors_per_spatial_dim=None, min_shape=100):
    """Levels of a view dict ``{"data", "origin", "spacing"}`` (msi_utils.py:49-77 per step)."""
    dims = ["z", "y", "x"][-view["data"].ndim:]
    _, rel, _ = calc_resolution_levels(dict(zip(dims, view["data"].shape)), downscale_factors_per_spatial_dim, min_shape)
    levels = [view]
    for step in rel[1:]:
        prev = levels[-1]
        levels.append({
            "data": coarsen(prev["data"], [step[d] for d in dims]),
            "spacing": {d: prev["spacing"][d] * step[d] for d in dims},
            "origin": {d: prev["origin"][d] + (step[d] - 1) * prev["spacing"][d] / 2 for d in dims},
        })
    return levels
