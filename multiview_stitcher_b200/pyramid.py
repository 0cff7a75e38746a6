"""Multiscale pyramid of a fused stack on the device (the output
side): the level-to-level mean-binning the reference runs through dask / xarray when it
writes an OME-Zarr (``ngff_utils.write_and_return_downsampled_sim`` with
``mean_dtype = np.mean(...).astype(dtype)``, ngff_utils.py:1284-1330, :1456-1463) or builds
a MultiscaleSpatialImage (``msi_utils._downsample_sim``, msi_utils.py:49-77).

Every level is one launch of ``mvs_bin_mean`` over the previous level, which stays in HBM;
a level is read once and 1/prod(factors) of it written: HBM-bound, no reuse to exploit.
Encoding / writing the chunks (zarr, compression) stays with the caller.
"""

from __future__ import annotations

from ._lib import EngineError


def calc_resolution_levels(spatial_shape, downscale_factors_per_spatial_dim=None, min_shape=100):
    """``msi_utils.calc_resolution_levels`` (msi_utils.py:279-327): a dimension keeps being
    halved (or divided by its factor) while the result stays above ``min_shape``.  Returns
    ``(shapes, relative factors, absolute factors)``, level 0 included."""
    dims = list(spatial_shape.keys())
    f = downscale_factors_per_spatial_dim or {d: 2 for d in dims}
    shapes = [{d: int(spatial_shape[d]) for d in dims}]
    rel = [{d: 1 for d in dims}]
    absf = [{d: 1 for d in dims}]
    while True:
        step = {d: (int(f[d]) if shapes[-1][d] // int(f[d]) > min_shape else 1) for d in dims}
        if not any(v > 1 for v in step.values()):
            return shapes, rel, absf
        shapes.append({d: shapes[-1][d] // step[d] for d in dims})
        absf.append({d: absf[-1][d] * step[d] for d in dims})
        rel.append(step)


def downsample(view, factors):
    """One pyramid step of a resident view: ``coarsen(factors, boundary="trim")`` with
    ``mean(...).astype(dtype)``; spacing ``* f``, origin ``+ (f - 1) * spacing / 2``
    (msi_utils.py:63-72).  Returns a ``DeviceView``."""
    from .fusion import DeviceView, to_device_view
    from .pairs import bin_view

    dv = to_device_view(view)
    fac = {d: int(factors.get(d, 1)) for d in dv.dims}
    if any(v < 1 for v in fac.values()):
        raise EngineError(f"downscale factors must be >= 1, got {fac}")
    t = bin_view(dv, fac, skip_nan=False)
    spacing = {d: dv.spacing[d] * fac[d] for d in dv.dims}
    origin = {d: dv.origin[d] + (fac[d] - 1) * dv.spacing[d] / 2 for d in dv.dims}
    return DeviceView(t, origin, spacing)


def build_pyramid(view, downscale_factors_per_spatial_dim=None, min_shape=100):
    """All resolution levels of a (fused) view, level 0 first, as ``DeviceView``s."""
    from .fusion import to_device_view

    dv = to_device_view(view)
    _, rel, _ = calc_resolution_levels(dict(zip(dv.dims, dv.shape)), downscale_factors_per_spatial_dim, min_shape)
    levels = [dv]
    for step in rel[1:]:
        levels.append(downsample(levels[-1], step))
    return levels
