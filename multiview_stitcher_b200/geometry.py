"""Host-side geometry of the fusion path (numpy, O(tiles + chunks) work).

Produces exactly the numbers the reference hands to its resampler so the CUDA
kernel samples the same positions:

* ``pixel_affine``      <- transformation.transform_sim, transformation.py:31-83
* ``blending_table``    <- weights.get_blending_weights, weights.py:430-470
* ``union_stack_props`` <- fusion/_core.py:1821-1992
* ``chunk_grid``        <- mv_graph.get_chunk_bbs, mv_graph.py:934-986

Conventions: spatial dims are ordered (z,) y, x; bounding boxes are the
reference's ``{"origin": {dim: ..}, "spacing": {dim: ..}, "shape": {dim: ..}}``
dicts; affines are (ndim+1)^2 float64 arrays in physical units.
"""

from __future__ import annotations

import numpy as np

SPATIAL_DIMS = ["z", "y", "x"]
DEFAULT_BLENDING_WIDTHS = {"z": 3, "y": 10, "x": 10}  # weights.py:430-431
DEFAULT_CHUNKSIZE_2D = {"y": 2048, "x": 2048}  # spatial_image_utils.py:22
DEFAULT_CHUNKSIZE_3D = {"z": 256, "y": 256, "x": 256}  # spatial_image_utils.py:21


def spatial_dims(ndim):
    return SPATIAL_DIMS[-ndim:]


def bb_arrays(bb, dims):
    return (
        np.array([bb["origin"][d] for d in dims], dtype=np.float64),
        np.array([bb["spacing"][d] for d in dims], dtype=np.float64),
        np.array([bb["shape"][d] for d in dims], dtype=np.float64),
    )


def pixel_affine(p, out_origin, out_spacing, in_origin, in_spacing):
    """Physical affine ``p`` (output space -> input space) as the pixel
    matrix / offset of ``scipy.ndimage.affine_transform``.

    ``x_in_px = matrix @ x_out_px + offset``.  Both are rounded to 10 decimals
    and offsets within 1e-6 of an integer are snapped, as the reference does
    before calling scipy (transformation.py:72-83), so the device kernel sees
    the same float64 numbers.
    """
    p = np.asarray(p, dtype=np.float64)
    ndim = len(out_spacing)
    lin = p[:ndim, :ndim]
    trans = p[:ndim, ndim]
    sp_out = np.asarray(out_spacing, dtype=np.float64)
    sp_in = np.asarray(in_spacing, dtype=np.float64)
    o_out = np.asarray(out_origin, dtype=np.float64)
    o_in = np.asarray(in_origin, dtype=np.float64)

    # Sy^-1 (M Sx) with diagonal spacing matrices: the reference's
    # np.linalg.solve(Sy, np.dot(M, Sx)) reduces to one multiply and one divide per
    # entry (the LU of a diagonal matrix is the matrix itself), bit for bit
    matrix = (lin * sp_out[None, :]) / sp_in[:, None]
    # both origins relative to the output origin (transformation.py:60-65)
    rel = trans + (lin - np.eye(ndim)) @ o_out
    offset = (rel - (o_in - o_out)) / sp_in

    matrix = np.around(matrix, decimals=10)
    offset = np.around(offset, decimals=10)
    snapped = np.round(offset)
    close = np.abs(offset - snapped) <= 1e-6  # np.isclose(rtol=0, atol=1e-6)
    offset[close] = snapped[close]
    return matrix, offset


def embed3(matrix, offset):
    """(ndim x ndim, ndim) -> row-major 3x3 / 3 with identity on a leading z."""
    ndim = len(offset)
    m3 = np.eye(3)
    o3 = np.zeros(3)
    m3[3 - ndim :, 3 - ndim :] = matrix
    o3[3 - ndim :] = offset
    return m3.reshape(9), o3


def blending_table(source_bb, blending_widths=None, shrink_distance=0):
    """The view's 5^ndim blending-support table and its placement.

    The reference builds a 5^ndim mask whose inner 3^ndim nodes are 1 and runs
    an anisotropic Euclidean distance transform on it (weights.py:439-464).
    For that mask the nearest zero node of an inner node lies along one axis,
    so the transform has the closed form
    ``E[k] = min_d( min(k_d, 4 - k_d) * s_d )`` with the sampling
    ``s_d = (shape_d + 1) / 4 * spacing_d / blending_width_d``.

    Returns ``(table float32 (5,)*ndim, origin (ndim,), spacing (ndim,))`` with
    the table placed at ``origin - spacing`` with node pitch
    ``(shape + 1) / 4 * spacing`` (weights.py:448-470).
    """
    if blending_widths is None:
        blending_widths = DEFAULT_BLENDING_WIDTHS
    dims = sorted(source_bb["origin"].keys())[::-1]
    ndim = len(dims)
    origin, spacing, shape = bb_arrays(source_bb, dims)
    if shrink_distance:
        # weights.py:348-388
        if isinstance(shrink_distance, (int, float)):
            shrink = np.full(ndim, float(shrink_distance))
        else:
            shrink = np.array([shrink_distance.get(d, 0) for d in dims], dtype=np.float64)
        origin = origin + shrink
        shape = shape - 2 * shrink / spacing
    support_spacing = (shape - 1) / 4 * spacing
    node_pitch = support_spacing * (shape - 1 + 2 * 1) / (shape - 1)
    table_origin = origin - 1 * spacing
    sampling = node_pitch / np.array([blending_widths[d] for d in dims], dtype=np.float64)
    k = np.arange(5)
    ramp = np.minimum(k, 4 - k).astype(np.float64)
    table = None
    for d in range(ndim):
        shp = [1] * ndim
        shp[d] = 5
        # sqrt of the squared distance, like the EDT computes it
        axis_dist = np.sqrt((ramp * sampling[d]) ** 2).reshape(shp)
        table = axis_dist if table is None else np.minimum(table, axis_dist)
    return table.astype(np.float32), table_origin, node_pitch


def union_stack_props(view_bbs, params, spacing, mode="union"):
    """Output stack enclosing all transformed views (pixel-centre corners,
    fusion/_core.py:1957-1962; ``shape = floor(extent/spacing + 1e-9) + 1``,
    :1982-1985)."""
    ndim = len(spacing)
    dims = spatial_dims(ndim)
    sp = np.array([spacing[d] for d in dims], dtype=np.float64)
    corners = np.array(list(np.ndindex(*([2] * ndim))), dtype=np.float64)
    lows, highs = [], []
    for bb, p in zip(view_bbs, params):
        o, s, n = bb_arrays(bb, dims)
        p = np.asarray(p, dtype=np.float64)
        v = corners * (n - 1) * s + o
        vt = np.dot(p[:ndim, :ndim], v.T).T + p[:ndim, ndim]
        lows.append(vt.min(0))
        highs.append(vt.max(0))
    if mode == "union":
        lo, hi = np.min(lows, 0), np.max(highs, 0)
    elif mode == "intersection":
        lo, hi = np.max(lows, 0), np.min(highs, 0)
    else:
        raise NotImplementedError(f"output_stack_mode {mode!r}")
    shape = np.floor((hi - lo) / sp + 1e-9).astype(np.uint64) + 1
    return {
        "origin": {d: float(lo[i]) for i, d in enumerate(dims)},
        "spacing": {d: float(sp[i]) for i, d in enumerate(dims)},
        "shape": {d: int(shape[i]) for i, d in enumerate(dims)},
    }


def chunk_grid(array_bb, chunksize):
    """Regular chunk grid over ``array_bb``: list of (start index tuple,
    shape tuple) in (z,) y, x order."""
    dims = sorted(array_bb["origin"].keys())[::-1]
    per_dim = []
    for d in dims:
        n, c = int(array_bb["shape"][d]), int(chunksize[d])
        per_dim.append([(s, min(c, n - s)) for s in range(0, n, c)])
    out = []
    for idx in np.ndindex(*[len(b) for b in per_dim]):
        out.append(
            (
                tuple(per_dim[i][idx[i]][0] for i in range(len(dims))),
                tuple(per_dim[i][idx[i]][1] for i in range(len(dims))),
            )
        )
    return out


def transformed_aabb(bb, param, dims):
    """Axis-aligned physical bounds of a view's pixel-centre corners after
    ``param``."""
    ndim = len(dims)
    o, s, n = bb_arrays(bb, dims)
    corners = np.array(list(np.ndindex(*([2] * ndim))), dtype=np.float64)
    v = corners * (n - 1) * s + o
    p = np.asarray(param, dtype=np.float64)
    vt = np.dot(p[:ndim, :ndim], v.T).T + p[:ndim, ndim]
    return vt.min(0), vt.max(0)
