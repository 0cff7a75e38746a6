"""CPU oracle for the Zarr v2 / OME-Zarr 0.4 output side.
TEST INFRASTRUCTURE ONLY.

Restates, on numpy, multiview-stitcher @ 629f72d: the chunk encoding and array metadata the
reference itself spells out in ``ngff_utils.VirtualOMEZarr`` (``array_zarray`` :306-325,
``read_chunk`` :372-395, ``_pad_edge_chunk`` :425-436, ``_build_root_zattrs`` :224-294,
``_zarr_dtype`` :121-128, ``_fill_value_for_dtype`` :131-139), the per-level NGFF transforms
(``calc_ngff_coordinate_transformations_and_axes`` :1493-1561), the multiscales document
``write_multiscales_metadata`` leaves for NGFF 0.4 (:1185-1224), the ``omero`` block
(:1714-1747) and the level loop of ``write_sim_to_ome_zarr`` (:1564-1712; every chunk
written, ``fill_value=0``).  zarr-python itself (third party, absent here) is what moves the
bytes in the reference; its on-disk v2 layout is the one VirtualOMEZarr serves.

Pinned bit for bit against fixtures produced by the reference's own VirtualOMEZarr /
calc_ngff_coordinate_transformations_and_axes (tests/golden/make_golden_ngff.py ->
ngff_golden.npz; tests/test_oracle_ngff.py).  A store is a plain dict key -> bytes / JSON."""

from __future__ import annotations

import numpy as np

from . import pyramid

SPATIAL = ("z", "y", "x")


def zarr_dtype(dtype):
    dtype = np.dtype(dtype)
    if dtype.byteorder == "=":
        if dtype.itemsize == 1:
            dtype = dtype.newbyteorder("|")
        else:
            dtype = dtype.newbyteorder("<" if np.little_endian else ">")
    return dtype.str


def fill_value_for_dtype(dtype):
    dtype = np.dtype(dtype)
    if np.issubdtype(dtype, np.floating):
        return 0.0
    if np.issubdtype(dtype, np.integer):
        return 0
    if np.issubdtype(dtype, np.bool_):
        return False
    return 0


def array_zarray(shape, chunks, dtype, compressor=None):
    return {
        "zarr_format": 2,
        "shape": [int(s) for s in shape],
        "chunks": [int(c) for c in chunks],
        "dtype": zarr_dtype(dtype),
        "compressor": compressor,
        "fill_value": fill_value_for_dtype(dtype),
        "order": "C",
        "filters": None,
        "dimension_separator": "/",
    }


def encode_chunk(data, chunks, index):
    """Bytes of chunk ``index``: the C-order box, padded to ``chunks`` with the fill value."""
    sl = tuple(slice(i * c, min((i + 1) * c, n)) for i, c, n in zip(index, chunks, data.shape))
    chunk = data[sl]
    if tuple(chunk.shape) != tuple(chunks):
        padded = np.full(chunks, fill_value_for_dtype(data.dtype), dtype=data.dtype)
        padded[tuple(slice(0, s) for s in chunk.shape)] = chunk
        chunk = padded
    return np.ascontiguousarray(chunk).tobytes(order="C")


def encode_array(data, chunks, prefix=""):
    """All chunk files of an array: ``{prefix + "i/j/k": bytes}``."""
    grid = [-(-n // c) for n, c in zip(data.shape, chunks)]
    return {prefix + "/".join(str(i) for i in idx): encode_chunk(data, chunks, idx) for idx in np.ndindex(*grid)}


def decode_array(store, shape, chunks, dtype, prefix=""):
    """Inverse of ``encode_array`` (a missing chunk reads as the fill value)."""
    out = np.full(shape, fill_value_for_dtype(dtype), dtype=dtype)
    grid = [-(-n // c) for n, c in zip(shape, chunks)]
    for idx in np.ndindex(*grid):
        raw = store.get(prefix + "/".join(str(i) for i in idx))
        if raw is None:
            continue
        chunk = np.frombuffer(raw, dtype=dtype).reshape(chunks)
        sl = tuple(slice(i * c, min((i + 1) * c, n)) for i, c, n in zip(idx, chunks, shape))
        out[sl] = chunk[tuple(slice(0, s.stop - s.start) for s in sl)]
    return out


def calc_ngff_coordinate_transformations_and_axes(stack_properties_res0, res_abs_factors, nsdims=None, time_transform=None):
    spacing = stack_properties_res0["spacing"]
    origin = stack_properties_res0["origin"]
    sdims = list(spacing.keys())
    nsdims = list(nsdims or [])
    tt = {"scale": 1.0, "translation": 0.0, "unit": None}
    tt.update(time_transform or {})
    nsdim_scales = [float(tt["scale"]) if dim == "t" else 1.0 for dim in nsdims]
    nsdim_translations = [float(tt["translation"]) if dim == "t" else 0 for dim in nsdims]
    coordtfs = []
    for level in range(len(res_abs_factors)):
        f = res_abs_factors[level]
        coordtfs.append([
            {"type": "scale", "scale": nsdim_scales + [float(s * f[dim]) for dim, s in spacing.items()]},
            {"type": "translation",
             "translation": nsdim_translations + [origin[dim] + (f[dim] - 1) * spacing[dim] / 2 for dim in sdims]},
        ])
    axes = []
    for dim in nsdims + sdims:
        ax = {"name": dim, "type": "channel" if dim == "c" else ("time" if dim == "t" else "space")}
        if dim in sdims:
            ax.update({"unit": "micrometer"})
        if dim == "t" and tt["unit"]:
            ax.update({"unit": tt["unit"]})
        axes.append(ax)
    return coordtfs, axes


def virtual_root_zattrs(levels, dims, name="image"):
    """``VirtualOMEZarr._build_root_zattrs``: levels = [(origin, spacing)] per scale."""
    dim_type = {"t": "time", "c": "channel"}
    axes = []
    for dim in dims:
        ax = {"name": dim, "type": dim_type.get(dim, "space")}
        if dim not in dim_type:
            ax["unit"] = "micrometer"
        axes.append(ax)
    datasets = []
    for i, (origin, spacing) in enumerate(levels):
        datasets.append({
            "path": str(i),
            "coordinateTransformations": [
                {"type": "scale", "scale": [float(spacing[d]) if d in SPATIAL else 1.0 for d in dims]},
                {"type": "translation", "translation": [float(origin[d]) if d in SPATIAL else 0.0 for d in dims]},
            ],
        })
    return {"multiscales": [{"version": "0.4", "name": name, "axes": axes, "datasets": datasets}]}


def write_sim_to_ome_zarr(data, dims, origin, spacing, chunks, downscale_factors_per_spatial_dim=None, c_coords=None,
                          min_shape=100):
    """The store ``write_sim_to_ome_zarr`` produces for an in-memory image, as a dict:
    ``"<level>/.zarray"`` (dict), ``"<level>/<chunk key>"`` (bytes), ``".zgroup"``, ``".zattrs"``."""
    dims = list(dims)
    sdims = [d for d in dims if d in SPATIAL]
    nsdims = [d for d in dims if d not in SPATIAL]
    spatial_shape = {d: int(data.shape[dims.index(d)]) for d in sdims}
    chunk_shape = tuple(1 if d in nsdims else min(int(chunks[d]), spatial_shape[d]) for d in dims)
    res_shapes, res_rel, res_abs = pyramid.calc_resolution_levels(spatial_shape, downscale_factors_per_spatial_dim, min_shape)
    coordtfs, axes = calc_ngff_coordinate_transformations_and_axes(
        {"spacing": {d: spacing[d] for d in sdims}, "origin": {d: origin[d] for d in sdims}, "shape": spatial_shape},
        res_abs, nsdims=nsdims)
    store = {".zgroup": {"zarr_format": 2}}
    cur = data
    for lvl, rel in enumerate(res_rel):
        if max(rel.values()) > 1:
            cur = pyramid.coarsen(cur, [rel[d] if d in sdims else 1 for d in dims])
        store[f"{lvl}/.zarray"] = array_zarray(cur.shape, chunk_shape, cur.dtype)
        store.update(encode_array(cur, chunk_shape, prefix=f"{lvl}/"))
    ms = {"axes": axes, "datasets": [{"path": f"{l}", "coordinateTransformations": coordtfs[l]} for l in range(len(res_rel))],
          "name": "/", "version": "0.4"}
    zattrs = {"multiscales": [ms]}
    if "c" in dims:
        other = tuple(i for i, d in enumerate(dims) if d != "c")
        cmin, cmax = np.array(cur.min(axis=other)), np.array(cur.max(axis=other))
        labels = c_coords if c_coords is not None else list(range(len(cmin)))
        zattrs["omero"] = {"channels": [
            {"color": "ffffff", "label": f"{ch}", "active": True,
             "window": {"end": int(cmax[i]), "max": int(cmax[i]), "min": 0, "start": int(cmin[i])}}
            for i, ch in enumerate(labels)]}
    store[".zattrs"] = zattrs
    return store
