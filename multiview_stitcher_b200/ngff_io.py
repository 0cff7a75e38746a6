"""Output / input side of the hot path: Zarr v2 chunk encode +
write of fused stacks and their pyramid levels, and chunk read + decode of input tiles,
with the (de)chunking done on the device.

What the reference does here is zarr-python under dask: ``fuse(output_zarr_url=...)`` opens
an array (fusion/_core.py:2160-2283), every fused block goes through
``da.to_zarr(region=...)`` (:2130-2150), then ``ngff_utils.write_sim_to_ome_zarr``
(ngff_utils.py:1564-1749, called at fusion/_core.py:1160-1168) mean-bins level after level
(``write_and_return_downsampled_sim`` :1288-1408, every chunk written --
``write_empty_chunks=True``, ``fill_value=0``) and writes the NGFF ``multiscales`` document
(``write_multiscales_metadata`` :1185-1230, ``calc_ngff_coordinate_transformations_and_axes``
:1493-1561).  The reference's own statement of the array metadata and of the bytes of a chunk
is ``VirtualOMEZarr.array_zarray`` / ``read_chunk`` / ``_pad_edge_chunk`` (:306-325, :372-395,
:425-436): C order, edge chunks padded to the chunk shape with the fill value, key
``i/j/k`` under ``dimension_separator="/"`` (:1258-1266).

Here a level stays in HBM: ``mvs_chunks_pack`` reorders it into chunk-major order (one
contiguous device range per chunk file), ``mvs_chunks_store`` moves each chunk through a
per-thread pinned buffer into its file while the other threads' DMAs and writes overlap;
the next level is one ``mvs_bin_mean`` launch over the resident one.  Reading is the mirror
image (``mvs_chunks_load`` + ``mvs_chunks_unpack``).

Zarr-python and numcodecs are not installed in this image: raw chunks (``compressor=None``)
take the device path; ``zlib`` / ``gzip`` chunks are (de)compressed by host threads around the
same packed buffer; Blosc / Zstd stores raise ``EngineError``.  Zarr v3 (NGFF 0.5) is not
written.
"""

from __future__ import annotations

import ctypes
import json
import math
import os
import shutil
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import _lib, geometry
from ._lib import EngineError

SPATIAL_DIMS = ("z", "y", "x")


# --- metadata -----------------------------------------------------------------------


def zarr_dtype(dtype):
    """numpy dtype -> Zarr v2 dtype string (ngff_utils.py:121-128)."""
    dtype = np.dtype(dtype)
    if dtype.byteorder == "=":
        dtype = dtype.newbyteorder("|" if dtype.itemsize == 1 else ("<" if np.little_endian else ">"))
    return dtype.str


def fill_value_for_dtype(dtype):
    """ngff_utils.py:131-139."""
    dtype = np.dtype(dtype)
    if np.issubdtype(dtype, np.floating):
        return 0.0
    if np.issubdtype(dtype, np.bool_):
        return False
    return 0


def array_zarray(shape, chunks, dtype, compressor=None):
    """The ``.zarray`` document of a level (``VirtualOMEZarr.array_zarray``,
    ngff_utils.py:306-325)."""
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


def calc_ngff_coordinate_transformations_and_axes(stack_properties_res0, res_abs_factors, nsdims=None, time_transform=None):
    """``ngff_utils.calc_ngff_coordinate_transformations_and_axes`` (ngff_utils.py:1493-1561):
    per level ``scale = spacing * factor`` and ``translation = origin + (factor - 1) *
    spacing / 2``; non-spatial axes carry the time calibration / unit scale."""
    spacing = stack_properties_res0["spacing"]
    origin = stack_properties_res0["origin"]
    sdims = list(spacing.keys())
    nsdims = list(nsdims or [])
    tt = {"scale": 1.0, "translation": 0.0, "unit": None, **(time_transform or {})}
    ns_scales = [float(tt["scale"]) if d == "t" else 1.0 for d in nsdims]
    ns_trans = [float(tt["translation"]) if d == "t" else 0 for d in nsdims]
    coordtfs = [
        [
            {"type": "scale", "scale": ns_scales + [float(s * f[d]) for d, s in spacing.items()]},
            {"type": "translation", "translation": ns_trans + [origin[d] + (f[d] - 1) * spacing[d] / 2 for d in sdims]},
        ]
        for f in res_abs_factors
    ]
    axes = []
    for d in nsdims + sdims:
        ax = {"name": d, "type": "channel" if d == "c" else ("time" if d == "t" else "space")}
        if d in sdims:
            ax["unit"] = "micrometer"
        if d == "t" and tt["unit"]:
            ax["unit"] = tt["unit"]
        axes.append(ax)
    return coordtfs, axes


def multiscales_zattrs(axes, datasets, name="/", ngff_version="0.4"):
    """The group attributes ``write_multiscales_metadata`` leaves behind for NGFF 0.4
    (ngff_utils.py:1185-1224): ``axes``, ``datasets``, ``name`` and ``version`` only."""
    if not str(ngff_version).startswith("0.4"):
        raise EngineError(f"ngff_version {ngff_version}: only NGFF 0.4 (Zarr v2) is written by the engine")
    ms = {
        "axes": [dict(a) for a in axes],
        "datasets": [
            {"path": str(ds["path"]),
             "coordinateTransformations": [_plain_transform(t) for t in ds["coordinateTransformations"]]}
            for ds in datasets
        ],
        "name": name,
        "version": str(ngff_version),
    }
    return {"multiscales": [ms]}


def _plain_transform(t):
    if t.get("type") == "scale":
        return {"scale": [float(v) for v in t["scale"]], "type": "scale"}
    if t.get("type") == "translation":
        return {"translation": [float(v) for v in t["translation"]], "type": "translation"}
    return {"type": "identity"}


def _write_json(path, obj):
    tmp = path + ".tmp"
    with open(tmp, "w") as f:
        json.dump(obj, f, indent=4)
    os.replace(tmp, path)


# --- compressors ---------------------------------------------------------------------


def _codec(config):
    """(encode, decode) host functions for a Zarr v2 compressor config, None for raw."""
    if config is None:
        return None
    if hasattr(config, "get_config"):
        config = config.get_config()
    cid = config.get("id")
    level = int(config.get("level", 1))
    if cid == "zlib":
        return (lambda b: zlib.compress(b, level)), zlib.decompress
    if cid == "gzip":
        import gzip

        return (lambda b: gzip.compress(b, compresslevel=level, mtime=0)), gzip.decompress
    raise EngineError(f"compressor {cid!r} is not available in this build (raw, zlib and gzip chunks only)")


def _compressor_config(compressor):
    if compressor is None:
        return None
    cfg = compressor.get_config() if hasattr(compressor, "get_config") else dict(compressor)
    _codec(cfg)
    return cfg


# --- the array ------------------------------------------------------------------------


class ZarrArray:
    """A Zarr v2 array in a directory store with ``/``-separated chunk keys.

    Host side it behaves like the ``output_zarr_array`` the reference hands to hook C:
    ``arr[region] = block`` / ``arr[region]`` with numpy semantics on basic slices.  Device side
    ``write_device`` / ``read_device`` move whole spatial boxes between HBM and the chunk
    files without a host copy of the dense array."""

    def __init__(self, path, meta):
        self.path = str(path)
        self.meta = meta
        self.shape = tuple(int(s) for s in meta["shape"])
        self.chunks = tuple(int(c) for c in meta["chunks"])
        self.dtype = np.dtype(meta["dtype"])
        self.ndim = len(self.shape)
        self.fill_value = meta.get("fill_value") or 0
        self.sep = meta.get("dimension_separator", ".")
        if meta.get("order", "C") != "C":
            raise EngineError("only C-order Zarr arrays are supported")
        if meta.get("filters"):
            raise EngineError("Zarr filters are not supported")
        self._codec = _codec(meta.get("compressor"))
        self.grid = tuple(-(-s // c) for s, c in zip(self.shape, self.chunks))
        self.chunk_bytes = int(np.prod(self.chunks)) * self.dtype.itemsize
        self.bytes_written = 0
        self.bytes_read = 0

    # -- creation ------------------------------------------------------------------
    @classmethod
    def create(cls, path, shape, chunks, dtype, compressor=None, overwrite=True, **_ignored):
        """``zarr.open(path, shape=, chunks=, dtype=, fill_value=0, mode="w",
        dimension_separator="/", zarr_format=2)`` (fusion/_core.py:2256-2283,
        ngff_utils.py:1353-1362)."""
        path = str(path)
        if os.path.exists(path):
            if not overwrite:
                raise EngineError(f"{path} exists")
            shutil.rmtree(path)
        os.makedirs(path)
        chunks = tuple(int(c) for c in chunks)  # kept as declared, like zarr.create (edge chunks are padded)
        meta = array_zarray(shape, chunks, dtype, _compressor_config(compressor))
        _write_json(os.path.join(path, ".zarray"), meta)
        return cls(path, meta)

    @classmethod
    def open(cls, path):
        p = os.path.join(str(path), ".zarray")
        if not os.path.exists(p):
            raise EngineError(f"{path} is not a Zarr v2 array (no .zarray)")
        with open(p) as f:
            return cls(path, json.load(f))

    # -- chunk files ---------------------------------------------------------------
    def chunk_path(self, idx):
        return os.path.join(self.path, *[str(int(i)) for i in idx]) if self.sep == "/" else os.path.join(
            self.path, ".".join(str(int(i)) for i in idx))

    def _mkdirs(self, indices):
        if self.sep != "/":
            return
        for d in {os.path.dirname(self.chunk_path(i)) for i in indices}:
            os.makedirs(d, exist_ok=True)

    def _read_chunk(self, idx):
        p = self.chunk_path(idx)
        if not os.path.exists(p):
            return np.full(self.chunks, self.fill_value, dtype=self.dtype)
        with open(p, "rb") as f:
            raw = f.read()
        if self._codec is not None:
            raw = self._codec[1](raw)
        self.bytes_read += len(raw)
        return np.frombuffer(raw, dtype=self.dtype).reshape(self.chunks).copy()

    def _write_chunk(self, idx, chunk):
        raw = np.ascontiguousarray(chunk, dtype=self.dtype).tobytes(order="C")
        if self._codec is not None:
            raw = self._codec[0](raw)
        p = self.chunk_path(idx)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        with open(p, "wb") as f:
            f.write(raw)
        self.bytes_written += len(raw)

    # -- numpy-style host access ----------------------------------------------------
    def _region(self, key):
        if not isinstance(key, tuple):
            key = (key,)
        if any(k is Ellipsis for k in key):
            i = [k is Ellipsis for k in key].index(True)
            key = key[:i] + (slice(None),) * (self.ndim - len(key) + 1) + key[i + 1:]
        key = key + (slice(None),) * (self.ndim - len(key))
        lo, hi, squeeze = [], [], []
        for k, n in zip(key, self.shape):
            if isinstance(k, (int, np.integer)):
                k = int(k) + (n if k < 0 else 0)
                lo.append(k), hi.append(k + 1), squeeze.append(True)
            elif isinstance(k, slice):
                a, b, st = k.indices(n)
                if st != 1:
                    raise EngineError("ZarrArray supports unit-step slices only")
                lo.append(a), hi.append(max(a, b)), squeeze.append(False)
            else:
                raise EngineError(f"unsupported index {k!r}")
        return lo, hi, squeeze

    def _touched(self, lo, hi):
        ranges = [range(a // c, -(-b // c)) if b > a else range(0) for a, b, c in zip(lo, hi, self.chunks)]
        return list(np.ndindex(*[len(r) for r in ranges])), ranges

    def __setitem__(self, key, value):
        lo, hi, squeeze = self._region(key)
        value = np.asarray(value)
        full_shape = tuple(b - a for a, b in zip(lo, hi))
        sq_shape = tuple(n for n, s in zip(full_shape, squeeze) if not s)
        if value.shape not in (sq_shape, full_shape):
            value = np.broadcast_to(value, sq_shape)
        value = value.reshape(full_shape)
        rel, ranges = self._touched(lo, hi)
        for r in rel:
            idx = tuple(rg[i] for rg, i in zip(ranges, r))
            c0 = [i * c for i, c in zip(idx, self.chunks)]
            a = [max(l, c) for l, c in zip(lo, c0)]
            b = [min(h, c + cs, n) for h, c, cs, n in zip(hi, c0, self.chunks, self.shape)]
            covers = all(x == c and (y == c + cs or y == n) for x, y, c, cs, n in zip(a, b, c0, self.chunks, self.shape))
            chunk = np.full(self.chunks, self.fill_value, dtype=self.dtype) if covers else self._read_chunk(idx)
            chunk[tuple(slice(x - c, y - c) for x, y, c in zip(a, b, c0))] = value[tuple(slice(x - l, y - l) for x, y, l in zip(a, b, lo))]
            self._write_chunk(idx, chunk)

    def __getitem__(self, key):
        lo, hi, squeeze = self._region(key)
        out = np.empty(tuple(b - a for a, b in zip(lo, hi)), dtype=self.dtype)
        rel, ranges = self._touched(lo, hi)
        for r in rel:
            idx = tuple(rg[i] for rg, i in zip(ranges, r))
            c0 = [i * c for i, c in zip(idx, self.chunks)]
            a = [max(l, c) for l, c in zip(lo, c0)]
            b = [min(h, c + cs) for h, c, cs in zip(hi, c0, self.chunks)]
            out[tuple(slice(x - l, y - l) for x, y, l in zip(a, b, lo))] = self._read_chunk(idx)[
                tuple(slice(x - c, y - c) for x, y, c in zip(a, b, c0))]
        return out.reshape([n for n, s in zip(out.shape, squeeze) if not s])

    def __array__(self, dtype=None, copy=None):
        a = self[...]
        return a.astype(dtype) if dtype is not None else a

    # -- device access ---------------------------------------------------------------
    def _spatial_split(self, lead):
        lead = tuple(int(i) for i in lead)
        nsp = self.ndim - len(lead)
        if nsp not in (2, 3):
            raise EngineError(f"device access needs 2 or 3 trailing spatial axes, got {nsp}")
        if any(c != 1 for c in self.chunks[: len(lead)]):
            raise EngineError("device access needs chunk size 1 along the non-spatial axes")
        return lead, nsp

    def _box_chunks(self, lead, start, shape, nsp):
        """chunk index tuples (C order over the box's chunk grid) of a chunk-aligned box."""
        sch, ssh = self.chunks[-nsp:], self.shape[-nsp:]
        for a, n, c, s in zip(start, shape, sch, ssh):
            if a % c or not ((a + n) % c == 0 or a + n == s) or a + n > s or n < 1:
                raise EngineError(f"device region start {tuple(start)} shape {tuple(shape)} is not aligned to chunks {sch} of {ssh}")
        first = [a // c for a, c in zip(start, sch)]
        counts = [-(-n // c) for n, c in zip(shape, sch)]
        return [lead + tuple(f + i for f, i in zip(first, r)) for r in np.ndindex(*counts)], counts

    def write_device(self, tensor, lead=(), start=None):
        """Encode a CUDA tensor (spatial box, contiguous x) into this array's chunk files:
        box start ``start`` (spatial, chunk aligned; default the origin), at the non-spatial
        index ``lead``.  Returns the bytes written."""
        import torch

        lib = _lib.load(require_device=True)
        lead, nsp = self._spatial_split(lead)
        if tensor.ndim != nsp or not tensor.is_cuda:
            raise EngineError(f"write_device needs a CUDA tensor with {nsp} axes")
        if _np_dtype_of(tensor) != self.dtype:
            raise EngineError(f"dtype mismatch: tensor {tensor.dtype}, array {self.dtype}")
        if tensor.stride(-1) != 1:
            tensor = tensor.contiguous()
        start = tuple(int(s) for s in (start if start is not None else (0,) * nsp))
        indices, counts = self._box_chunks(lead, start, tuple(tensor.shape), nsp)
        es = self.dtype.itemsize
        sch = self.chunks[-nsp:]
        n_chunks = len(indices)
        packed = torch.empty(n_chunks * self.chunk_bytes, dtype=torch.uint8, device=tensor.device)
        st = _lib.current_stream_ptr()
        sh3, str3, ch3 = _triples(tuple(tensor.shape), tuple(tensor.stride()), sch)
        _lib.check(lib.mvs_chunks_pack(ctypes.c_void_p(tensor.data_ptr()), es, sh3, str3, ch3,
                                       ctypes.c_void_p(packed.data_ptr()), st), "mvs_chunks_pack")
        self._mkdirs(indices)
        if self._codec is None:
            paths = (ctypes.c_char_p * n_chunks)(*[self.chunk_path(i).encode() for i in indices])
            _lib.check(lib.mvs_chunks_store(ctypes.c_void_p(packed.data_ptr()), self.chunk_bytes, n_chunks, paths, st),
                       "mvs_chunks_store")
            n = n_chunks * self.chunk_bytes
        else:
            host = packed.cpu().numpy().reshape(n_chunks, self.chunk_bytes)

            def enc(k):
                raw = self._codec[0](host[k].tobytes())
                with open(self.chunk_path(indices[k]), "wb") as f:
                    f.write(raw)
                return len(raw)

            with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as ex:
                n = sum(ex.map(enc, range(n_chunks)))
        self.bytes_written += n
        return n

    def read_device(self, lead=(), start=None, shape=None, device="cuda"):
        """Decode a chunk-aligned spatial box into a new CUDA tensor."""
        import torch

        lib = _lib.load(require_device=True)
        lead, nsp = self._spatial_split(lead)
        start = tuple(int(s) for s in (start if start is not None else (0,) * nsp))
        shape = tuple(int(s) for s in (shape if shape is not None else [n - a for n, a in zip(self.shape[-nsp:], start)]))
        indices, counts = self._box_chunks(lead, start, shape, nsp)
        n_chunks = len(indices)
        es = self.dtype.itemsize
        out = torch.empty(shape, dtype=_torch_dtype_of(self.dtype), device=device)
        st = _lib.current_stream_ptr()
        if self._codec is None:
            packed = torch.empty(n_chunks * self.chunk_bytes, dtype=torch.uint8, device=out.device)
            paths = (ctypes.c_char_p * n_chunks)(*[self.chunk_path(i).encode() for i in indices])
            _lib.check(lib.mvs_chunks_load(ctypes.c_void_p(packed.data_ptr()), self.chunk_bytes, n_chunks, paths, st),
                       "mvs_chunks_load")
        else:
            host = np.zeros((n_chunks, self.chunk_bytes), dtype=np.uint8)

            def dec(k):
                p = self.chunk_path(indices[k])
                if os.path.exists(p):
                    with open(p, "rb") as f:
                        raw = self._codec[1](f.read())
                    if len(raw) != self.chunk_bytes:
                        raise EngineError(f"{p}: decoded chunk has {len(raw)} bytes, expected {self.chunk_bytes}")
                    host[k] = np.frombuffer(raw, dtype=np.uint8)

            with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as ex:
                list(ex.map(dec, range(n_chunks)))
            packed = torch.from_numpy(host.reshape(-1)).to(out.device)
        sh3, str3, ch3 = _triples(shape, tuple(out.stride()), self.chunks[-nsp:])
        _lib.check(lib.mvs_chunks_unpack(ctypes.c_void_p(packed.data_ptr()), es, sh3, str3, ch3,
                                         ctypes.c_void_p(out.data_ptr()), st), "mvs_chunks_unpack")
        self.bytes_read += n_chunks * self.chunk_bytes
        return out


def _triples(shape, stride, chunk):
    pad = 3 - len(shape)
    sh = (ctypes.c_int32 * 3)(*([1] * pad + [int(s) for s in shape]))
    big = int(shape[0]) * int(stride[0]) if len(shape) else 1
    st = (ctypes.c_int64 * 3)(*([big] * pad + [int(s) for s in stride]))
    ch = (ctypes.c_int32 * 3)(*([1] * pad + [int(c) for c in chunk]))
    return sh, st, ch


def _np_dtype_of(tensor):
    import torch

    table = {torch.uint8: np.uint8, torch.uint16: np.uint16, torch.float32: np.float32, torch.int16: np.int16,
             torch.int32: np.int32, torch.float64: np.float64, torch.int8: np.int8, torch.uint32: np.uint32}
    if tensor.dtype not in table:
        raise EngineError(f"unsupported tensor dtype {tensor.dtype}")
    return np.dtype(table[tensor.dtype])


def _torch_dtype_of(dtype):
    import torch

    table = {"u1": torch.uint8, "u2": torch.uint16, "f4": torch.float32, "i2": torch.int16, "i4": torch.int32,
             "f8": torch.float64, "i1": torch.int8, "u4": torch.uint32}
    key = np.dtype(dtype).str.lstrip("<|=")
    if np.dtype(dtype).byteorder == ">" or key not in table:
        raise EngineError(f"unsupported Zarr dtype {np.dtype(dtype).str}")
    return table[key]


# --- OME-Zarr writer / reader -----------------------------------------------------------


def _as_levels_input(image, dims):
    """(tensor on the device, dims list, origin, spacing) of a DeviceView / dict / array."""
    import torch

    from .fusion import DeviceView

    if isinstance(image, DeviceView):
        return image.tensor, list(dims or image.dims), dict(image.origin), dict(image.spacing), {}
    if isinstance(image, dict):
        data = image["data"]
        t = data if isinstance(data, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(data))
        dims = list(dims or image.get("dims") or geometry.spatial_dims(t.ndim))
        coords = {k: image[k] for k in ("t_coords", "c_coords") if k in image}
        return t.cuda(), dims, dict(image["origin"]), dict(image["spacing"]), coords
    raise EngineError(f"cannot interpret image of type {type(image)}")


def write_sim_to_ome_zarr(image, output_zarr_url, downscale_factors_per_spatial_dim=None, overwrite=False,
                          ngff_version="0.4", zarr_array_creation_kwargs=None, chunks=None, dims=None,
                          time_transform=None, min_shape=100):
    """``ngff_utils.write_sim_to_ome_zarr`` (ngff_utils.py:1564-1749) for a stack resident in
    HBM: every resolution level is mean-binned from the previous one on the device
    (``pyramid.downsample`` = ``mean_dtype`` under ``da.coarsen(trim_excess=True)``,
    :1284-1330), chunk-encoded and written; existing levels are kept and read back when
    ``overwrite`` is false (:1306-1308) -- that is how ``fuse(output_zarr_url=...,
    zarr_options={"ome_zarr": True})`` completes the pyramid over the level-0 array its blocks
    were written to (fusion/_core.py:1160-1168).  The NGFF 0.4 ``multiscales`` (and ``omero``
    when there is a channel axis, :1714-1747) attributes are rewritten in any case.

    ``image``: ``DeviceView`` (spatial) or ``{"data": tensor / ndarray, "dims": [...], "origin":
    {...}, "spacing": {...}[, "c_coords": [...]]}`` with leading non-spatial axes.  ``chunks``:
    ``{dim: size}`` for the spatial axes (default: the reference's 256^3 / 2048^2,
    spatial_image_utils.py:21-22); non-spatial axes are chunked by 1.  Returns a dict with the
    level arrays (``ZarrArray``), their shapes and the bytes written."""
    import torch

    from .pairs import bin_view
    from .fusion import DeviceView
    from .pyramid import calc_resolution_levels

    if not str(ngff_version).startswith("0.4"):
        raise EngineError(f"ngff_version {ngff_version}: only NGFF 0.4 (Zarr v2) is written by the engine")
    kw = dict(zarr_array_creation_kwargs or {})
    compressor = kw.get("compressor")
    data, dims, origin, spacing, coords = _as_levels_input(image, dims)
    sdims = [d for d in dims if d in SPATIAL_DIMS]
    nsdims = [d for d in dims if d not in SPATIAL_DIMS]
    if dims != nsdims + sdims or len(sdims) not in (2, 3):
        raise EngineError(f"dims {dims}: non-spatial axes must lead, followed by (z,) y, x")
    default_chunks = {"z": 256, "y": 256, "x": 256} if len(sdims) == 3 else {"y": 2048, "x": 2048}
    # _chunk_shape_from_sim (ngff_utils.py:142-185): the level-0 chunk shape, no larger than the level-0
    # extent, is used for every level
    chunks = {d: min(int((chunks or default_chunks)[d]), int(data.shape[dims.index(d)])) for d in sdims}
    url = str(output_zarr_url)
    if overwrite and os.path.exists(url):
        shutil.rmtree(url)
    os.makedirs(url, exist_ok=True)
    spatial_shape = {d: int(data.shape[dims.index(d)]) for d in sdims}
    res_shapes, res_rel, res_abs = calc_resolution_levels(spatial_shape, downscale_factors_per_spatial_dim, min_shape)
    coordtfs, axes = calc_ngff_coordinate_transformations_and_axes(
        {"spacing": {d: spacing[d] for d in sdims}, "origin": {d: origin[d] for d in sdims}, "shape": spatial_shape},
        res_abs, nsdims=nsdims, time_transform=time_transform)
    ns_shape = tuple(int(data.shape[i]) for i in range(len(nsdims)))
    ns_indices = list(np.ndindex(*ns_shape)) if nsdims else [()]
    cur = {idx: data[idx] for idx in ns_indices}  # per non-spatial index: the current level's spatial tensor
    arrays, written = [], 0
    lo = hi = None
    for lvl, (shape_l, rel) in enumerate(zip(res_shapes, res_rel)):
        path = os.path.join(url, str(lvl))
        full_shape = ns_shape + tuple(shape_l[d] for d in sdims)
        full_chunks = (1,) * len(nsdims) + tuple(chunks[d] for d in sdims)
        if not overwrite and os.path.exists(os.path.join(path, ".zarray")):
            arr = ZarrArray.open(path)  # "Found existing resolution level": read it back as the next base
            if arr.shape != full_shape:
                raise EngineError(f"existing level {lvl} has shape {arr.shape}, expected {full_shape}")
            cur = {idx: arr.read_device(idx) for idx in ns_indices}
        else:
            arr = ZarrArray.create(path, full_shape, full_chunks, _np_dtype_of(data), compressor=compressor)
            for idx in ns_indices:
                t = cur[idx]
                if any(v > 1 for v in rel.values()):
                    dv = DeviceView(t, {d: 0.0 for d in sdims}, {d: 1.0 for d in sdims})
                    t = bin_view(dv, {d: int(rel[d]) for d in sdims}, skip_nan=False)
                    cur[idx] = t
                written += arr.write_device(t, lead=idx)
        arrays.append(arr)
    if "c" in nsdims:  # contrast limits from the last level, per channel (ngff_utils.py:1714-1747)
        ci = nsdims.index("c")
        n_c = ns_shape[ci]
        lo, hi = [None] * n_c, [None] * n_c
        for idx in ns_indices:
            t = cur[idx].float() if cur[idx].dtype == torch.uint16 else cur[idx]
            a, b = float(t.min()), float(t.max())
            k = idx[ci]
            lo[k] = a if lo[k] is None else min(lo[k], a)
            hi[k] = b if hi[k] is None else max(hi[k], b)
    datasets = [{"path": f"{lvl}", "coordinateTransformations": coordtfs[lvl]} for lvl in range(len(res_shapes))]
    zattrs = multiscales_zattrs(axes, datasets, name="/", ngff_version=ngff_version)
    if "c" in nsdims:
        labels = coords.get("c_coords", list(range(len(lo))))
        zattrs["omero"] = {"channels": [
            {"color": "ffffff", "label": f"{ch}", "active": True,
             "window": {"end": int(hi[i]), "max": int(hi[i]), "min": 0, "start": int(lo[i])}}
            for i, ch in enumerate(labels)]}
    _write_json(os.path.join(url, ".zgroup"), {"zarr_format": 2})
    _write_json(os.path.join(url, ".zattrs"), zattrs)
    return {"arrays": arrays, "shapes": res_shapes, "abs_factors": res_abs, "bytes_written": written,
            "zattrs": zattrs}


def read_sim_from_ome_zarr(zarr_path, resolution_level=0, lead=None):
    """``ngff_utils.read_sim_from_ome_zarr`` (ngff_utils.py:1752-1811) onto the device: the
    spatial stack of one resolution level (at the non-spatial index ``lead``, default all
    zeros) as a ``DeviceView`` with origin / spacing from the level's NGFF ``scale`` /
    ``translation`` -- the input-tile decode of the hot path."""
    from .fusion import DeviceView

    with open(os.path.join(str(zarr_path), ".zattrs")) as f:
        attrs = json.load(f)
    ms = (attrs.get("multiscales") or attrs.get("ome", {}).get("multiscales"))[0]
    names = [a["name"] for a in ms["axes"]]
    ds = ms["datasets"][int(resolution_level)]
    scale = next(t["scale"] for t in ds["coordinateTransformations"] if t["type"] == "scale")
    trans = next((t["translation"] for t in ds["coordinateTransformations"] if t["type"] == "translation"), [0.0] * len(names))
    arr = ZarrArray.open(os.path.join(str(zarr_path), ds["path"]))
    sdims = [d for d in names if d in SPATIAL_DIMS]
    n_ns = len(names) - len(sdims)
    lead = tuple(lead) if lead is not None else (0,) * n_ns
    t = arr.read_device(lead)
    origin = {d: float(trans[names.index(d)]) for d in sdims}
    spacing = {d: float(scale[names.index(d)]) for d in sdims}
    return DeviceView(t, origin, spacing)
