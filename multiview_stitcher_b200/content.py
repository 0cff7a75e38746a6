"""Multi-pass fusion with a ``weights_func`` (content-weighted fusion) or an
arbitrary ``fusion_func``: the structure of ``fuse_np`` (fusion/_core.py:1513-1733)
with every array step on the GPU.

    resample views + blending weights (mvs_resample_views)
      -> mask + normalize_weights          (_core.py:1647-1649)
      -> weights_func(transformed_views, blending_weights, ...)   (_core.py:1665-1680)
      -> fusion_func(transformed_views, blending_weights, fusion_weights, ...)
      -> trim halo, nan_to_num, cast        (_core.py:1687-1713)

The engine's own hooks (``hooks.content_based``, ``hooks.weighted_average_fusion``
...) keep the stacks on the device; foreign callables receive numpy arrays, as
they would from the reference.
"""

from __future__ import annotations

import ctypes
import inspect
import os

import numpy as np

from . import _lib, deconv, geometry, hooks
from ._lib import EngineError

_N_STREAMS = max(1, min(3, int(os.environ.get("MVS_CONTENT_STREAMS", "2"))))  # the engine keeps 3 workspaces
_TWO_STREAMS = _N_STREAMS > 1

_ENGINE_FUNCS = {
    "weighted_average_fusion": hooks.weighted_average_fusion,
    "max_fusion": hooks.max_fusion,
    "simple_average_fusion": hooks.simple_average_fusion,
    "content_based": hooks.content_based,
    "content_based_dct": hooks.content_based_dct,
    "multi_view_deconvolution": deconv.multi_view_deconvolution,
}


def _has_keyword(func, name):
    if func is None:
        return False
    try:
        return name in inspect.signature(func).parameters
    except (TypeError, ValueError):
        return False


def _resolve(func):
    """Engine implementation for the reference's built-ins (matched by name),
    else the callable itself (called with numpy arrays)."""
    if func is None:
        return None, True
    name = getattr(func, "__name__", None)
    mod = getattr(func, "__module__", "") or ""
    # the reference's own built-ins and this package's selectors share names
    if name in _ENGINE_FUNCS and mod.split(".")[0] in ("multiview_stitcher", "multiview_stitcher_b200"):
        return _ENGINE_FUNCS[name], True
    return func, False


def resample_stack(dviews, params, chunk_props, interpolation_order=1, full_view_bbs=None, spacings=None,
                   blending_widths=None, shrink_distance=0, want_weights=True):
    """(V, *chunk) float32 CUDA stacks of the transformed views (NaN outside) and
    their un-normalised, validity-masked blending weights."""
    import torch

    from .fusion import build_work_list

    lib = _lib.load(require_device=True)
    ndim = dviews[0].ndim
    dims = geometry.spatial_dims(ndim)
    shape = tuple(int(chunk_props["shape"][d]) for d in dims)
    work = build_work_list(
        dviews, params, chunk_props, dict(zip(dims, shape)), 0,
        [chunk_props["origin"][d] for d in dims], full_view_bbs, spacings, blending_widths, shrink_distance,
    )
    # build_work_list drops views that cannot touch the chunk; fuse_np keeps every
    # view it is given (all-NaN rows), so re-expand to the full list
    xarr, tables, vidx = work["xforms"], work["tables"], work["view_index"]
    V = len(dviews)
    all_kept = len(xarr) == V and list(vidx) == list(range(V))
    tv = bw = None
    if not all_kept:
        tv = torch.full((V,) + shape, float("nan"), dtype=torch.float32, device="cuda")
        bw = torch.zeros((V,) + shape, dtype=torch.float32, device="cuda") if want_weights else None
    if len(xarr):
        n = len(xarr)
        tv_c = torch.empty((n,) + shape, dtype=torch.float32, device="cuda")
        bw_c = torch.empty((n,) + shape, dtype=torch.float32, device="cuda") if want_weights else None
        shp = (ctypes.c_int32 * 3)(*((1,) * (3 - ndim) + shape))
        halo = (ctypes.c_int32 * 3)(0, 0, 0)
        _lib.check(
            lib.mvs_resample_views(
                xarr.ctypes.data_as(ctypes.c_void_p), n, tables.ctypes.data_as(ctypes.c_void_p), len(tables), shp, halo,
                ndim, int(interpolation_order), ctypes.c_void_p(tv_c.data_ptr()),
                ctypes.c_void_p(bw_c.data_ptr() if want_weights else 0), _lib.current_stream_ptr(),
            ),
            "mvs_resample_views",
        )
        if all_kept:
            return tv_c, bw_c  # every view touches the chunk: the stacks are already in view order
        idx = torch.tensor(vidx, device="cuda")
        tv[idx] = tv_c
        if want_weights:
            bw[idx] = bw_c
    return tv, bw


def _call(func, is_engine, kwargs):
    import torch

    if is_engine:
        return func(**kwargs)
    host = {k: (v.cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in kwargs.items()}
    res = func(**host)
    return torch.from_numpy(np.ascontiguousarray(res, dtype=np.float32)).to("cuda")


def fuse_np_with_weights(
    sims, params, output_properties, fusion_func=None, fusion_func_kwargs=None, weights_func=None,
    weights_func_kwargs=None, trim_overlap_in_pixels=0, interpolation_order=1, full_view_bbs=None,
    spacings=None, blending_widths=None, shrink_distance=0, output_on_backend=False,
):
    """``fuse_np`` (fusion/_core.py:1513-1733) for the general case: any
    ``fusion_func`` / ``weights_func`` pair, halo + trim, on device stacks."""
    import torch

    from .fusion import to_device_view, weighted_average_fusion as _sel_wavg

    lib = _lib.load(require_device=True)
    dviews = [to_device_view(s) for s in sims]
    ndim = dviews[0].ndim
    dims = geometry.spatial_dims(ndim)
    in_dtype = {torch.uint8: np.uint8, torch.uint16: np.uint16, torch.float32: np.float32}[dviews[0].tensor.dtype]
    if fusion_func is None:
        fusion_func = _sel_wavg
    ffunc, f_engine = _resolve(fusion_func)
    wfunc, w_engine = _resolve(weights_func)
    fusion_func_kwargs = dict(fusion_func_kwargs or {})
    weights_func_kwargs = dict(weights_func_kwargs or {})
    params = [np.asarray(p, dtype=np.float64) for p in params]

    requires_bw = _has_keyword(ffunc, "blending_weights") or _has_keyword(wfunc, "blending_weights")
    tv, bw = resample_stack(
        dviews, params, output_properties, interpolation_order, full_view_bbs, spacings, blending_widths,
        shrink_distance, want_weights=requires_bw,
    )
    V = tv.shape[0]
    N = tv[0].numel()
    if requires_bw:
        _lib.check(lib.mvs_normalize_weights(ctypes.c_void_p(bw.data_ptr()), V, N, _lib.current_stream_ptr()), "mvs_normalize_weights")

    fusion_func_kwargs["transformed_views"] = tv
    if _has_keyword(ffunc, "params"):
        fusion_func_kwargs["params"] = params
    if requires_bw:
        fusion_func_kwargs["blending_weights"] = bw
    if _has_keyword(ffunc, "output_spacing") and "output_spacing" not in fusion_func_kwargs:
        fusion_func_kwargs["output_spacing"] = output_properties["spacing"]
    if wfunc is not None and _has_keyword(ffunc, "fusion_weights"):
        weights_func_kwargs["transformed_views"] = tv
        if _has_keyword(wfunc, "params"):
            weights_func_kwargs["params"] = params
        if _has_keyword(wfunc, "blending_weights"):
            weights_func_kwargs["blending_weights"] = bw
        if _has_keyword(wfunc, "output_chunksize") and "output_chunksize" not in weights_func_kwargs:
            weights_func_kwargs["output_chunksize"] = output_properties["shape"]
        fusion_func_kwargs["fusion_weights"] = _call(wfunc, w_engine, weights_func_kwargs)
    fused = _call(ffunc, f_engine, fusion_func_kwargs)
    if not isinstance(fused, torch.Tensor):
        fused = torch.from_numpy(np.ascontiguousarray(fused, dtype=np.float32)).to("cuda")
    fused = fused.to(torch.float32).contiguous()

    if not isinstance(trim_overlap_in_pixels, dict):
        trim_overlap_in_pixels = {d: int(trim_overlap_in_pixels) for d in dims}
    trim = [int(trim_overlap_in_pixels[d]) for d in dims]
    shape = tuple(fused.shape)
    out_shape = tuple(s - 2 * t for s, t in zip(shape, trim))
    tdt = {np.uint8: torch.uint8, np.uint16: torch.uint16, np.float32: torch.float32}[in_dtype]
    out = torch.empty(out_shape, dtype=tdt, device="cuda")
    shp = (ctypes.c_int32 * 3)(*((1,) * (3 - ndim) + shape))
    trm = (ctypes.c_int32 * 3)(*((0,) * (3 - ndim) + tuple(trim)))
    ostr = (ctypes.c_int64 * 3)(*((0,) * (3 - ndim) + tuple(int(s) for s in out.stride())))
    _lib.check(
        lib.mvs_trim_cast(ctypes.c_void_p(fused.data_ptr()), shp, trm, ctypes.c_void_p(out.data_ptr()),
                          _lib.mvs_dtype(in_dtype), ostr, _lib.current_stream_ptr()),
        "mvs_trim_cast",
    )
    if output_on_backend:
        return out
    return out.cpu().numpy()


def fuse_with_weights(dviews, params, osp, output_chunksize, fusion_func, weights_func, weights_func_kwargs,
                      interpolation_order, blending_widths, overlap_in_pixels=None, chunk_subset=None):
    """Chunked multi-pass fusion of whole views: one ``fuse_np_with_weights`` per
    output chunk with the halo the hooks ask for (fusion/_core.py:1194-1254) and
    the views restricted to those that can touch the chunk (:582-653).
    ``chunk_subset``: linear indices of the chunks to fuse (sharded jobs); the rest of
    the returned stack stays zero."""
    import torch

    from .fusion import _required_overlap

    ndim = dviews[0].ndim
    dims = geometry.spatial_dims(ndim)
    full_shape = tuple(int(osp["shape"][d]) for d in dims)
    if output_chunksize is None:
        output_chunksize = geometry.DEFAULT_CHUNKSIZE_2D if ndim == 2 else geometry.DEFAULT_CHUNKSIZE_3D
    wfunc, _ = _resolve(weights_func)
    ffunc, _ = _resolve(fusion_func)
    # per-axis halo: the maximum over what the hooks ask for (fusion/_core.py:1194-1219)
    ov = np.zeros(ndim, dtype=np.int64)
    if overlap_in_pixels is None:
        for f, kw in ((wfunc if weights_func is not None else None, weights_func_kwargs), (ffunc, None)):
            o = _required_overlap(f, kw, output_chunksize)
            o = np.array([o[d] for d in dims] if isinstance(o, dict) else [o] * ndim, dtype=np.int64)
            ov = np.maximum(ov, o)
    elif isinstance(overlap_in_pixels, dict):
        ov = np.array([int(overlap_in_pixels[d]) for d in dims], dtype=np.int64)
    else:
        ov = np.full(ndim, int(overlap_in_pixels), dtype=np.int64)
    bbs = [v.bb() for v in dviews]
    o_org, o_sp, _ = geometry.bb_arrays(osp, dims)
    aabbs = [geometry.transformed_aabb(bb, p, dims) for bb, p in zip(bbs, params)]
    out = torch.zeros(full_shape, dtype=dviews[0].tensor.dtype, device="cuda")
    grid = geometry.chunk_grid(osp, output_chunksize)
    if chunk_subset is not None:
        grid = [grid[i] for i in chunk_subset]
    # Consecutive chunks alternate between two streams: nothing in a chunk's pipeline returns to the
    # host, so the FP64-bound Gaussian passes of one chunk run beside the HBM-bound resampling and
    # element-wise passes of the next (each stream has its own engine workspace).
    cur = torch.cuda.current_stream()
    streams = _side_streams() if len(grid) > 1 and _TWO_STREAMS else [cur]
    for st in streams:
        st.wait_stream(cur)
    for k, (start, shape) in enumerate(grid):
        with torch.cuda.stream(streams[k % len(streams)]):
            _fuse_one_chunk(dviews, params, osp, dims, start, shape, ov, o_org, o_sp, aabbs, bbs, fusion_func,
                            weights_func, weights_func_kwargs, interpolation_order, blending_widths, out)
    for st in streams:
        cur.wait_stream(st)
    return out


_SIDE = {}


def _side_streams():
    """Two long-lived streams per device (the engine keeps one workspace per stream in use)."""
    import torch

    dev = torch.cuda.current_device()
    if dev not in _SIDE:
        _SIDE[dev] = [torch.cuda.Stream() for _ in range(_N_STREAMS)]
    return _SIDE[dev]


def _fuse_one_chunk(dviews, params, osp, dims, start, shape, ov, o_org, o_sp, aabbs, bbs, fusion_func, weights_func,
                    weights_func_kwargs, interpolation_order, blending_widths, out):
    """One output chunk of ``fuse_with_weights`` (enqueued on the current stream)."""
    if True:
        start_a = np.array(start)
        c_org = (o_org + o_sp * start_a) - ov * o_sp
        hbb = {
            "origin": dict(zip(dims, map(float, c_org))),
            "spacing": osp["spacing"],
            "shape": {d: int(s) + 2 * int(h) for d, s, h in zip(dims, shape, ov)},
        }
        lo, hi = c_org, c_org + (np.array(shape) + 2 * ov - 1) * o_sp
        sel = [i for i, (alo, ahi) in enumerate(aabbs) if not (np.any(ahi < lo - 1e-6) or np.any(alo > hi + 1e-6))]
        if not sel:
            return
        res = fuse_np_with_weights(
            [dviews[i] for i in sel], [params[i] for i in sel], hbb, fusion_func=fusion_func,
            weights_func=weights_func, weights_func_kwargs=weights_func_kwargs,
            trim_overlap_in_pixels={d: int(h) for d, h in zip(dims, ov)},
            interpolation_order=interpolation_order, full_view_bbs=[bbs[i] for i in sel],
            blending_widths=blending_widths, output_on_backend=True,
        )
        sl = tuple(slice(int(s), int(s) + int(n)) for s, n in zip(start, shape))
        out[sl] = res
