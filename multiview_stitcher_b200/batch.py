"""Hook C: a ``batch_func`` for ``fusion.fuse(output_zarr_url=..., batch_options=...)``
(fusion/_core.py:1133-1141) -- the one hook at which the
resampling itself can run inside the fused kernel without patching the reference.

The reference hands ``batch_func(fuse_chunk, block_ids, **batch_func_kwargs)`` a
``functools.partial`` of ``_fuse_chunk_to_zarr`` (fusion/_core.py:2044-2154) whose
keywords expose everything a chunk needs (views, transform key, output geometry,
destination array).  Where the reference would run one dask graph per block --
``fuse()`` on a one-chunk output stack, then ``da.to_zarr(region=...)`` -- the engine
fuses ALL blocks of a batch that share their non-spatial coordinate (t, c) in one
launch of the fused kernel (``FusionPlan(chunk_subset=...)``) over views that were
uploaded once, and writes the regions itself.

    from multiview_stitcher_b200.batch import BatchFuser
    fusion.fuse(sims, transform_key=..., output_zarr_url=...,
                batch_options={"batch_func": BatchFuser(), "n_batch": 64})
"""

from __future__ import annotations

import ctypes

import numpy as np

from . import geometry
from ._lib import EngineError

_BUILTIN = ("weighted_average_fusion", "max_fusion", "simple_average_fusion")


def _lib_copy_d2h(dst_array, src_tensor):
    from . import _lib

    return _lib.copy_d2h(dst_array, src_tensor)


def _select_ns(sim, ns_coord):
    """The spatial image at one non-spatial coordinate (the reference's
    ``sim_sel_coords(sim, {dim: sim.coords[dim][[ic]]})``, _core.py:2105-2111)."""
    if not ns_coord:
        return sim
    if not hasattr(sim, "isel"):
        raise EngineError("views with non-spatial dimensions must support .isel()")
    return sim.isel({d: int(i) for d, i in ns_coord.items()})


def _affine_at(sim, transform_key, ns_coord):
    """(ndim+1)^2 array of ``sim.attrs["transforms"][transform_key]`` at the block's time
    point (the transforms carry a "t" axis when the sim has one)."""
    if isinstance(sim, dict):
        return np.asarray(sim["transforms"][transform_key], dtype=np.float64)
    aff = sim.attrs["transforms"][transform_key]
    if hasattr(aff, "dims") and "t" in aff.dims:
        aff = aff.isel({"t": int(ns_coord.get("t", 0))})
    return np.asarray(getattr(aff, "data", aff), dtype=np.float64)


def block_geometry(osp, output_chunksize):
    """Spatial chunk index tuple -> (linear index into ``geometry.chunk_grid``, offset,
    shape): the regular grid ``normalize_chunks`` gives ``_fuse_chunk_to_zarr``
    (_core.py:2065-2090)."""
    dims = geometry.spatial_dims(len(osp["shape"]))
    grid = geometry.chunk_grid(osp, output_chunksize)
    counts = [-(-int(osp["shape"][d]) // int(output_chunksize[d])) for d in dims]
    table = {}
    for lin, idx in enumerate(np.ndindex(*counts)):
        table[tuple(int(i) for i in idx)] = (lin, grid[lin][0], grid[lin][1])
    return table


def _host_array(data):
    """numpy view of a host buffer (numpy array or CPU torch tensor), else None."""
    import torch

    if isinstance(data, np.ndarray):
        return data
    if isinstance(data, torch.Tensor) and not data.is_cuda:
        return data.numpy()
    return None


class BatchFuser:
    """Callable ``batch_func``.

    Built-in fusion functions run as a host-to-host pipeline per call: the blocks of the
    batch are ordered into bands along the slowest axis; the views a band reads are sent
    up first (pageable numpy arrays through the engine's pinned staging ring,
    ``mvs_copy_h2d_2d``), the band is fused by ONE launch over all its blocks as soon as
    they have landed, and finished bands stream back into the destination array
    (``mvs_copy_d2h_2d``) from a second thread while later views are still going up.
    Views stay resident for the (t, c) slice they belong to (consecutive batches of one
    ``fuse`` walk the blocks of a slice before moving on), so every view crosses PCIe once
    per slice, not once per batch; device buffers and plans are reused across slices of the
    same geometry.  A ``weights_func`` or a foreign ``fusion_func`` takes the multi-pass
    device path per block."""

    def __init__(self):
        self._key = None
        self._sims_ref = None   # strong reference: ids in the key stay unique while cached
        self._views = None      # DeviceViews of the current slice
        self._host = None       # per view: host ndarray still to be uploaded, or None
        self._events = None
        self._params = None
        self._plans = {}
        self._outbuf = {}
        self._streams = None
        self._pool = None
        self.launches = 0
        self.blocks_written = 0
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def reset(self):
        """Forget the resident views (next call uploads again); buffers and plans stay."""
        self._key = None

    # -- views of one (t, c) slice ---------------------------------------------------
    def _slice_views(self, fuse_kwargs, ns_coord, zarr_out):
        import torch

        from .fusion import DeviceView, _np_to_torch, _view_fields

        sims = fuse_kwargs.get("images")
        if sims is None:
            sims = fuse_kwargs.get("sims")
        if sims is None:
            raise EngineError("batch_func: fuse_kwargs carries no views (zarr-serialised sims are not supported)")
        tkey = fuse_kwargs["transform_key"]
        key = (tuple(id(s) for s in sims), tkey, tuple(sorted(ns_coord.items())), id(zarr_out))
        if key == self._key:
            return self._views, self._params
        self._sims_ref = (list(sims), zarr_out)
        old = self._views or []
        views, host = [], []
        for k, s in enumerate(sims):
            data, origin, spacing = _view_fields(_select_ns(s, ns_coord))
            harr = _host_array(data)
            if harr is None:
                # device-resident input (CUDA tensor, or anything exposing __cuda_array_interface__
                # such as the CuPy arrays of fuse(backend="cupy"), _core.py:1579-1587): used in place
                t = data if isinstance(data, torch.Tensor) else torch.as_tensor(data, device="cuda")
                views.append(DeviceView(t, origin, spacing))
                host.append(None)
                continue
            if harr.ndim and harr.strides[-1] != harr.itemsize:
                harr = np.ascontiguousarray(harr)
            tdt = _np_to_torch(harr.dtype)
            if k < len(old) and self._host is not None and tuple(old[k].tensor.shape) == harr.shape and old[k].tensor.dtype == tdt:
                t = old[k].tensor  # same geometry as the previous slice: plans keep their pointers
            else:
                t = torch.empty(harr.shape, dtype=tdt, device="cuda")
            views.append(DeviceView(t, origin, spacing))
            host.append(harr)
        self._views, self._host = views, host
        self._events = [None] * len(views)
        self._params = [_affine_at(s, tkey, ns_coord) for s in sims]
        self._key = key
        return self._views, self._params

    def _upload(self, vi, h2d):
        """Enqueue the upload of view vi (if still pending) on the copy stream."""
        import torch

        from . import _lib

        if self._host[vi] is None or self._events[vi] is not None:
            return
        self.h2d_bytes += _lib.copy_h2d(self._views[vi].tensor, self._host[vi], ctypes.c_void_p(h2d.cuda_stream))
        ev = torch.cuda.Event()
        ev.record(h2d)
        self._events[vi] = ev

    def __call__(self, fuse_chunk, block_ids, **_ignored):
        kw = fuse_chunk.keywords
        osp = kw["output_stack_properties"]
        nsdims = list(kw["nsdims"])
        fk = kw["fuse_kwargs"]
        chunksize = kw["output_chunksize"]
        zarr_out = kw["output_zarr_array"]
        dims = geometry.spatial_dims(len(osp["shape"]))
        if fk.get("backend") not in (None, "numpy", "cupy"):
            raise EngineError(f"batch_func: unknown backend {fk.get('backend')!r}")
        table = block_geometry(osp, chunksize)

        by_slice = {}
        for bid in block_ids:
            bid = tuple(int(b) for b in bid)
            by_slice.setdefault(bid[: len(nsdims)], []).append(bid[len(nsdims):])
        for ns_idx, spatial in by_slice.items():
            ns_coord = dict(zip(nsdims, ns_idx))
            views, params = self._slice_views(fk, ns_coord, zarr_out)
            blocks = sorted((table[s] for s in spatial), key=lambda b: b[0])
            fusion_func, weights_func = fk.get("fusion_func"), fk.get("weights_func")
            if weights_func is None and (fusion_func is None or getattr(fusion_func, "__name__", None) in _BUILTIN):
                self._fuse_blocks_pipelined(views, params, osp, chunksize, fk, blocks, dims, ns_idx, len(nsdims), zarr_out)
                continue
            for vi in range(len(views)):  # multi-pass path: everything resident first
                self._upload_now(vi)
            fused = self._fuse_blocks_multipass(views, params, osp, fk, blocks, dims)  # device tensors
            from .ngff_io import ZarrArray

            dev_store = (isinstance(zarr_out, ZarrArray) and zarr_out._codec is None
                         and tuple(zarr_out.chunks) == (1,) * len(nsdims) + tuple(int(chunksize[d]) for d in dims))
            for (lin, start, shape), data in zip(blocks, fused):
                if isinstance(zarr_out, np.ndarray):
                    region = tuple(int(i) for i in ns_idx) + tuple(slice(int(a), int(a) + int(n)) for a, n in zip(start, shape))
                    self.d2h_bytes += _lib_copy_d2h(zarr_out[region], data)
                elif dev_store:
                    self.d2h_bytes += zarr_out.write_device(data, lead=ns_idx, start=start)
                else:
                    region = tuple(slice(i, i + 1) for i in ns_idx) + tuple(slice(int(a), int(a) + int(n)) for a, n in zip(start, shape))
                    zarr_out[region] = data.cpu().numpy().reshape((1,) * len(nsdims) + tuple(shape))
                self.blocks_written += 1

    def _upload_many(self, vis, h2d):
        """Enqueue the upload of all still-pending views of ``vis`` as ONE staged transfer (the copy
        pool keeps working across tiles); they share one completion event."""
        import torch

        from . import _lib

        todo = [vi for vi in vis if self._host[vi] is not None and self._events[vi] is None]
        if not todo:
            return
        if len(todo) == 1 or not all(self._host[vi].flags.c_contiguous and self._views[vi].tensor.is_contiguous() for vi in todo):
            for vi in todo:
                self._upload(vi, h2d)
            return
        self.h2d_bytes += _lib.copy_h2d_many([self._views[vi].tensor for vi in todo], [self._host[vi] for vi in todo],
                                             ctypes.c_void_p(h2d.cuda_stream))
        ev = torch.cuda.Event()
        ev.record(h2d)
        for vi in todo:
            self._events[vi] = ev

    def _upload_now(self, vi):
        import torch

        if self._streams is None:
            self._streams = (torch.cuda.Stream(), torch.cuda.Stream())
        self._upload(vi, self._streams[0])
        if self._events[vi] is not None:
            torch.cuda.current_stream().wait_event(self._events[vi])

    def _fuse_blocks_pipelined(self, views, params, osp, chunksize, fk, blocks, dims, ns_idx, n_ns, zarr_out):
        import torch
        from concurrent.futures import ThreadPoolExecutor

        from . import _lib
        from .fusion import FusionPlan, _np_to_torch, _torch_to_np

        ndim = len(dims)
        if self._streams is None:
            self._streams = (torch.cuda.Stream(), torch.cuda.Stream())
        if self._pool is None:
            self._pool = ThreadPoolExecutor(max_workers=1)
        h2d, d2h = self._streams
        cur = torch.cuda.current_stream()
        # device buffer: bounding box of the batch's blocks
        lo = np.min([b[1] for b in blocks], axis=0)
        hi = np.max([np.add(b[1], b[2]) for b in blocks], axis=0)
        np_dt = _torch_to_np(views[0].tensor.dtype)
        okey = (tuple(int(v) for v in hi - lo), np.dtype(np_dt).str)
        out = self._outbuf.get(okey)
        if out is None:
            self._outbuf.clear()  # one buffer at a time
            out = torch.zeros(okey[0], dtype=_np_to_torch(np_dt), device="cuda")
            self._outbuf[okey] = out
        pkey = (tuple(v.tensor.data_ptr() for v in views), tuple(p.tobytes() for p in params),
                tuple(b[0] for b in blocks), tuple(int(v) for v in lo), out.data_ptr(),
                getattr(fk.get("fusion_func"), "__name__", None), fk.get("interpolation_order", 1),
                repr(fk.get("blending_widths")), repr(sorted(osp["origin"].items())), repr(sorted(osp["shape"].items())))
        plan = self._plans.get(pkey)
        if plan is None:
            for p_ in self._plans.values():
                p_.close()
            self._plans.clear()
            plan = FusionPlan(views, params, osp, output_chunksize=chunksize, fusion_func=fk.get("fusion_func"),
                              interpolation_order=fk.get("interpolation_order", 1), blending_widths=fk.get("blending_widths"),
                              chunk_subset=[b[0] for b in blocks], out=out, out_start=lo)
            self._plans[pkey] = plan
        host_out = zarr_out if isinstance(zarr_out, np.ndarray) else None
        from .ngff_io import ZarrArray

        # the engine's own Zarr v2 array: blocks are chunk-encoded on the device and go from HBM to
        # their chunk files without a dense host copy (raw chunks; the output chunk grid is the array's)
        dev_store = (isinstance(zarr_out, ZarrArray) and zarr_out._codec is None
                     and tuple(zarr_out.chunks) == (1,) * n_ns + tuple(int(chunksize[d]) for d in dims))
        h2d.wait_stream(cur)
        d2h_ptr = ctypes.c_void_p(d2h.cuda_stream)
        futures = []

        def drain(done, band_blocks):
            torch.cuda.set_device(out.device)
            d2h.wait_event(done)
            n = 0
            if host_out is not None and len(band_blocks) > 1:
                # blocks that tile a box (a band of a regular chunk grid) go down as ONE pitched copy:
                # long rows and a deep pipeline instead of a short transfer per block
                b_lo = np.min([b[1] for b in band_blocks], axis=0)
                b_hi = np.max([np.add(b[1], b[2]) for b in band_blocks], axis=0)
                if int(np.prod(b_hi - b_lo)) == sum(int(np.prod(b[2])) for b in band_blocks):
                    band_blocks = [(None, b_lo, b_hi - b_lo)]
            for lin, start, shape in band_blocks:
                win = out[tuple(slice(int(a - o), int(a - o + m)) for a, o, m in zip(start, lo, shape))]
                region = tuple(int(i) for i in ns_idx) + tuple(slice(int(a), int(a) + int(m)) for a, m in zip(start, shape))
                if host_out is not None:
                    n += _lib.copy_d2h(host_out[region], win, d2h_ptr)
                elif dev_store:
                    with torch.cuda.stream(d2h):
                        n += zarr_out.write_device(win, lead=ns_idx, start=start)
                else:
                    with torch.cuda.stream(d2h):
                        arr = win.cpu().numpy()
                    zarr_out[tuple(slice(i, i + 1) for i in ns_idx) + region[n_ns:]] = arr.reshape((1,) * n_ns + tuple(shape))
                    n += arr.nbytes
            return n

        pos = 0
        for first, n, row0, nrows, vidx in plan.bands():
            self._upload_many(vidx, h2d)
            for vi in vidx:
                if self._events[vi] is not None:
                    cur.wait_event(self._events[vi])
            plan.run_chunks(first, n)
            self.launches += plan.launches_per_run
            done = torch.cuda.Event()
            done.record(cur)
            futures.append(self._pool.submit(drain, done, blocks[pos : pos + n]))
            pos += n
        for f in futures:
            self.d2h_bytes += f.result()
        self.blocks_written += len(blocks)

    def _fuse_blocks_multipass(self, views, params, osp, fk, blocks, dims):
        """Device tensors of the fused blocks, in order: any other fusion_func / a weights_func
        runs the multi-pass device path, one output stack per block exactly like the
        reference's per-chunk fuse() (_core.py:2118-2128).  Consecutive blocks alternate between
        two streams (see content.fuse_with_weights); nothing returns to the host in between."""
        import torch

        from . import content

        fusion_func, weights_func = fk.get("fusion_func"), fk.get("weights_func")
        order = fk.get("interpolation_order", 1)
        widths = fk.get("blending_widths")
        o_org = np.array([osp["origin"][d] for d in dims], dtype=np.float64)
        o_sp = np.array([osp["spacing"][d] for d in dims], dtype=np.float64)
        cur = torch.cuda.current_stream()
        streams = content._side_streams() if len(blocks) > 1 and content._TWO_STREAMS else [cur]
        for st in streams:
            st.wait_stream(cur)
        res = []
        for k, (_, start, shape) in enumerate(blocks):
            c_org = np.array(start) * o_sp + o_org  # _core.py:2083-2086
            sub = {"origin": dict(zip(dims, map(float, c_org))), "spacing": dict(osp["spacing"]),
                   "shape": {d: int(n) for d, n in zip(dims, shape)}}
            with torch.cuda.stream(streams[k % len(streams)]):
                out = content.fuse_with_weights(
                    views, params, sub, {d: int(n) for d, n in zip(dims, shape)}, fusion_func, weights_func,
                    fk.get("weights_func_kwargs"), order, widths, fk.get("overlap_in_pixels"),
                )
            res.append(out)
        for st in streams:
            cur.wait_stream(st)
        return res

    def close(self):
        for p_ in self._plans.values():
            p_.close()
        self._plans.clear()
        if self._pool is not None:
            self._pool.shutdown()
            self._pool = None


def batch_func(fuse_chunk, block_ids, **batch_func_kwargs):
    """Stateless form: ``batch_options={"batch_func": batch_func}`` (views are uploaded
    per call; prefer one ``BatchFuser()`` per ``fuse``)."""
    return BatchFuser()(fuse_chunk, block_ids, **batch_func_kwargs)
