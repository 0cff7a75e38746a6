"""Multi-GPU sharding of the hot path (one process per GPU, torch.distributed).

Both stages shard into independent units:

* registration -- one unit per overlap pair (registration.py:2657-2664); pairs
  are dealt round-robin, results gathered as small host objects.  No data-path
  collective.
* fusion -- one unit per output chunk (fusion/_core.py:1133-1141; the
  reference's own precedent for disjoint chunk sets per worker is
  browser/executors.py:232-288).
  - ``fuse_sharded``: bands of the chunk grid per rank, every rank holding
    (replicas of) the tiles its band touches -> no communication; the optional
    gather broadcasts each rank's band in the output dtype.
  - ``fuse_tile_partitioned``: the TILES are partitioned (each tile lives on
    exactly one GPU; this is how a stack larger than one GPU's HBM is fused).
    Every output chunk has an owner rank.  Chunks fed by one rank only are fused
    straight into the owner's slab.  For a chunk that draws from tiles on
    several GPUs only the box the foreign tiles can reach is exchanged, in one
    of two ways (one NCCL send/recv message per rank pair over NVLink, issued
    BEFORE the direct launch so the transfer hides behind it):
      mode="partial"  every contributing rank produces un-normalised partial
        sums (sum_i v_i*b_i, sum_i b_i) of ITS tiles over that box (8 bytes per
        voxel) and the owner adds them to its own, divides and casts.  Valid
        because normalisation is linear: sum_i v_i b_i / sum_i b_i (the per-view
        normalisers of weights.py:340-345 cancel).  Within 1 LSB / 1e-6 of the
        one-GPU result.
      mode="halo" (default)  the contributing rank sends the raw WINDOW of each
        of its tiles that the box can sample (input dtype: 2 bytes per voxel for
        uint16, a quarter of the partial sums) and the owner fuses the box with
        the ordinary fused kernel from local tiles + received windows, in global
        view order -- the one-GPU result up to float32 rounding of the box origin
        (<= 1e-6 relative, <= 1 LSB for uint16), no extra passes.
    The rest of the chunk is fused directly.
  - ``fuse_partial``: whole-volume variant (all-reduce of full accumulators;
    small stacks, ``max_fusion``).
"""

from __future__ import annotations

import ctypes

import numpy as np

from . import _lib, geometry
from ._lib import EngineError


def world():
    """(rank, world_size) of the default process group, (0, 1) without one."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_round_robin(n_units, rank, world_size):
    """Indices of the units rank ``rank`` owns (pairs: registration.py:2657-2664)."""
    return list(range(rank, n_units, world_size))


def shard_slabs(n_units, rank, world_size):
    """Contiguous, balanced slab of ``range(n_units)`` for ``rank`` (output chunks
    in C order: a slab is a band of the chunk grid along the slowest axis)."""
    base, extra = divmod(n_units, world_size)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


def gather_objects(local, n_units, owned):
    """All ranks' per-unit results in unit order (host objects, tiny)."""
    import torch.distributed as dist

    rank, ws = world()
    if ws == 1:
        return list(local)
    parts = [None] * ws
    dist.all_gather_object(parts, (list(owned), list(local)))
    out = [None] * n_units
    for idx, vals in parts:
        for i, v in zip(idx, vals):
            out[i] = v
    return out


def register_pairs_sharded(fixed_list, moving_list, **kwargs):
    """Registers this rank's share of the pairs on its GPU and returns the full
    result list on every rank.  fixed_list / moving_list must be indexable on
    every rank (only the owned pairs are touched)."""
    from . import registration

    rank, ws = world()
    n = len(fixed_list)
    owned = shard_round_robin(n, rank, ws)
    local = registration.register_pairs([fixed_list[i] for i in owned], [moving_list[i] for i in owned], **kwargs) if owned else []
    return gather_objects(local, n, owned)


def register_views_sharded(views, affines, pairs, **kwargs):
    """``pairs.register_views`` with the pairs dealt round-robin over the ranks: every rank
    plans, uploads (only the views its pairs touch), prepares and registers its share and
    all ranks receive the full result list.  ``views`` must be indexable on every rank."""
    from . import pairs as pairs_mod

    rank, ws = world()
    owned = shard_round_robin(len(pairs), rank, ws)
    local = pairs_mod.register_views(views, affines, [pairs[i] for i in owned], **kwargs) if owned else []
    return gather_objects(local, len(pairs), owned)


def _band_axis(n_per_axis):
    """Axis (not x) with the most chunks: bands along it balance best."""
    cand = list(range(len(n_per_axis) - 1))
    return max(cand, key=lambda a: (n_per_axis[a], -a))


class ShardedFuser:
    """This rank's band of output chunks, planned once and fused as often as needed (time lapses /
    channels share the geometry, like ``fusion.HostFuser``): tiles are replicated where bands meet,
    there is no communication.  ``out`` is the full-size output tensor with only the owned chunks
    written; its memory is laid out with the band axis slowest, so a band is one contiguous range."""

    def __init__(self, views, params, output_stack_properties, output_chunksize=None, band_axis=None, **plan_kwargs):
        import torch

        from .fusion import FusionPlan, _np_to_torch, _torch_to_np, to_device_view

        self.rank, self.ws = world()
        dviews = [to_device_view(v) for v in views]
        ndim = dviews[0].ndim
        dims = geometry.spatial_dims(ndim)
        if output_chunksize is None:
            output_chunksize = geometry.DEFAULT_CHUNKSIZE_2D if ndim == 2 else geometry.DEFAULT_CHUNKSIZE_3D
        cs = {d: int(output_chunksize[d]) for d in dims}
        osp = output_stack_properties
        full = [int(osp["shape"][d]) for d in dims]
        counts = [-(-full[i] // cs[d]) for i, d in enumerate(dims)]
        if band_axis is None:
            band_axis = _band_axis(counts)
        grid = geometry.chunk_grid(osp, cs)
        self.bands = [shard_slabs(counts[band_axis], r, self.ws) for r in range(self.ws)]
        mine = set(self.bands[self.rank])
        self.owned = [i for i, (start, _) in enumerate(grid) if start[band_axis] // cs[dims[band_axis]] in mine]
        # memory order: band axis first
        self.perm = [band_axis] + [a for a in range(ndim) if a != band_axis]
        inv = [self.perm.index(a) for a in range(ndim)]
        np_dt = np.dtype(plan_kwargs.pop("out_dtype", None) or _torch_to_np(dviews[0].tensor.dtype))
        self.out = torch.zeros([full[a] for a in self.perm], dtype=_np_to_torch(np_dt), device="cuda").permute(inv)
        self.band_axis, self.band_size, self.full = band_axis, cs[dims[band_axis]], full
        self.plan = None
        if self.owned:
            self.plan = FusionPlan(dviews, params, osp, output_chunksize=cs, chunk_subset=self.owned, out=self.out,
                                   out_dtype=np_dt, **plan_kwargs)
        self.last_gather_bytes = 0

    def run(self):
        if self.plan is not None:
            self.plan.run()
        return self.out

    def gather(self):
        """Broadcast every rank's band in the OUTPUT dtype (each rank ends up holding the whole
        stack; bytes on the wire = stack bytes x (N-1)/N per rank, no float32 round trip)."""
        import torch
        import torch.distributed as dist

        nbytes = 0
        if self.ws > 1:
            mem = self.out.permute(self.perm)  # contiguous, band axis leading
            for r in range(self.ws):
                if not self.bands[r]:
                    continue
                a = self.bands[r][0] * self.band_size
                b = min((self.bands[r][-1] + 1) * self.band_size, self.full[self.band_axis])
                slab = mem[a:b].view(torch.uint8)
                dist.broadcast(slab, src=r)
                if r != self.rank:
                    nbytes += slab.numel()
        self.last_gather_bytes = nbytes
        return self.out

    def close(self):
        if self.plan is not None:
            self.plan.close()
            self.plan = None


def fuse_sharded(views, params, output_stack_properties, output_chunksize=None, gather=False, band_axis=None,
                 **plan_kwargs):
    """One-shot ``ShardedFuser``: fuses this rank's band of output chunks (tiles replicated where
    bands meet, no communication).  Returns ``(out, owned_chunks)``; ``gather=True`` also
    broadcasts every rank's band in the output dtype."""
    f = ShardedFuser(views, params, output_stack_properties, output_chunksize, band_axis, **plan_kwargs)
    out = f.run()
    if gather:
        f.gather()
    fuse_sharded.last_gather_bytes = f.last_gather_bytes
    f.close()
    return out, f.owned


# --- tile-partitioned fusion ---------------------------------------------------


def _box_minus(box, sub):
    """``box`` minus ``sub`` (both (lo, hi) inclusive int arrays, sub inside box)
    as a list of disjoint boxes."""
    lo, hi = np.array(box[0]), np.array(box[1])
    out = []
    for a in range(len(lo)):
        if sub[0][a] > lo[a]:
            h = hi.copy()
            h[a] = sub[0][a] - 1
            out.append((lo.copy(), h))
            lo[a] = sub[0][a]
        if sub[1][a] < hi[a]:
            l = lo.copy()
            l[a] = sub[1][a] + 1
            out.append((l, hi.copy()))
            hi[a] = sub[1][a]
    return out


class TilePartition:
    """Host plan of a tile-partitioned fusion job; identical on every rank (it is a
    pure function of the geometry and the tile -> rank map).

    ``direct[r]``   boxes ``(start, shape)`` rank r fuses straight into its slab
    ``entries``     border boxes: ``{"chunk", "owner", "contrib" (non-owner ranks),
                    "start", "shape", "nvox"}``, ordered by (owner, contrib, chunk)
    ``owner_of``    owner rank per chunk; ``slab[r]`` = (start, shape) bounding the
                    chunks r owns
    """

    def __init__(self, view_bbs, params, owners, osp, chunksize, world_size):
        ndim = len(osp["shape"])
        dims = geometry.spatial_dims(ndim)
        self.dims, self.ndim, self.world_size = dims, ndim, int(world_size)
        o_org, o_sp, _ = geometry.bb_arrays(osp, dims)
        full = np.array([int(osp["shape"][d]) for d in dims], dtype=np.int64)
        cs = {d: int(chunksize[d]) for d in dims}
        self.chunksize = cs
        self.grid = geometry.chunk_grid(osp, cs)
        owners = [int(o) for o in owners]
        if len(owners) != len(view_bbs) or len(params) != len(view_bbs):
            raise EngineError("need one owner rank and one affine per view")
        if owners and (min(owners) < 0 or max(owners) >= world_size):
            raise EngineError("owner rank out of range")
        # output-pixel boxes the views can reach (one pixel of margin)
        vlo, vhi = [], []
        for bb, p in zip(view_bbs, params):
            alo, ahi = geometry.transformed_aabb(bb, p, dims)
            vlo.append(np.floor((alo - o_org) / o_sp - 1e-6).astype(np.int64) - 1)
            vhi.append(np.ceil((ahi - o_org) / o_sp + 1e-6).astype(np.int64) + 1)
        self.direct = [[] for _ in range(world_size)]
        self.owner_of = []
        entries = []
        n_chunks = len(self.grid)
        for ci, (start, shape) in enumerate(self.grid):
            clo = np.array(start, dtype=np.int64)
            chi = clo + np.array(shape, dtype=np.int64) - 1
            vol = np.zeros(world_size, dtype=np.float64)
            touching = []
            for vi in range(len(view_bbs)):
                lo, hi = np.maximum(vlo[vi], clo), np.minimum(vhi[vi], chi)
                if np.any(hi < lo):
                    continue
                touching.append((vi, lo, hi))
                vol[owners[vi]] += float(np.prod(hi - lo + 1))
            if not touching:
                owner = min(ci * world_size // max(n_chunks, 1), world_size - 1)
                self.owner_of.append(owner)
                self.direct[owner].append((tuple(start), tuple(shape)))
                continue
            owner = int(np.argmax(vol))  # first maximum: ties go to the lowest rank
            self.owner_of.append(owner)
            foreign = [(vi, lo, hi) for vi, lo, hi in touching if owners[vi] != owner]
            if not foreign:
                self.direct[owner].append((tuple(start), tuple(shape)))
                continue
            ulo = np.min([lo for _, lo, _ in foreign], axis=0)
            uhi = np.max([hi for _, _, hi in foreign], axis=0)
            for blo, bhi in _box_minus((clo, chi), (ulo, uhi)):
                self.direct[owner].append((tuple(int(v) for v in blo), tuple(int(v) for v in bhi - blo + 1)))
            contrib = tuple(sorted({owners[vi] for vi, _, _ in foreign}))
            ushape = uhi - ulo + 1
            entries.append({"chunk": ci, "owner": owner, "contrib": contrib, "start": tuple(int(v) for v in ulo),
                            "shape": tuple(int(v) for v in ushape), "nvox": int(np.prod(ushape)),
                            "foreign": [vi for vi, _, _ in foreign]})
        entries.sort(key=lambda e: (e["owner"], e["contrib"], e["chunk"]))
        self.entries = entries
        self.slab = []
        for r in range(world_size):
            mine = [self.grid[ci] for ci in range(n_chunks) if self.owner_of[ci] == r]
            if not mine:
                self.slab.append((tuple([0] * ndim), tuple([0] * ndim)))
                continue
            lo = np.min([np.array(s) for s, _ in mine], axis=0)
            hi = np.max([np.array(s) + np.array(n) for s, n in mine], axis=0)
            self.slab.append((tuple(int(v) for v in lo), tuple(int(v) for v in hi - lo)))
        self.full_shape = tuple(int(v) for v in full)
        # halo mode: per (owner rank, foreign view) the window of the view (its own pixel
        # indices, inclusive) that can be sampled from the owner's border boxes
        self.windows = {}
        for e in entries:
            c_org = o_org + o_sp * np.array(e["start"], dtype=np.float64)
            corners = np.array(list(np.ndindex(*([2] * ndim))), dtype=np.float64) * (np.array(e["shape"], dtype=np.float64) - 1)
            for vi in e["foreign"]:
                bb = view_bbs[vi]
                in_org, in_sp, in_n = geometry.bb_arrays(bb, dims)
                m, off = geometry.pixel_affine(np.linalg.inv(np.asarray(params[vi], dtype=np.float64)), c_org, o_sp, in_org, in_sp)
                pts = corners @ m.T + off
                lo = np.maximum(np.floor(pts.min(0)).astype(np.int64) - 2, 0)
                hi = np.minimum(np.ceil(pts.max(0)).astype(np.int64) + 2, in_n.astype(np.int64) - 1)
                if np.any(hi < lo):
                    continue
                lo[-1] = (lo[-1] // 16) * 16  # rows start 16-byte aligned for every dtype (TMA)
                key = (e["owner"], vi)
                if key in self.windows:
                    plo, phi = self.windows[key]
                    lo, hi = np.minimum(lo, plo), np.maximum(hi, phi)
                self.windows[key] = (lo, hi)
        self.view_owner = owners

    def halo_sends(self, src, dst):
        """Windows ``(view index, lo, hi)`` of ``src``'s tiles that ``dst`` needs for the
        border boxes it owns (halo mode), in view order."""
        return [(vi, lo, hi) for (o, vi), (lo, hi) in sorted(self.windows.items(), key=lambda kv: kv[0])
                if o == dst and self.view_owner[vi] == src]

    def halo_bytes(self, itemsize):
        """Raw tile bytes crossing NVLink per job in halo mode."""
        return sum(int(np.prod(hi - lo + 1)) * itemsize for (lo, hi) in self.windows.values())

    def own_entries(self, rank):
        return [e for e in self.entries if e["owner"] == rank]

    def send_entries(self, src, dst):
        """Entries rank ``src`` contributes partial sums to, owned by ``dst``."""
        return [e for e in self.entries if e["owner"] == dst and src in e["contrib"]]

    def exchanged_bytes(self):
        """float32 (num, den) bytes crossing NVLink per job, all ranks together."""
        return sum(8 * e["nvox"] * len(e["contrib"]) for e in self.entries)


class _Runner:
    """A prepared launch: ``run()`` enqueues it, ``launches`` kernels per run."""

    def __init__(self, plan=None):
        self.plan = plan
        self.launches = plan.launches_per_run if plan is not None else 0

    def run(self):
        if self.plan is not None:
            self.plan.run()

    def close(self):
        if self.plan is not None:
            self.plan.close()
            self.plan = None


class _CudaEngine:
    """Device half of ``TilePartitionedFuser`` (the CPU test substitutes an oracle-backed
    engine to exercise the exchange protocol over gloo)."""

    def __init__(self, **plan_kwargs):
        self.kw = plan_kwargs

    def zeros(self, n, np_dtype=np.float32):
        import torch

        from .fusion import _np_to_torch

        return torch.zeros(n, dtype=_np_to_torch(np.dtype(np_dtype)), device="cuda")

    def direct_plan(self, views, params, osp, chunksize, boxes, out, out_start):
        from .fusion import FusionPlan

        if not boxes or not views:
            return _Runner()
        return _Runner(FusionPlan(views, params, osp, output_chunksize=chunksize, chunk_list=boxes, out=out,
                                  out_start=out_start, **self.kw))

    def border_plan(self, views, params, full_bbs, osp, chunksize, boxes, out, out_start):
        """Border boxes fused from local views and received windows of foreign views
        (``full_bbs``: the bounding boxes of the WHOLE views, for the blending weights)."""
        from .fusion import FusionPlan

        if not boxes or not views:
            return _Runner()
        return _Runner(FusionPlan(views, params, osp, output_chunksize=chunksize, chunk_list=boxes, out=out,
                                  out_start=out_start, full_view_bbs=full_bbs, **self.kw))

    def view_tensor(self, view):
        return view.tensor

    def make_view(self, tensor, origin, spacing):
        from .fusion import DeviceView

        return DeviceView(tensor, origin, spacing)

    def partial_plan(self, views, params, osp, chunksize, boxes, targets):
        """boxes[i] accumulated into the packed float32 (num, den) windows targets[i] =
        (buffer tensor, element offset of num, element offset of den)."""
        from .fusion import FusionPlan

        if not boxes or not views:
            return _Runner()
        tg = []
        for (start, shape), (buf, o_num, o_den) in zip(boxes, targets):
            strides = [int(np.prod(shape[i + 1:])) for i in range(len(shape))]
            tg.append((buf.data_ptr() + 4 * o_num, buf.data_ptr() + 4 * o_den, strides))
        kw = {k: v for k, v in self.kw.items() if k != "fusion_func"}
        return _Runner(FusionPlan(views, params, osp, output_chunksize=chunksize, chunk_list=boxes, partial=True,
                                  chunk_targets=tg, **kw))

    def finalize(self, buf, items, out, out_start, np_dtype):
        """items: (element offset of num, of den, start, shape) per box."""
        if not items:
            return 0
        lib = _lib.load(require_device=True)
        ndim = out.ndim
        boxes = np.zeros(len(items), dtype=_lib.CHUNK_DTYPE)
        ostride = [0] * (3 - ndim) + [int(s) for s in out.stride()]
        for b, (o_num, o_den, start, shape) in zip(boxes, items):
            off = int(np.dot(np.asarray(start, dtype=np.int64) - np.asarray(out_start, dtype=np.int64), ostride[3 - ndim:]))
            b["out"] = out.data_ptr() + off * out.element_size()
            b["out_dtype"] = _lib.mvs_dtype(np_dtype)
            b["shape"] = [1] * (3 - ndim) + [int(n) for n in shape]
            b["stride"] = ostride
            b["acc_num"] = buf.data_ptr() + 4 * o_num
            b["acc_den"] = buf.data_ptr() + 4 * o_den
        _lib.check(lib.mvs_fuse_finalize_boxes(boxes.ctypes.data_as(ctypes.c_void_p), len(items), _lib.current_stream_ptr()),
                   "mvs_fuse_finalize_boxes")
        return 1


class TilePartitionedFuser:
    """Fusion of a stack whose tiles are partitioned over the ranks (see module docstring).

    ``local_views``: {global view index: view} for the views THIS rank holds (exactly the
    indices ``i`` with ``owners[i] == rank``); ``view_bbs`` / ``params`` / ``owners``: bounding
    box, affine and owner rank of EVERY view of the job (metadata, identical on all ranks).
    Weighted-average fusion with blending weights only.

    Construction plans the partition, allocates this rank's slab ``out`` (the chunks it
    owns; ``out_start`` = stack index of its first voxel), the packed partial-sum buffers
    and the launches; ``run()`` fuses: direct boxes -> partial sums of the border boxes ->
    one NCCL message per (contributor -> owner) pair -> add, divide, cast.
    """

    def __init__(self, local_views, view_bbs, params, owners, output_stack_properties, output_chunksize=None,
                 out_dtype=None, engine=None, partition=None, mode="halo", **plan_kwargs):
        if mode not in ("halo", "partial"):
            raise EngineError(f"unknown exchange mode {mode!r} (halo, partial)")
        self.mode = mode
        rank, ws = world()
        self.rank, self.ws = rank, ws
        osp = output_stack_properties
        ndim = len(osp["shape"])
        dims = geometry.spatial_dims(ndim)
        if output_chunksize is None:
            output_chunksize = geometry.DEFAULT_CHUNKSIZE_2D if ndim == 2 else geometry.DEFAULT_CHUNKSIZE_3D
        cs = {d: int(output_chunksize[d]) for d in dims}
        ff = plan_kwargs.get("fusion_func")
        if ff is not None and getattr(ff, "__name__", None) != "weighted_average_fusion":
            raise EngineError("tile-partitioned fusion supports weighted_average_fusion (use fuse_partial for max_fusion)")
        mine = sorted(i for i, o in enumerate(owners) if int(o) == rank)
        if sorted(local_views) != mine:
            raise EngineError(f"rank {rank} must hold exactly the views it owns: {mine}, got {sorted(local_views)}")
        self.engine = engine = engine or _CudaEngine(**plan_kwargs)
        self.partition = part = partition or TilePartition(view_bbs, params, owners, osp, cs, ws)
        lviews = [local_views[i] for i in mine]
        lparams = [params[i] for i in mine]
        if out_dtype is None:
            t = getattr(lviews[0], "tensor", None) if lviews else None
            if t is not None:
                from .fusion import _torch_to_np

                out_dtype = _torch_to_np(t.dtype)
            else:
                out_dtype = np.asarray(lviews[0]["data"]).dtype if lviews else np.float32
        self.np_dtype = np.dtype(out_dtype)
        self.out_start, slab_shape = part.slab[rank]
        self.out = engine.zeros(int(np.prod(slab_shape)), self.np_dtype).reshape(slab_shape)
        self.direct = engine.direct_plan(lviews, lparams, osp, cs, part.direct[rank], self.out, self.out_start)
        self.own = own = part.own_entries(rank)
        self.launches = 0
        if mode == "halo":
            self._init_halo(engine, part, local_views, mine, view_bbs, params, osp, cs)
            return
        # packed (num | den) windows: my own border boxes first, then one send buffer per owner
        self.own_off, n = [], 0
        for e in own:
            self.own_off.append(n)
            n += 2 * e["nvox"]
        self.acc = engine.zeros(n)
        self.send = {}
        boxes = [(e["start"], e["shape"]) for e in own]
        targets = [(self.acc, o, o + e["nvox"]) for e, o in zip(own, self.own_off)]
        for dst in range(ws):
            es = part.send_entries(rank, dst) if dst != rank else []
            if not es:
                continue
            m = sum(2 * e["nvox"] for e in es)
            buf = engine.zeros(m)
            self.send[dst] = buf
            o = 0
            for e in es:
                boxes.append((e["start"], e["shape"]))
                targets.append((buf, o, o + e["nvox"]))
                o += 2 * e["nvox"]
        self.partial = engine.partial_plan(lviews, lparams, osp, cs, boxes, targets)
        self.recv = {}
        for src in range(ws):
            es = part.send_entries(src, rank) if src != rank else []
            if es:
                self.recv[src] = (engine.zeros(sum(2 * e["nvox"] for e in es)), es)
        # owner-side adds: runs of consecutive own boxes per source -> one flat add each
        own_index = {e["chunk"]: k for k, e in enumerate(own)}
        self.adds = []  # (src, acc offset, recv offset, length)
        for src in sorted(self.recv):
            es = self.recv[src][1]
            k, pos = 0, 0
            while k < len(es):
                j = k
                first = own_index[es[k]["chunk"]]
                while j + 1 < len(es) and own_index[es[j + 1]["chunk"]] == first + (j + 1 - k):
                    j += 1
                length = sum(2 * e["nvox"] for e in es[k : j + 1])
                self.adds.append((src, self.own_off[first], pos, length))
                pos += length
                k = j + 1
        self.final_items = [(o, o + e["nvox"], e["start"], e["shape"]) for e, o in zip(own, self.own_off)]
        self.sent_bytes = int(sum(4 * b.numel() for b in self.send.values()))
        self.recv_bytes = int(sum(4 * b.numel() for b, _ in self.recv.values()))
        self.launches = 0

    # -- halo mode: raw windows of the foreign tiles travel, the owner fuses its border
    #    boxes from local views + received windows with the ordinary fused kernel ------
    def _init_halo(self, engine, part, local_views, mine, view_bbs, params, osp, cs):
        import torch

        rank, ws = self.rank, self.ws
        self.partial = _Runner()
        self.adds, self.final_items = [], []

        def layout(items, itemsize):
            offs, n = [], 0
            for vi, lo, hi in items:
                offs.append(n)
                n += -(-int(np.prod(hi - lo + 1)) * itemsize // 16) * 16
            return offs, n

        self.send, self.packs = {}, []  # packs: (buffer window tensor, source window tensor)
        for dst in range(ws):
            items = part.halo_sends(rank, dst) if dst != rank else []
            if not items:
                continue
            t0 = engine.view_tensor(local_views[items[0][0]])
            offs, n = layout(items, t0.element_size())
            buf = engine.zeros(n, np.uint8)
            self.send[dst] = buf
            for (vi, lo, hi), o in zip(items, offs):
                t = engine.view_tensor(local_views[vi])
                shape = tuple(int(v) for v in hi - lo + 1)
                nb = int(np.prod(shape)) * t.element_size()
                dstw = buf[o : o + nb].view(t.dtype).view(shape)
                self.packs.append((dstw, t[tuple(slice(int(a), int(b) + 1) for a, b in zip(lo, hi))]))
        self.recv = {}
        border_views = [(i, local_views[i]) for i in mine]
        dims = geometry.spatial_dims(len(osp["shape"]))
        tdt = engine.view_tensor(local_views[mine[0]]).dtype if mine else torch.uint8
        es = torch.empty(0, dtype=tdt).element_size()
        for src in range(ws):
            items = part.halo_sends(src, rank) if src != rank else []
            if not items:
                continue
            offs, n = layout(items, es)
            buf = engine.zeros(n, np.uint8)
            self.recv[src] = (buf, items)
            for (vi, lo, hi), o in zip(items, offs):
                shape = tuple(int(v) for v in hi - lo + 1)
                nb = int(np.prod(shape)) * es
                win = buf[o : o + nb].view(tdt).view(shape)
                bb = view_bbs[vi]
                origin = {d: bb["origin"][d] + float(lo[k]) * bb["spacing"][d] for k, d in enumerate(dims)}
                border_views.append((vi, engine.make_view(win, origin, bb["spacing"])))
        border_views.sort(key=lambda kv: kv[0])  # global view order = the one-GPU summation order
        boxes = [(e["start"], e["shape"]) for e in self.own]
        self.border = engine.border_plan([v for _, v in border_views], [params[i] for i, _ in border_views],
                                         [view_bbs[i] for i, _ in border_views], osp, cs, boxes, self.out, self.out_start)
        self.sent_bytes = int(sum(b.numel() for b in self.send.values()))
        self.recv_bytes = int(sum(b.numel() for b, _ in self.recv.values()))

    def _run_halo(self):
        import torch.distributed as dist

        for dstw, src in self.packs:
            dstw.copy_(src)
        works = []
        if self.ws > 1 and (self.send or self.recv):
            ops = [dist.P2POp(dist.isend, self.send[dst], dst) for dst in sorted(self.send)]
            ops += [dist.P2POp(dist.irecv, self.recv[src][0], src) for src in sorted(self.recv)]
            works = dist.batch_isend_irecv(ops)
        self.direct.run()  # overlaps the exchange
        for w in works:
            w.wait()
        self.border.run()
        self.launches = len(self.packs) + self.direct.launches + self.border.launches
        return self.out

    def run(self):
        import torch.distributed as dist

        if self.mode == "halo":
            return self._run_halo()
        n = self.direct.launches + self.partial.launches
        self.partial.run()
        works = []
        if self.ws > 1 and (self.send or self.recv):
            ops = [dist.P2POp(dist.isend, self.send[dst], dst) for dst in sorted(self.send)]
            ops += [dist.P2POp(dist.irecv, self.recv[src][0], src) for src in sorted(self.recv)]
            works = dist.batch_isend_irecv(ops)
        self.direct.run()  # overlaps the exchange
        for w in works:
            w.wait()
        for src, a_off, r_off, length in self.adds:
            self.acc[a_off : a_off + length] += self.recv[src][0][r_off : r_off + length]
        n += len(self.adds)
        n += self.engine.finalize(self.acc, self.final_items, self.out, self.out_start, self.np_dtype)
        self.launches = n
        return self.out

    def info(self):
        return {"partition": self.partition, "mode": self.mode, "sent_bytes": self.sent_bytes,
                "recv_bytes": self.recv_bytes, "border_boxes": len(self.own), "launches": int(self.launches)}

    def close(self):
        self.direct.close()
        self.partial.close()
        if self.mode == "halo":
            self.border.close()


def fuse_tile_partitioned(local_views, view_bbs, params, owners, output_stack_properties, output_chunksize=None,
                          out_dtype=None, engine=None, partition=None, mode="halo", **plan_kwargs):
    """One-shot ``TilePartitionedFuser``.  Returns ``(out, out_start, info)``: this rank's
    slab of the fused stack (covering the chunks it owns), the stack index of its first
    voxel and ``info`` = {"partition", "sent_bytes", "recv_bytes", "border_boxes", "launches"}."""
    f = TilePartitionedFuser(local_views, view_bbs, params, owners, output_stack_properties, output_chunksize,
                             out_dtype, engine, partition, mode, **plan_kwargs)
    try:
        out = f.run()
        return out, f.out_start, f.info()
    finally:
        f.close()


def fuse_partial(local_views, local_params, output_stack_properties, output_chunksize=None, fusion_func=None,
                 out_dtype=None, **plan_kwargs):
    """Whole-volume variant of tile-partitioned fusion for stacks that fit every GPU:
    every rank contributes partial weighted sums for the views IT holds over the WHOLE
    stack; one in-place all-reduce (NCCL over NVLink) sums them; the divide, NaN->0 and
    cast run locally, so every rank ends up with the full fused stack (CUDA).  Prefer
    ``fuse_tile_partitioned`` (border boxes only, each rank keeps its slab).  Single-view
    voxels come out as (v*b)/b, i.e. within 1 LSB of the one-GPU result after the
    truncating cast."""
    import torch
    import torch.distributed as dist

    from .fusion import FusionPlan, _fusion_mode, _np_to_torch, _torch_to_np, to_device_view

    rank, ws = world()
    mode = _fusion_mode(fusion_func)
    if mode == _lib.MVS_FUSE_MEAN:
        raise EngineError("partial fusion supports weighted_average_fusion and max_fusion")
    dviews = [to_device_view(v) for v in local_views]
    lib = _lib.load(require_device=True)
    ndim = len(output_stack_properties["shape"])
    dims = geometry.spatial_dims(ndim)
    full_shape = tuple(int(output_stack_properties["shape"][d]) for d in dims)
    np_dtype = np.dtype(out_dtype or (_torch_to_np(dviews[0].tensor.dtype) if dviews else np.float32))
    if mode == _lib.MVS_FUSE_MAX:
        # max of the per-rank maxima.  A voxel no local view covers is 0 in the per-rank result
        # (like the reference's nan_to_num of an all-NaN nanmax), which is neutral under MAX
        # for non-negative data -- the only kind this path accepts (validity is decided by the
        # coordinate predicate inside the kernel, not by a blending weight).
        for v in dviews:
            if v.tensor.dtype == torch.float32 and bool((torch.nan_to_num(v.tensor) < 0).any()):
                raise EngineError("fuse_partial(max_fusion) needs non-negative data (uncovered voxels are 0)")
        if dviews:
            plan = FusionPlan(dviews, local_params, output_stack_properties, output_chunksize=output_chunksize,
                              fusion_func=fusion_func, out_dtype=np.float32, **plan_kwargs)
            part = plan.run()
            plan.close()
        else:
            part = torch.zeros(full_shape, dtype=torch.float32, device="cuda")
        if ws > 1:
            dist.all_reduce(part, op=dist.ReduceOp.MAX)
        return part.to(_np_to_torch(np_dtype)) if np_dtype != np.float32 else part
    if dviews:
        plan = FusionPlan(dviews, local_params, output_stack_properties, output_chunksize=output_chunksize,
                          partial=True, **plan_kwargs)
        plan.run()
        both = plan.acc
        plan.close()
    else:
        both = torch.zeros((2,) + full_shape, dtype=torch.float32, device="cuda")
    if ws > 1:
        # whole-volume variant: one in-place sum of the (num, den) buffer
        dist.all_reduce(both, op=dist.ReduceOp.SUM)
    num, den = both[0], both[1]
    out = torch.empty(full_shape, dtype=_np_to_torch(np_dtype), device="cuda")
    _lib.check(
        lib.mvs_fuse_finalize(ctypes.c_void_p(num.data_ptr()), ctypes.c_void_p(den.data_ptr()),
                              ctypes.c_void_p(out.data_ptr()), _lib.mvs_dtype(np_dtype), num.numel(),
                              _lib.current_stream_ptr()),
        "mvs_fuse_finalize",
    )
    return out
