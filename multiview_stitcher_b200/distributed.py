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
    several GPUs only the box the foreign tiles can reach is exchanged: every
    contributing rank produces un-normalised partial sums (sum_i v_i*b_i,
    sum_i b_i) of ITS tiles over that box, sends them to the owner (NCCL
    send/recv over NVLink), and the owner adds them to its own partial sums,
    divides and casts.  Valid because normalisation is linear:
    sum_i v_i b_i / sum_i b_i (the per-view normalisers of weights.py:340-345
    cancel).  The rest of the chunk is fused directly.
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


def fuse_sharded(views, params, output_stack_properties, output_chunksize=None, gather=False, band_axis=None,
                 **plan_kwargs):
    """Fuses this rank's band of output chunks (tiles replicated where bands meet, no
    communication).  Returns ``(out, owned_chunks)``: ``out`` is the full-size output tensor
    with only the owned chunks written; its memory is laid out with the band axis slowest,
    so a band is one contiguous range.  ``gather=True`` broadcasts every rank's band in
    the OUTPUT dtype (each rank ends up holding the whole stack; bytes on the wire =
    stack bytes x (N-1)/N per rank, no float32 round trip)."""
    import torch
    import torch.distributed as dist

    from .fusion import FusionPlan, _np_to_torch, _torch_to_np, to_device_view

    rank, ws = world()
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
    bands = [shard_slabs(counts[band_axis], r, ws) for r in range(ws)]
    mine = set(bands[rank])
    owned = [i for i, (start, _) in enumerate(grid) if start[band_axis] // cs[dims[band_axis]] in mine]
    # memory order: band axis first
    perm = [band_axis] + [a for a in range(ndim) if a != band_axis]
    inv = [perm.index(a) for a in range(ndim)]
    np_dt = np.dtype(plan_kwargs.pop("out_dtype", None) or _torch_to_np(dviews[0].tensor.dtype))
    out = torch.zeros([full[a] for a in perm], dtype=_np_to_torch(np_dt), device="cuda").permute(inv)
    if owned:
        plan = FusionPlan(dviews, params, osp, output_chunksize=cs, chunk_subset=owned, out=out, out_dtype=np_dt,
                          **plan_kwargs)
        plan.run()
        plan.close()
    nbytes = 0
    if gather and ws > 1:
        mem = out.permute(perm)  # contiguous, band axis leading
        c = cs[dims[band_axis]]
        for r in range(ws):
            if not bands[r]:
                continue
            a, b = bands[r][0] * c, min((bands[r][-1] + 1) * c, full[band_axis])
            slab = mem[a:b].view(torch.uint8)
            dist.broadcast(slab, src=r)
            if r != rank:
                nbytes += slab.numel()
    fuse_sharded.last_gather_bytes = nbytes
    return out, owned


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
                            "shape": tuple(int(v) for v in ushape), "nvox": int(np.prod(ushape))})
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

    def own_entries(self, rank):
        return [e for e in self.entries if e["owner"] == rank]

    def send_entries(self, src, dst):
        """Entries rank ``src`` contributes partial sums to, owned by ``dst``."""
        return [e for e in self.entries if e["owner"] == dst and src in e["contrib"]]

    def exchanged_bytes(self):
        """float32 (num, den) bytes crossing NVLink per job, all ranks together."""
        return sum(8 * e["nvox"] * len(e["contrib"]) for e in self.entries)


class _CudaEngine:
    """Device half of ``fuse_tile_partitioned`` (the CPU test substitutes an oracle-backed
    engine to exercise the exchange protocol over gloo)."""

    def __init__(self, **plan_kwargs):
        self.kw = plan_kwargs

    def zeros(self, n, np_dtype=np.float32):
        import torch

        from .fusion import _np_to_torch

        return torch.zeros(n, dtype=_np_to_torch(np.dtype(np_dtype)), device="cuda")

    def fuse_direct(self, views, params, osp, chunksize, boxes, out, out_start):
        from .fusion import FusionPlan

        if not boxes:
            return 0
        plan = FusionPlan(views, params, osp, output_chunksize=chunksize, chunk_list=boxes, out=out,
                          out_start=out_start, **self.kw)
        plan.run()
        n = plan.launches_per_run
        plan.close()
        return n

    def fuse_partial(self, views, params, osp, chunksize, boxes, targets):
        """boxes[i] accumulated into the packed float32 (num, den) windows targets[i] =
        (buffer tensor, element offset of num, element offset of den)."""
        from .fusion import FusionPlan

        if not boxes:
            return 0
        tg = []
        for (start, shape), (buf, o_num, o_den) in zip(boxes, targets):
            strides = [int(np.prod(shape[i + 1:])) for i in range(len(shape))]
            tg.append((buf.data_ptr() + 4 * o_num, buf.data_ptr() + 4 * o_den, strides))
        kw = {k: v for k, v in self.kw.items() if k != "fusion_func"}
        plan = FusionPlan(views, params, osp, output_chunksize=chunksize, chunk_list=boxes, partial=True,
                          chunk_targets=tg, **kw)
        plan.run()
        n = plan.launches_per_run
        plan.close()
        return n

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


def fuse_tile_partitioned(local_views, view_bbs, params, owners, output_stack_properties, output_chunksize=None,
                          out_dtype=None, engine=None, partition=None, **plan_kwargs):
    """Fusion of a stack whose tiles are partitioned over the ranks.

    ``local_views``: {global view index: view} for the views THIS rank holds (exactly the
    indices ``i`` with ``owners[i] == rank``); ``view_bbs`` / ``params`` / ``owners``: bounding
    box, affine and owner rank of EVERY view of the job (metadata, identical on all ranks).
    Weighted-average fusion with blending weights only.

    Returns ``(out, out_start, info)``: this rank's slab of the fused stack (a CUDA tensor
    covering the chunks it owns), the stack index of its first voxel and ``info`` =
    {"partition", "sent_bytes", "recv_bytes", "border_boxes", "launches"}.
    """
    import torch
    import torch.distributed as dist

    rank, ws = world()
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
    if engine is None:
        engine = _CudaEngine(**plan_kwargs)
    part = partition or TilePartition(view_bbs, params, owners, osp, cs, ws)
    lviews = [local_views[i] for i in mine]
    lparams = [params[i] for i in mine]
    if out_dtype is None:
        t = getattr(lviews[0], "tensor", None) if lviews else None
        if t is not None:
            from .fusion import _torch_to_np

            out_dtype = _torch_to_np(t.dtype)
        else:
            out_dtype = np.asarray(lviews[0]["data"]).dtype if lviews else np.float32
    np_dt = np.dtype(out_dtype)
    slab_start, slab_shape = part.slab[rank]
    out = engine.zeros(int(np.prod(slab_shape)), np_dt).reshape(slab_shape)
    launches = 0
    if lviews:
        launches += engine.fuse_direct(lviews, lparams, osp, cs, part.direct[rank], out, slab_start)
    # packed (num | den) windows: my own border boxes first, then one send buffer per owner
    own = part.own_entries(rank)
    own_off, n = [], 0
    for e in own:
        own_off.append(n)
        n += 2 * e["nvox"]
    acc = engine.zeros(n)
    send, send_off = {}, {}
    for dst in range(ws):
        es = part.send_entries(rank, dst) if dst != rank else []
        if es:
            offs, m = [], 0
            for e in es:
                offs.append(m)
                m += 2 * e["nvox"]
            send[dst], send_off[dst] = engine.zeros(m), (es, offs)
    boxes, targets = [], []
    for e, o in zip(own, own_off):
        boxes.append((e["start"], e["shape"]))
        targets.append((acc, o, o + e["nvox"]))
    for dst, (es, offs) in send_off.items():
        for e, o in zip(es, offs):
            boxes.append((e["start"], e["shape"]))
            targets.append((send[dst], o, o + e["nvox"]))
    if lviews:
        launches += engine.fuse_partial(lviews, lparams, osp, cs, boxes, targets)
    # exchange: one message per (contributor -> owner) pair
    recv = {}
    for src in range(ws):
        if src == rank:
            continue
        es = part.send_entries(src, rank)
        if es:
            recv[src] = (engine.zeros(sum(2 * e["nvox"] for e in es)), es)
    sent_bytes = sum(4 * b.numel() for b in send.values())
    recv_bytes = sum(4 * b.numel() for b, _ in recv.values())
    if ws > 1 and (send or recv):
        ops = [dist.P2POp(dist.isend, send[dst], dst) for dst in sorted(send)]
        ops += [dist.P2POp(dist.irecv, recv[src][0], src) for src in sorted(recv)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    # owner: add the neighbours' partial sums (runs of consecutive boxes -> one add each)
    own_index = {e["chunk"]: k for k, e in enumerate(own)}
    for src in sorted(recv):
        buf, es = recv[src]
        k, pos = 0, 0
        while k < len(es):
            j = k
            first = own_index[es[k]["chunk"]]
            while j + 1 < len(es) and own_index[es[j + 1]["chunk"]] == first + (j + 1 - k):
                j += 1
            length = sum(2 * e["nvox"] for e in es[k : j + 1])
            acc[own_off[first] : own_off[first] + length] += buf[pos : pos + length]
            pos += length
            k = j + 1
    launches += engine.finalize(acc, [(o, o + e["nvox"], e["start"], e["shape"]) for e, o in zip(own, own_off)],
                                out, slab_start, np_dt)
    info = {"partition": part, "sent_bytes": int(sent_bytes), "recv_bytes": int(recv_bytes),
            "border_boxes": len(own), "launches": int(launches)}
    return out, slab_start, info


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
        # float32 max of the per-rank maxima; uncovered voxels carry -inf
        if dviews:
            plan = FusionPlan(dviews, local_params, output_stack_properties, output_chunksize=output_chunksize,
                              fusion_func=fusion_func, out_dtype=np.float32, **plan_kwargs)
            part = plan.run()
            cov = FusionPlan(dviews, local_params, output_stack_properties, output_chunksize=output_chunksize,
                             partial=True, **plan_kwargs)
            cov.run()
            part = torch.where(cov.acc_den > 0, part, torch.full_like(part, float("-inf")))
            plan.close(); cov.close()
        else:
            part = torch.full(full_shape, float("-inf"), dtype=torch.float32, device="cuda")
        if ws > 1:
            dist.all_reduce(part, op=dist.ReduceOp.MAX)
        part = torch.where(torch.isinf(part), torch.zeros_like(part), part)
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
