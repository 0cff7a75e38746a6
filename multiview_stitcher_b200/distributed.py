"""Multi-GPU sharding of the hot path (one process per GPU, torch.distributed).

Both stages shard into independent units:

* registration -- one unit per overlap pair (registration.py:2657-2664); pairs
  are dealt round-robin, results gathered as small host objects.  No data-path
  collective.
* fusion -- one unit per output chunk (fusion/_core.py:1133-1141).
  - ``fuse_sharded``: contiguous slabs of the chunk grid per rank, every rank
    holding (replicas of) the tiles its slab touches -> no communication.
  - ``fuse_partial``: the tiles themselves are partitioned; every rank produces
    un-normalised partial sums (sum_i v_i*b_i, sum_i b_i) for its tiles and the
    partials are summed with ONE NCCL all-reduce over NVLink before the divide.
    Valid because normalisation is linear:
    sum_i v_i b_i / sum_i b_i (the per-view normalisers of weights.py:340-345
    cancel).  ``max_fusion`` reduces with MAX instead.
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


def fuse_sharded(views, params, output_stack_properties, output_chunksize=None, gather=False, **plan_kwargs):
    """Fuses this rank's slab of output chunks.  Returns ``(out, owned_chunks)``:
    ``out`` is the full-size output tensor with only the owned chunks written
    (``gather=True`` sums the disjoint slabs so every rank holds the whole stack)."""
    import torch
    import torch.distributed as dist

    from .fusion import FusionPlan, to_device_view

    rank, ws = world()
    dviews = [to_device_view(v) for v in views]
    ndim = dviews[0].ndim
    dims = geometry.spatial_dims(ndim)
    if output_chunksize is None:
        output_chunksize = geometry.DEFAULT_CHUNKSIZE_2D if ndim == 2 else geometry.DEFAULT_CHUNKSIZE_3D
    cs = {d: int(output_chunksize[d]) for d in dims}
    n_chunks = len(geometry.chunk_grid(output_stack_properties, cs))
    owned = shard_slabs(n_chunks, rank, ws)
    plan = FusionPlan(dviews, params, output_stack_properties, output_chunksize=cs, chunk_subset=owned, **plan_kwargs)
    out = plan.run()
    plan.close()
    if gather and ws > 1:
        # slabs are disjoint and zero elsewhere: a sum assembles the stack
        buf = out.to(torch.float32) if out.dtype != torch.float32 else out
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
        out = buf.to(out.dtype) if buf is not out else out
    return out, owned


def fuse_partial(local_views, local_params, output_stack_properties, output_chunksize=None, fusion_func=None,
                 out_dtype=None, **plan_kwargs):
    """Tile-partitioned fusion: every rank contributes partial weighted sums for
    the views IT holds; one all-reduce (NCCL over NVLink) sums them; the divide,
    NaN->0 and cast run locally on every rank.  Returns the fused stack (CUDA)."""
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
        num, den = plan.acc_num, plan.acc_den
        plan.close()
    else:
        num = torch.zeros(full_shape, dtype=torch.float32, device="cuda")
        den = torch.zeros_like(num)
    if ws > 1:
        # the only data-path collective of the engine: sum of two float32 volumes
        both = torch.stack([num, den])
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
