"""torchrun worker: sharded registration, slab-sharded fusion and tile-partitioned
(partial-sum) fusion over NCCL must reproduce the single-GPU results."""

import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from multiview_stitcher_b200 import distributed, fusion, geometry, registration, synthetic  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ws = dist.get_world_size()
    views, stage, true = synthetic.make_grid((2, 4), (192, 256), (40, 48), np.float32, jitter=2, seed=7, subpixel=True)
    osp = geometry.union_stack_props([v.bb() for v in views], true, views[0].spacing)
    ref, _ = fusion.fuse(views, true, output_stack_properties=osp, output_on_backend=True)

    # tiles partitioned by grid column blocks: border boxes only cross NVLink
    owners = [(i % 4) * ws // 4 for i in range(len(views))]
    bbs = [v.bb() for v in views]
    local = {i: views[i] for i in range(len(views)) if owners[i] == rank}
    for mode in ("halo", "partial"):
        slab, start, info = distributed.fuse_tile_partitioned(local, bbs, true, owners, osp, {"y": 96, "x": 128}, mode=mode)
        part = info["partition"]
        got = torch.zeros_like(ref)
        got[tuple(slice(a, a + n) for a, n in zip(start, slab.shape))] = slab
        mask = torch.zeros_like(ref, dtype=torch.bool)
        for ci, (cs_, cn_) in enumerate(part.grid):
            if part.owner_of[ci] == rank:
                mask[tuple(slice(a, a + n) for a, n in zip(cs_, cn_))] = True
        err = ((got - ref).abs() * mask).max().item()
        # halo: same kernel, same views in the same order -> float32 rounding of the box origin at most
        tol = (1e-6 if mode == "halo" else 1e-4) * ref.abs().max().item()
        assert err <= tol, f"fuse_tile_partitioned[{mode}] mismatch {err}"
        assert ws == 1 or info["sent_bytes"] + info["recv_bytes"] > 0
    assert part.halo_bytes(4) < part.exchanged_bytes(), "raw windows should be smaller than partial sums"
    cover = mask.to(torch.int32)
    dist.all_reduce(cover)
    assert int(cover.min()) == 1 and int(cover.max()) == 1, "chunk ownership does not tile the stack"
    assert ws == 1 or part.exchanged_bytes() > 0
    assert part.exchanged_bytes() < 0.5 * 8 * ref.numel(), "border exchange should be a fraction of the stack"

    # slab-sharded chunks, no data-path communication until the optional gather
    out, owned = distributed.fuse_sharded(views, true, osp, output_chunksize={"y": 96, "x": 128}, gather=True)
    assert torch.allclose(out, ref, rtol=1e-5, atol=1e-6), "fuse_sharded mismatch"

    # whole-volume variant: this rank only holds views rank, rank+ws, ...
    mine = list(range(rank, len(views), ws))
    part = distributed.fuse_partial([views[i] for i in mine], [true[i] for i in mine], osp)
    assert torch.allclose(part, ref, rtol=1e-4, atol=1e-5), "fuse_partial mismatch"

    # sharded registration of the horizontal neighbour pairs
    fixed = [views[i].tensor[:, -48:].contiguous() for i in (0, 1, 2, 4, 5, 6)]
    moving = [views[i + 1].tensor[:, :48].contiguous() for i in (0, 1, 2, 4, 5, 6)]
    res = distributed.register_pairs_sharded(fixed, moving)
    solo = registration.register_pairs(fixed, moving)
    for a, b in zip(res, solo):
        assert np.array_equal(a["affine_matrix"], b["affine_matrix"]) and a["quality"] == b["quality"]
    dist.barrier()
    if rank == 0:
        print("DIST_OK", ws)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
