"""Host-side logic of the multi-GPU path on CPU: unit sharding and the result
gather, exercised with a world_size-2 gloo process group."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multiview_stitcher_b200 import distributed as D


def test_round_robin_and_slabs_partition_units():
    for n in (0, 1, 7, 40, 64):
        for ws in (1, 2, 3, 8):
            rr = [D.shard_round_robin(n, r, ws) for r in range(ws)]
            sl = [D.shard_slabs(n, r, ws) for r in range(ws)]
            for parts in (rr, sl):
                flat = sorted(i for p in parts for i in p)
                assert flat == list(range(n))
                sizes = [len(p) for p in parts]
                assert max(sizes) - min(sizes) <= 1
            for p in sl:  # slabs are contiguous
                assert p == list(range(p[0], p[0] + len(p))) if p else True


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, ws, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        n = 11
        owned = D.shard_round_robin(n, rank, ws)
        local = [{"affine_matrix": np.eye(3) * (i + 1), "quality": float(i)} for i in owned]
        full = D.gather_objects(local, n, owned)
        ok = all(full[i]["quality"] == float(i) and full[i]["affine_matrix"][0, 0] == i + 1 for i in range(n))
        # the partial-sum reduce of fuse_partial, on CPU tensors
        num = torch.full((4, 5), float(rank + 1))
        den = torch.full((4, 5), 0.5)
        both = torch.stack([num, den])
        dist.all_reduce(both, op=dist.ReduceOp.SUM)
        ok = ok and torch.all(both[0] == sum(range(1, ws + 1))).item() and torch.all(both[1] == 0.5 * ws).item()
        # register_views_sharded: pairs dealt round-robin, each rank registers only its share
        # (the device call is stubbed: this checks the dealing and the gather, not the kernels)
        from multiview_stitcher_b200 import pairs as pairs_mod

        seen = []

        def fake_register_views(views, affines, pairs, **kw):
            seen.extend(pairs)
            return [{"transform": np.eye(3) * (10 * a + b), "quality": float(a + b), "bbox": np.zeros((2, 2))} for a, b in pairs]

        real = pairs_mod.register_views
        pairs_mod.register_views = fake_register_views
        try:
            all_pairs = [(0, 1), (1, 2), (2, 3), (0, 4), (1, 5), (2, 6), (3, 7)]
            res = D.register_views_sharded([None] * 8, [np.eye(3)] * 8, all_pairs, registration_binning={"y": 1, "x": 1})
        finally:
            pairs_mod.register_views = real
        ok = ok and seen == all_pairs[rank::ws]
        ok = ok and all(r["transform"][0, 0] == 10 * a + b and r["quality"] == float(a + b) for r, (a, b) in zip(res, all_pairs))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_gather_and_reduce_world_size_2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    res = dict(q.get(timeout=10) for _ in range(2))
    assert res == {0: True, 1: True}
