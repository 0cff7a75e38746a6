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


# --- tile-partitioned fusion: the plan, and the exchange protocol over gloo ----


def _grid_views(rng, grid=(2, 3), tile=(40, 48), ov=(10, 12)):
    """Small 2-D tile grid with fractional jitter, host views + bounding boxes."""
    views, params = [], []
    for iy in range(grid[0]):
        for ix in range(grid[1]):
            data = rng.random(tile).astype(np.float32) * 100
            views.append({"data": data, "origin": {"y": float(iy * (tile[0] - ov[0])), "x": float(ix * (tile[1] - ov[1]))},
                          "spacing": {"y": 1.0, "x": 1.0}})
            p = np.eye(3)
            p[:2, 2] = np.round(rng.uniform(-2, 2, 2) * 64) / 64
            params.append(p)
    return views, params


def test_tile_partition_plan_covers_every_voxel_once():
    from oracle import fusion as of

    rng = np.random.default_rng(0)
    views, params = _grid_views(rng)
    bbs = [of.view_bb(v) for v in views]
    osp = of.calc_stack_properties(bbs, params, views[0]["spacing"])
    for ws, owners in ((2, [0, 0, 0, 1, 1, 1]), (3, [0, 1, 2, 0, 1, 2]), (2, [0, 1, 0, 1, 0, 1]), (4, [0, 1, 2, 3, 0, 1])):
        part = D.TilePartition(bbs, params, owners, osp, {"y": 32, "x": 32}, ws)
        cover = np.zeros(part.full_shape, dtype=np.int32)
        for r in range(ws):
            for start, shape in part.direct[r]:
                cover[tuple(slice(a, a + n) for a, n in zip(start, shape))] += 1
            lo, n = part.slab[r]
            for e in part.own_entries(r):
                cover[tuple(slice(a, a + m) for a, m in zip(e["start"], e["shape"]))] += 1
                assert all(a >= l and a + m <= l + k for a, m, l, k in zip(e["start"], e["shape"], lo, n))
                assert r not in e["contrib"] and e["contrib"]
        assert cover.min() == 1 and cover.max() == 1
        # a direct box may only be reached by tiles of its owner
        o_org = np.array([osp["origin"][d] for d in "yx"])
        for r in range(ws):
            for start, shape in part.direct[r]:
                lo = o_org + np.array(start)
                hi = lo + np.array(shape) - 1
                for vi, (bb, p) in enumerate(zip(bbs, params)):
                    if owners[vi] == r:
                        continue
                    alo, ahi = D.geometry.transformed_aabb(bb, p, ["y", "x"])
                    assert np.any(ahi < lo) or np.any(alo > hi), (ws, r, start, shape, vi)
        assert part.exchanged_bytes() == sum(8 * e["nvox"] * len(e["contrib"]) for e in part.entries)


class _OracleEngine:
    """CPU stand-in for the device half of fuse_tile_partitioned: partial sums and direct
    fusion computed with the oracle (numpy / scipy) on CPU tensors."""

    def zeros(self, n, np_dtype=np.float32):
        return torch.zeros(n, dtype={np.dtype(np.float32): torch.float32, np.dtype(np.uint16): torch.uint16, np.dtype(np.uint8): torch.uint8}[np.dtype(np_dtype)])

    @staticmethod
    def _props(osp, start, shape):
        dims = list("zyx")[-len(shape):]
        return {"origin": {d: osp["origin"][d] + a * osp["spacing"][d] for d, a in zip(dims, start)},
                "spacing": dict(osp["spacing"]), "shape": {d: int(n) for d, n in zip(dims, shape)}}

    class _Run:
        launches = 1

        def __init__(self, fn):
            self.run = fn

        def close(self):
            pass

    def direct_plan(self, views, params, osp, chunksize, boxes, out, out_start):
        return self._Run(lambda: self._direct(views, params, osp, boxes, out, out_start))

    def border_plan(self, views, params, full_bbs, osp, chunksize, boxes, out, out_start):
        return self._Run(lambda: self._direct(views, params, osp, boxes, out, out_start, full_bbs))

    def view_tensor(self, view):
        return torch.from_numpy(view["data"])

    def make_view(self, tensor, origin, spacing):
        return {"data": tensor.numpy(), "origin": dict(origin), "spacing": dict(spacing)}

    def partial_plan(self, views, params, osp, chunksize, boxes, targets):
        return self._Run(lambda: self._partial(views, params, osp, boxes, targets))

    def _direct(self, views, params, osp, boxes, out, out_start, full_bbs=None):
        from oracle import fusion as of

        bbs = full_bbs or [of.view_bb(v) for v in views]
        for start, shape in boxes:
            res = of.fuse_np(views, params, self._props(osp, start, shape), full_view_bbs=bbs)
            sl = tuple(slice(a - o, a - o + n) for a, o, n in zip(start, out_start, shape))
            out[sl] = torch.from_numpy(res.astype(np.float32))

    def _partial(self, views, params, osp, boxes, targets):
        from oracle import fusion as of

        for (start, shape), (buf, o_num, o_den) in zip(boxes, targets):
            props = self._props(osp, start, shape)
            num = np.zeros(shape, dtype=np.float32)
            den = np.zeros(shape, dtype=np.float32)
            for v, p in zip(views, params):
                t = of.transform_view({**v, "data": v["data"].astype(np.float32)}, np.linalg.inv(p), props, cval=np.nan)
                b = of.get_blending_weights(props, of.view_bb(v), p) * ~np.isnan(t)
                num += np.nan_to_num(t) * b
                den += b
            n = int(np.prod(shape))
            buf[o_num : o_num + n] = torch.from_numpy(num.reshape(-1))
            buf[o_den : o_den + n] = torch.from_numpy(den.reshape(-1))

    def finalize(self, buf, items, out, out_start, np_dtype):
        for o_num, o_den, start, shape in items:
            n = int(np.prod(shape))
            num, den = buf[o_num : o_num + n], buf[o_den : o_den + n]
            den = torch.where(den == 0, torch.ones_like(den), den)
            sl = tuple(slice(a - o, a - o + m) for a, o, m in zip(start, out_start, shape))
            out[sl] = (num / den).reshape(shape)
        return 1


def _tp_worker(rank, ws, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        from oracle import fusion as of

        rng = np.random.default_rng(5)
        views, params = _grid_views(rng)
        bbs = [of.view_bb(v) for v in views]
        osp = of.calc_stack_properties(bbs, params, views[0]["spacing"])
        owners = [0, 0, 1, 0, 1, 1]  # ragged split: corner chunks draw from both ranks
        local = {i: views[i] for i in range(len(views)) if owners[i] == rank}
        ref, _ = of.fuse(views, params)
        # halo mode: raw windows of the foreign tiles travel and the owner fuses its border boxes
        # from local views + received windows
        out_h, start_h, info_h = D.fuse_tile_partitioned(local, bbs, params, owners, osp, {"y": 32, "x": 32},
                                                         out_dtype=np.float32, engine=_OracleEngine(), mode="halo")
        part_h = info_h["partition"]
        got_h = np.zeros(ref.shape, dtype=np.float32)
        got_h[tuple(slice(a, a + n) for a, n in zip(start_h, out_h.shape))] = out_h.numpy()
        mask_h = np.zeros(ref.shape, dtype=bool)
        for ci, (cs_, cn_) in enumerate(part_h.grid):
            if part_h.owner_of[ci] == rank:
                mask_h[tuple(slice(a, a + n) for a, n in zip(cs_, cn_))] = True
        # (numpy's reduction order depends on how many views a box / chunk stacks: 1 ulp)
        halo_ok = bool(np.abs(got_h - ref)[mask_h].max() <= 1e-6 * np.abs(ref).max()) and info_h["sent_bytes"] > 0
        halo_ok = halo_ok and info_h["sent_bytes"] == sum(
            -(-int(np.prod(hi - lo + 1)) * 4 // 16) * 16 for (o, vi), (lo, hi) in part_h.windows.items() if owners[vi] == rank)
        out, start, info = D.fuse_tile_partitioned(local, bbs, params, owners, osp, {"y": 32, "x": 32},
                                                   out_dtype=np.float32, engine=_OracleEngine(), mode="partial")
        sl = tuple(slice(a, a + n) for a, n in zip(start, out.shape))
        part = info["partition"]
        # compare only the chunks this rank owns inside its slab's bounding box
        mask = np.zeros(ref.shape, dtype=bool)
        for ci, (cs_, cn_) in enumerate(part.grid):
            if part.owner_of[ci] == rank:
                mask[tuple(slice(a, a + n) for a, n in zip(cs_, cn_))] = True
        got = np.zeros(ref.shape, dtype=np.float32)
        got[sl] = out.numpy()
        err = np.abs(got - ref)[mask].max() if mask.any() else 0.0
        tol = 1e-4 * np.abs(ref).max()
        ok = halo_ok and bool(err <= tol) and info["sent_bytes"] == sum(8 * e["nvox"] for e in part.entries if rank in e["contrib"])
        ok = ok and info["recv_bytes"] == sum(8 * e["nvox"] * len(e["contrib"]) for e in part.own_entries(rank))
        q.put((rank, ok, float(err), int(mask.sum()), info["sent_bytes"]))
    finally:
        dist.destroy_process_group()


def test_tile_partitioned_fusion_world_size_2_matches_oracle():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_tp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    res = [q.get(timeout=10) for _ in range(2)]
    assert all(r[1] for r in res), res
    assert sum(r[3] for r in res) > 0 and sum(r[4] for r in res) > 0
