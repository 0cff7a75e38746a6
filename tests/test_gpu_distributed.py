"""Multi-GPU fusion paths.  The partial-sum mathematics is checked on one GPU by
playing both ranks in-process; the NCCL run needs >= 2 GPUs (torchrun)."""

import os
import subprocess
import sys

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partial_sums_reduce_to_full_fusion():
    import ctypes

    import torch

    from multiview_stitcher_b200 import _lib, fusion, geometry, synthetic

    views, stage, true = synthetic.make_grid((2, 3), (160, 200), (36, 44), np.uint16, jitter=2, seed=4)
    true[1][:2, 2] += (0.3, -0.6)
    osp = geometry.union_stack_props([v.bb() for v in views], true, views[0].spacing)
    full, _ = fusion.fuse(views, true, output_stack_properties=osp, output_on_backend=True)
    groups = [[0, 2, 4], [1, 3, 5]]  # "rank 0" and "rank 1" tiles
    num = den = None
    for g in groups:
        plan = fusion.FusionPlan([views[i] for i in g], [true[i] for i in g], osp, partial=True)
        plan.run()
        num = plan.acc_num.clone() if num is None else num + plan.acc_num
        den = plan.acc_den.clone() if den is None else den + plan.acc_den
        plan.close()
    out = torch.empty_like(full)
    lib = _lib.load()
    _lib.check(lib.mvs_fuse_finalize(ctypes.c_void_p(num.data_ptr()), ctypes.c_void_p(den.data_ptr()),
                                     ctypes.c_void_p(out.data_ptr()), _lib.MVS_U16, num.numel(), _lib.current_stream_ptr()), "finalize")
    d = (out.cpu().numpy().astype(np.int64) - full.cpu().numpy().astype(np.int64))
    assert np.abs(d).max() <= 1
    assert (d != 0).mean() < 0.02


def test_tile_partition_boxes_on_one_gpu_match_full_fusion():
    """Plays every rank of a 3-rank tile partition in-process on one GPU: direct boxes +
    packed partial sums + box finalisation must rebuild the single-plan result."""
    import torch

    from multiview_stitcher_b200 import distributed as D, fusion, geometry, synthetic

    for dtype, grid, tile, ov, cs in ((np.uint16, (1, 2, 3), (24, 96, 112), (0, 20, 24), {"z": 16, "y": 64, "x": 64}),
                                      (np.float32, (2, 3), (160, 200), (36, 44), {"y": 96, "x": 128})):
        views, stage, true = synthetic.make_grid(grid, tile, ov, dtype, jitter=2, seed=4, subpixel=True)
        bbs = [v.bb() for v in views]
        osp = geometry.union_stack_props(bbs, true, views[0].spacing)
        full, _ = fusion.fuse(views, true, output_stack_properties=osp, output_chunksize=cs, output_on_backend=True)
        ws = 3
        owners = [i % ws for i in range(len(views))]
        part = D.TilePartition(bbs, true, owners, osp, cs, ws)
        eng = D._CudaEngine()
        out = torch.zeros_like(full)
        zero = [0] * full.ndim
        accs = {}
        for r in range(ws):
            lv = [views[i] for i in range(len(views)) if owners[i] == r]
            lp = [true[i] for i in range(len(views)) if owners[i] == r]
            r_ = eng.direct_plan(lv, lp, osp, cs, part.direct[r], out, zero)
            r_.run(); r_.close()
            # every entry r takes part in, accumulated into one buffer per entry
            es = [e for e in part.entries if e["owner"] == r or r in e["contrib"]]
            bufs = [torch.zeros(2 * e["nvox"], device="cuda") for e in es]
            r_ = eng.partial_plan(lv, lp, osp, cs, [(e["start"], e["shape"]) for e in es],
                                  [(b, 0, e["nvox"]) for b, e in zip(bufs, es)])
            r_.run(); r_.close()
            for e, b in zip(es, bufs):
                accs[e["chunk"]] = accs.get(e["chunk"], 0) + b
        for e in part.entries:
            eng.finalize(accs[e["chunk"]], [(0, e["nvox"], e["start"], e["shape"])], out, zero, np.dtype(dtype))
        a, b = out.cpu().numpy().astype(np.float64), full.cpu().numpy().astype(np.float64)
        if dtype == np.uint16:
            assert np.abs(a - b).max() <= 1
        else:
            assert np.all(np.abs(a - b) <= 1e-4 * np.abs(b) + 1e-6 * np.abs(b).max())
        assert part.entries and part.exchanged_bytes() < 8 * full.numel()


def test_fuse_partial_single_rank_equals_fuse():
    from multiview_stitcher_b200 import distributed, fusion, geometry

    case = cases.fusion_cases()["3d_f32_affine_lin"]
    dviews = [fusion.to_device_view(v) for v in case["views"]]
    osp = geometry.union_stack_props([v.bb() for v in dviews], case["params"], dviews[0].spacing)
    ref, _ = fusion.fuse(dviews, case["params"], output_stack_properties=osp)
    got = distributed.fuse_partial(dviews, case["params"], osp).cpu().numpy()
    tol = 1e-4 * np.abs(ref) + 1e-6 * np.abs(ref).max()
    assert np.all(np.abs(got - ref) <= tol)
    gmax = distributed.fuse_partial(dviews, case["params"], osp, fusion_func=fusion.max_fusion).cpu().numpy()
    rmax, _ = fusion.fuse(dviews, case["params"], output_stack_properties=osp, fusion_func=fusion.max_fusion)
    assert np.array_equal(gmax, rmax)


def test_two_gpu_nccl_paths():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29617", os.path.join(ROOT, "tests", "dist_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "DIST_OK" in r.stdout
