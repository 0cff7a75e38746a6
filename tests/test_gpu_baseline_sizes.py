"""Size-independent properties at BASELINE.json's full sizes (the oracle is too
slow there): tiles cut from ONE synthetic ground truth at integer positions must
fuse back to that ground truth, and every overlap pair must register to the
jitter difference the generator applied.

* C2: 5x5 grid of 2048x2048 float32 tiles, 15 % overlap -- fusion with
  cosine-edge blending and registration of all 40 pairs.
* C3-like: 2x2x2 grid of (256, 512, 512) uint16 tiles, 10 % overlap -- 3-D fusion
  and registration of the x-neighbour pairs.
"""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _grid(grid, tile, ov, dtype):
    from multiview_stitcher_b200 import geometry, synthetic

    views, stage, true = synthetic.make_grid(grid, tile, ov, dtype, jitter=2, seed=0)
    osp = geometry.union_stack_props([v.bb() for v in views], true, views[0].spacing)
    return views, true, osp


def _ground_truth(osp, dims, dtype):
    """The generator evaluated directly on the fused stack's grid."""
    from multiview_stitcher_b200 import synthetic

    shape = tuple(int(osp["shape"][d]) for d in dims)
    origin = tuple(int(round(osp["origin"][d])) for d in dims)
    return synthetic.make_tile(shape, origin, dtype, seed=0)


def test_c2_fusion_reproduces_ground_truth():
    import torch

    from multiview_stitcher_b200 import fusion

    views, true, osp = _grid((5, 5), (2048, 2048), (307, 307), np.float32)
    # tile origins (stage + true translation) are integers: resampling is exact
    plan = fusion.FusionPlan(views, true, osp)
    plan.run()
    torch.cuda.synchronize()
    fused = plan.out
    gt = _ground_truth(osp, "yx", np.float32)
    assert fused.shape == gt.shape
    covered = fused != 0
    err = (fused - gt).abs()
    # identical values in the overlaps: the weighted average returns them up to float32 rounding
    tol = 1e-4 * gt.abs() + 1e-6 * float(gt.abs().max())
    assert bool(((err <= tol) | ~covered).all()), float(err[covered].max())
    # single-view voxels are bit-exact: most of the stack
    assert float((err[covered] == 0).float().mean()) > 0.7
    # the union of the tiles covers the stack except the jitter fringe
    assert float(covered.float().mean()) > 0.99
    plan.close()


def test_c2_all_pairs_register_to_the_jitter():
    import bench

    from multiview_stitcher_b200 import registration

    views, true, _ = _grid(bench.GRID, bench.TILE, bench.OVERLAP, np.float32)
    pairs = bench.c2_pairs()
    fixed, moving = bench.pair_crops([v.tensor for v in views], pairs)
    res = registration.register_pairs([f.contiguous() for f in fixed], [m.contiguous() for m in moving])
    t = np.array([p[:2, 2] for p in true])
    for r, (a, b, _) in zip(res, pairs):
        np.testing.assert_allclose(r["affine_matrix"][:2, 2], -(t[b] - t[a]), atol=0.1)
        assert r["quality"] > 0.95


def test_c3_fusion_and_pairs():
    import torch

    from multiview_stitcher_b200 import fusion, registration

    grid, tile, ov = (2, 2, 2), (256, 512, 512), (26, 51, 51)
    views, true, osp = _grid(grid, tile, ov, np.uint16)
    plan = fusion.FusionPlan(views, true, osp)
    plan.run()
    torch.cuda.synchronize()
    fused = plan.out.to(torch.int32)
    gt = _ground_truth(osp, "zyx", np.uint16).to(torch.int32)
    covered = fused != 0
    d = (fused - gt).abs()
    assert int(d[covered].max()) <= 1  # <= 1 LSB on blended voxels, exact elsewhere
    assert float((d[covered] == 0).float().mean()) > 0.7
    plan.close()
    idx = list(np.ndindex(*grid))
    pairs = [(i, i + 1) for i, c in enumerate(idx) if c[2] + 1 < grid[2]]
    fixed = [views[a].tensor[:, :, -ov[2]:].to(torch.float32).contiguous() for a, b in pairs]
    moving = [views[b].tensor[:, :, : ov[2]].to(torch.float32).contiguous() for a, b in pairs]
    res = registration.register_pairs(fixed, moving)
    t = np.array([p[:3, 3] for p in true])
    for r, (a, b) in zip(res, pairs):
        np.testing.assert_allclose(r["affine_matrix"][:3, 3], -(t[b] - t[a]), atol=0.1)


def test_c2_all_pairs_from_tiles_and_register_then_fuse():
    """The whole path at C2's size without pre-cut crops: ``pairs.register_views`` on the
    resident tiles at their STAGE positions (overlap boxes, crop windows, resampling, 40
    registrations, physical transforms) recovers every pair's jitter difference; fusing
    with transforms chained from those pairwise results reproduces the ground truth."""
    import torch

    import bench

    from multiview_stitcher_b200 import fusion, geometry, pairs as pairs_mod, synthetic

    views, stage, true = synthetic.make_grid(bench.GRID, bench.TILE, bench.OVERLAP, np.float32, jitter=2, seed=0)
    edges = [(a, b) for a, b, _ in bench.c2_pairs()]
    res, prep = pairs_mod.register_views(views, stage, edges, registration_binning={"y": 1, "x": 1}, return_prepared=True)
    assert prep.launches == 2 and sorted({tuple(f.shape) for f in prep.fixed}) == [(307, 2048), (2048, 307)]
    t = np.array([p[:2, 2] for p in true])
    for r, (a, b) in zip(res, edges):
        np.testing.assert_allclose(r["transform"][:2, 2], -(t[b] - t[a]), atol=0.1)
        assert r["quality"] > 0.95
    # chain the pairwise transforms along a spanning tree rooted at tile 0 (first row, then
    # down each column): view-to-world translation of tile b = that of a minus the pair's
    # shift (transform maps fixed world -> moving world)
    nx = bench.GRID[1]
    shift = {e: r["transform"][:2, 2] for e, r in zip(edges, res)}
    pos = {0: np.zeros(2)}
    for k in range(1, len(views)):
        a = k - 1 if k < nx else k - nx
        pos[k] = pos[a] - shift[(a, k)]
    params = []
    for k in range(len(views)):
        p = np.eye(3)
        p[:2, 2] = np.round(pos[k] - pos[0] + t[0])  # anchor on tile 0's true offset
        params.append(p)
    for p, q in zip(params, true):
        np.testing.assert_array_equal(p, q)
    osp = geometry.union_stack_props([v.bb() for v in views], params, views[0].spacing)
    plan = fusion.FusionPlan(views, params, osp)
    plan.run()
    torch.cuda.synchronize()
    gt = _ground_truth(osp, "yx", np.float32)
    covered = plan.out != 0
    err = (plan.out - gt).abs()
    tol = 1e-4 * gt.abs() + 1e-6 * float(gt.abs().max())
    assert bool(((err <= tol) | ~covered).all())
    plan.close()
