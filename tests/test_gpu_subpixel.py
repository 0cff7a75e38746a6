"""Sub-pixel inputs at BASELINE sizes: tiles of the band-limited analytic field at
FRACTIONAL positions (jitter ~ U(-2, 2) px), engine vs ORACLE.

* one 2048 x 2048 output chunk of C2 (5x5 grid of 2048^2 float32 tiles) fed by 4-9
  views with fractional offsets -> fused float32 within 1e-4 relative of the oracle's
  fuse_np (every interpolation fraction non-zero: the dy / dz lerps of the stencil run);
* C2's crop shapes (2048 x 307, 307 x 2048): all stages of the registration vs the
  oracle -- shifts within 0.1 px, same winner, quality to 1e-6;
* a 3-D face pair at C3's crop shape (256 x 512 x 51): both normalisations' shift
  candidates vs the oracle (the oracle's candidate loop over 128 candidates of 6.7 Mvoxel
  is run at 64 x 256 x 51 instead, where it finishes in seconds).

Reference test this mirrors: _tests/test_registration.py:262-336 (fractional ground truth
0.488, 2.152; tolerance 0.1 px).
"""

import numpy as np
import pytest

from oracle import fusion as of
from oracle import registration as oreg

pytestmark = pytest.mark.gpu

TOL_PX = 0.1 + 1e-4  # north_star: recovered shifts within 0.1 px of the reference path


def _host(v):
    return {"data": v.tensor.cpu().numpy(), "origin": v.origin, "spacing": v.spacing}


def test_c2_chunk_with_fractional_views_matches_oracle():
    from multiview_stitcher_b200 import fusion, geometry, synthetic

    views, stage, true = synthetic.make_grid((5, 5), (2048, 2048), (307, 307), np.float32, jitter=2, seed=0, subpixel=True)
    offs = np.array([p[:2, 2] for p in true])
    assert np.all(offs != np.round(offs)), "every tile must sit at a fractional position"
    bbs = [v.bb() for v in views]
    osp = geometry.union_stack_props(bbs, true, bbs[0]["spacing"])
    plan = fusion.FusionPlan(views, true, osp)
    fused = plan.run()
    # the chunk at grid position (1, 1): rows / cols 2048..4095 -> touches 9 tiles
    cs = 2048
    grid = geometry.chunk_grid(osp, {"y": cs, "x": cs})
    ci = [i for i, (s, n) in enumerate(grid) if s == (cs, cs)][0]
    start, shape = grid[ci]
    first, count = plan.work["chunks"][ci][2], plan.work["chunks"][ci][3]
    sel = plan.work["view_index"][first : first + count]
    assert 4 <= len(sel) <= 9
    cprops = {"origin": {d: osp["origin"][d] + a * osp["spacing"][d] for d, a in zip("yx", start)},
              "spacing": osp["spacing"], "shape": {"y": shape[0], "x": shape[1]}}
    ref = of.fuse_np([_host(views[i]) for i in sel], [true[i] for i in sel], cprops, full_view_bbs=[bbs[i] for i in sel])
    got = fused[start[0] : start[0] + shape[0], start[1] : start[1] + shape[1]].cpu().numpy()
    tol = 1e-4 * np.abs(ref) + 1e-6 * np.abs(ref).max()
    bad = np.abs(got - ref) > tol
    assert not bad.any(), (int(bad.sum()), float(np.abs(got - ref).max()))
    # blended voxels exist and differ from any single view: the weights really ran
    assert float((ref != 0).mean()) > 0.99
    plan.close()


@pytest.mark.parametrize("shape,axis", [((2048, 307), 1), ((307, 2048), 0)])
def test_c2_crop_shapes_fractional_shift_matches_oracle(shape, axis):
    from multiview_stitcher_b200 import registration, synthetic

    rng = np.random.default_rng(7 + axis)
    fixed, moving, truth = [], [], []
    for k in range(2):
        ja = np.round(rng.uniform(-2, 2, 2) * 64) / 64
        jb = np.round(rng.uniform(-2, 2, 2) * 64) / 64
        fixed.append(synthetic.make_tile_field(shape, (3000 + ja[0], 1700 + ja[1]), np.float32, seed=2, tile_id=2 * k))
        moving.append(synthetic.make_tile_field(shape, (3000 + jb[0], 1700 + jb[1]), np.float32, seed=2, tile_id=2 * k + 1))
        truth.append(ja - jb)
    res = registration.register_pairs(fixed, moving, return_details=True)
    for f, m, r, t in zip(fixed, moving, res, truth):
        ref = oreg.phase_correlation_registration(f.cpu().numpy(), m.cpu().numpy(), return_details=True)
        for a, b in zip(r["shift_candidates"], ref["shift_candidates"]):
            assert np.abs(np.asarray(a) - np.asarray(b)).max() <= TOL_PX
        assert np.abs(r["affine_matrix"] - ref["affine_matrix"]).max() <= TOL_PX, (r["affine_matrix"][:2, 2], ref["affine_matrix"][:2, 2])
        if np.array_equal(r["affine_matrix"], ref["affine_matrix"]):
            assert abs(r["quality"] - ref["quality"]) < 1e-6
        # and the algorithm itself lands near the generator's fractional truth
        assert np.abs(ref["affine_matrix"][:2, 2] - t).max() <= 0.2, (ref["affine_matrix"][:2, 2], t)
        assert np.abs(r["affine_matrix"][:2, 2] - t).max() <= 0.2
        assert r["quality"] > 0.9


def _pair3d(shape, seed):
    from multiview_stitcher_b200 import synthetic

    rng = np.random.default_rng(seed)
    ja = np.round(rng.uniform(-2, 2, 3) * 64) / 64
    jb = np.round(rng.uniform(-2, 2, 3) * 64) / 64
    f = synthetic.make_tile_field(shape, tuple(100 + ja), np.float32, seed=seed, tile_id=0)
    m = synthetic.make_tile_field(shape, tuple(100 + jb), np.float32, seed=seed, tile_id=1)
    return f, m, ja - jb


def test_3d_pair_fractional_shift_full_loop_matches_oracle():
    from multiview_stitcher_b200 import registration

    f, m, t = _pair3d((64, 256, 51), 5)
    r = registration.register_pairs([f], [m], return_details=True)[0]
    ref = oreg.phase_correlation_registration(f.cpu().numpy(), m.cpu().numpy(), return_details=True)
    # upsample_factor 2: the sub-pixel grid is 0.5 px; equal bins or adjacent-bin ties only
    for a, b in zip(r["shift_candidates"], ref["shift_candidates"]):
        d = np.abs(np.asarray(a) - np.asarray(b))
        assert np.all((d <= 1e-6) | (np.abs(d - 0.5) <= 1e-6)), (a, b)
    d = np.abs(r["affine_matrix"][:3, 3] - ref["affine_matrix"][:3, 3])
    assert np.all(d <= TOL_PX), (r["affine_matrix"][:3, 3], ref["affine_matrix"][:3, 3])
    assert abs(r["quality"] - ref["quality"]) < 1e-6
    assert np.abs(ref["affine_matrix"][:3, 3] - t).max() <= 0.5


def test_c3_crop_shape_shift_candidates_match_oracle():
    """256 x 512 x 51 (C3's x-face crop): FFT -> cross power -> IFFT -> peak -> upsampled
    DFT against the oracle's scipy.fft path, both normalisations."""
    from multiview_stitcher_b200 import registration

    f, m, t = _pair3d((256, 512, 51), 9)
    r = registration.register_pairs([f], [m], return_details=True)[0]
    fh, mh = f.cpu().numpy(), m.cpu().numpy()
    lo, hi = fh.min(), fh.max()
    f01 = (fh - lo) / (hi - lo)
    lo, hi = mh.min(), mh.max()
    m01 = (mh - lo) / (hi - lo)
    nm = np.zeros(fh.shape, bool)
    ref = oreg.shift_candidates(f01, m01, f01, m01, nm, nm, 2)
    for a, b in zip(r["shift_candidates"], ref):
        assert np.abs(np.asarray(a) - np.asarray(b)).max() <= TOL_PX, (a, b)
    assert np.abs(r["affine_matrix"][:3, 3] - t).max() <= 0.5
