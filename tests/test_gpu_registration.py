"""Parity of the CUDA phase-correlation path (through the C ABI) with the
oracle / reference-generated fixtures.  Bar (BASELINE.json north_star):
recovered shifts within 0.1 px of the reference path; the sub-pixel grid is
1/upsample px, so adjacent-bin ties are the only admissible difference."""

import warnings

import numpy as np
import pytest
import scipy.fft as sfft
from scipy import ndimage, stats

import cases
from oracle import registration as oreg
from oracle import skimage_restated as sk

pytestmark = pytest.mark.gpu

TOL_PX = 0.1 + 1e-4


@pytest.fixture(scope="module")
def reg():
    from multiview_stitcher_b200 import registration

    return registration


def _rescaled(a):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return sk.rescale_intensity(a, in_range=(np.nanmin(a), np.nanmax(a)), out_range=(0, 1))


def _plan_for(reg, f, m, u=None):
    import torch

    u = u or (10 if f.ndim == 2 else 2)
    plan = reg.PhaseCorrPlan(f.shape, 1, u)
    stats_ = plan.load_pairs([torch.from_numpy(f).cuda()], [torch.from_numpy(m).cuda()])
    return plan, stats_


SHAPES = [(64, 128), (128, 77), (100, 100), (37, 307), (307, 2048), (2048, 307), (16, 32, 64), (24, 64, 51), (9, 20, 33), (40, 307, 64), (128, 256, 51)]


@pytest.mark.parametrize("shape", SHAPES)
def test_correlate_matches_numpy_fft(reg, shape):
    """Integer peaks and the upsampled-DFT samples equal a numpy evaluation of
    the same formulas (exercises Stockham + Bluestein on every axis)."""
    rng = np.random.default_rng(sum(shape))
    big = ndimage.gaussian_filter(rng.random(tuple(s + 12 for s in shape)), 1.0)
    sl = tuple(slice(6, 6 + s) for s in shape)
    shift = (2.3, -3.6, 1.2)[-len(shape):]
    f = big[sl].astype(np.float32)
    m = ndimage.shift(big, shift, order=3)[sl].astype(np.float32)
    plan, st = _plan_for(reg, f, m)
    peaks, updft = plan.correlate()
    r0, r1 = _rescaled(f), _rescaled(m)
    assert st[0, 0, 0] == f.min() and st[0, 1, 1] == m.max()
    F0, F1 = sfft.fftn(r0.astype(np.float64)), sfft.fftn(r1.astype(np.float64))
    P = F0 * F1.conj()
    eps = np.finfo(np.float32).eps
    Pn = P / np.maximum(np.abs(P), 100 * eps)
    ndim = len(shape)
    for slot, prod in ((0, P), (1, Pn)):
        cc = sfft.ifftn(prod)
        peak = np.array(np.unravel_index(np.argmax(np.abs(cc)), cc.shape))
        mid = np.array([np.fix(s / 2) for s in shape])
        wrapped = np.where(peak > mid, peak - np.array(shape), peak)
        assert np.array_equal(peaks[0, slot, 3 - ndim:], wrapped), (slot, peaks[0, slot], wrapped)
        u, R = plan.upsample, plan.region
        dftshift = np.fix(R / 2.0)
        off = dftshift - wrapped * u
        ref = sk._upsampled_dft(prod.conj(), R, u, off).conj()
        got = updft[0, slot].reshape((R,) * ndim)
        err = np.abs(got - ref).max() / np.abs(ref).max()
        assert err < 2e-3, (slot, err)
        assert np.argmax(np.abs(got)) == np.argmax(np.abs(ref))
    plan.close()


@pytest.mark.parametrize("name", ["strip_128x77", "strip_nan", "vol_24x64x51"])
def test_candidate_stages_match_scipy(reg, name):
    f, m, _ = cases.registration_cases()[name]
    plan, st = _plan_for(reg, f, m)
    r0, r1 = _rescaled(f), _rescaled(m)
    ndim = f.ndim
    rng = np.random.default_rng(3)
    cands = [np.zeros(ndim), np.full(ndim, 1.0), np.array([-3.3, 2.6, 1.5][:ndim]), -np.array([5.0, 7.5, 2.25][:ndim]),
             np.array([f.shape[d] - 1.0 for d in range(ndim)]), rng.uniform(-4, 4, ndim).astype(np.float32).astype(np.float64)]
    cs = plan.candidate_stats([0] * len(cands), np.array(cands))
    for t, got in zip(cands, cs):
        im1t = ndimage.affine_transform(r1, oreg.affine_from_translation(list(t)), order=1, mode="constant", cval=np.nan)
        valid = ~np.isnan(im1t)
        mask = valid & ~np.isnan(r0)
        assert got[0] == mask.sum() and got[1] == valid.sum(), (t, got[:2], mask.sum(), valid.sum())
        if valid.any():
            bb = oreg.get_bb_from_nanmask(valid)
            assert list(got[2 + 3 - ndim:5]) == [b[0] for b in bb]
            assert list(got[5 + 3 - ndim:8]) == [b[1] for b in bb]
        if mask.sum() > 50:
            bb0 = oreg.get_bb_from_nanmask(~np.isnan(r0))
            lo = [max(a[0], b[0]) for a, b in zip(bb0, bb)]
            hi = [min(a[1], b[1]) + 1 for a, b in zip(bb0, bb)]
            sl = tuple(slice(a, b) for a, b in zip(lo, hi))
            if min(b - a for a, b in zip(lo, hi)) >= 7:
                ref = sk.structural_similarity(np.nan_to_num(r0[sl]), np.nan_to_num(im1t[sl]), data_range=np.float32(1.0), win_size=7)
                got_s = plan.candidate_ssim([0], np.array([t]), np.array([[lo, hi]]), [7])[0]
                assert abs(got_s[0] - ref) < 2e-5, (t, got_s, ref)
                assert got_s[1] == np.nanmax(im1t[sl])
            rho = plan.spearman(0, t, int(mask.sum()))
            ref_rho = stats.spearmanr(r0[mask], im1t[mask] - 1).correlation
            assert abs(rho - ref_rho) < 1e-9, (t, rho, ref_rho)
    plan.close()


@pytest.mark.parametrize("name", sorted(cases.registration_cases().keys()))
def test_end_to_end_matches_reference_golden(reg, name, registration_golden):
    f, m, _ = cases.registration_cases()[name]
    res = reg.phase_correlation_registration(f, m)
    ref_aff = registration_golden[name + "/affine"]
    assert res["affine_matrix"].shape == ref_aff.shape
    assert np.abs(res["affine_matrix"] - ref_aff).max() <= TOL_PX
    assert abs(res["quality"] - float(registration_golden[name + "/quality"])) < 5e-3
    # identical when the same sub-pixel bin was chosen
    if np.array_equal(res["affine_matrix"], ref_aff):
        assert abs(res["quality"] - float(registration_golden[name + "/quality"])) < 1e-9


def test_reference_artificial_gt(reg):
    """_tests/test_registration.py:262-336 through the engine."""
    f, m, tr = cases.registration_cases()["blocks_100"]
    res = reg.phase_correlation_registration(f, m)
    A = np.eye(3)
    A[:2, 2] = tr
    assert np.allclose(res["affine_matrix"], A, atol=0.1)


def test_batched_pairs_mixed_shapes_match_oracle(reg):
    rng = np.random.default_rng(21)
    big = ndimage.gaussian_filter(rng.random((300, 300)), 1.5)
    fixed, moving, expect = [], [], []
    for k, (shape, shift) in enumerate([((200, 60), (1.4, -2.0)), ((60, 200), (-2.5, 0.7)), ((200, 60), (0.0, 3.0)), ((60, 200), (4.2, 4.9))]):
        sl = tuple(slice(20, 20 + s) for s in shape)
        fixed.append(big[sl].astype(np.float32))
        moving.append(ndimage.shift(big, shift, order=3)[sl].astype(np.float32))
    res = reg.register_pairs(fixed, moving)
    for f, m, r in zip(fixed, moving, res):
        ref = oreg.phase_correlation_registration(f, m)
        assert np.abs(r["affine_matrix"] - ref["affine_matrix"]).max() <= TOL_PX
        assert abs(r["quality"] - ref["quality"]) < 5e-3


def test_constant_image_guard(reg):
    a = np.zeros((32, 32), np.float32)
    b = np.random.default_rng(0).random((32, 32)).astype(np.float32)
    with pytest.warns(UserWarning):
        res = reg.phase_correlation_registration(a, b)
    assert np.array_equal(res["affine_matrix"], np.eye(3)) and np.isnan(res["quality"])


def test_synthetic_grid_pairs_recover_jitter(reg):
    """Size-independent property at C2-like crop shapes: the engine recovers
    the known integer jitter between overlapping synthetic tiles."""
    from multiview_stitcher_b200 import synthetic

    tile, ov = (512, 512), (77, 77)
    true, stage, idx = synthetic.grid_layout((2, 2), tile, ov, jitter=2, seed=5)
    tiles = [synthetic.make_tile(tile, o, np.float32, seed=5) for o in true]
    # horizontal pair (0,1): overlap strip in stage coordinates
    fixed = tiles[0][:, tile[1] - ov[1]:].contiguous()
    moving = tiles[1][:, : ov[1]].contiguous()
    res = reg.register_pairs([fixed], [moving])[0]
    d = (true[1] - stage[1].astype(np.int64)) - (true[0] - stage[0].astype(np.int64))
    # moving tile content at stage position is displaced by its jitter difference
    assert np.allclose(res["affine_matrix"][:2, 2], -d, atol=0.11), (res["affine_matrix"][:2, 2], d)
    assert res["quality"] > 0.9


def test_synthetic_3d_pair_matches_oracle(reg):
    """3-D face pair cut from the synthetic ground truth (integer jitter, default
    upsample_factor 2): engine == oracle within the half-pixel grid."""
    from multiview_stitcher_b200 import synthetic

    tile, ov = (40, 96, 64), (8, 16, 21)
    true, stage, idx = synthetic.grid_layout((1, 1, 2), tile, ov, jitter=2, seed=11)
    tiles = [synthetic.make_tile(tile, o, np.float32, seed=11) for o in true]
    fixed = tiles[0][:, :, tile[2] - ov[2]:].contiguous()
    moving = tiles[1][:, :, : ov[2]].contiguous()
    res = reg.register_pairs([fixed], [moving], return_details=True)[0]
    ref = oreg.phase_correlation_registration(fixed.cpu().numpy(), moving.cpu().numpy(), return_details=True)
    # upsample_factor 2: every shift lies on the 0.5 px grid.  Integer jitter puts the true peak ON a grid
    # node, where the two neighbouring upsampled-DFT samples tie to within float32 rounding: the only
    # admissible difference is such an adjacent-bin tie (exactly one bin), everything else must be equal
    # (the fractional-shift cases in test_gpu_subpixel.py hold 0.1 px)
    for a, b in zip(res["shift_candidates"], ref["shift_candidates"]):
        d = np.abs(np.asarray(a) - np.asarray(b))
        assert np.all((d <= 1e-6) | (np.abs(d - 0.5) <= 1e-6)), (a, b)
    d = np.abs(res["affine_matrix"][:3, 3] - ref["affine_matrix"][:3, 3])
    assert np.all((d <= 1e-6) | (np.abs(d - 0.5) <= 1e-6)), (res["affine_matrix"][:3, 3], ref["affine_matrix"][:3, 3])
    d = (true[1] - stage[1].astype(np.int64)) - (true[0] - stage[0].astype(np.int64))
    # both implementations land within one upsampling bin (0.5 px) of the true jitter
    assert np.abs(ref["affine_matrix"][:3, 3] + d).max() <= 0.5 + 1e-6
    assert np.abs(res["affine_matrix"][:3, 3] + d).max() <= 0.5 + 1e-6
    if np.array_equal(res["affine_matrix"], ref["affine_matrix"]):
        assert abs(res["quality"] - ref["quality"]) < 1e-9
