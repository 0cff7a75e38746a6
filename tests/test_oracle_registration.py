"""The oracle's registration restatement against fixtures produced by the
reference's own phase_correlation_registration candidate loop
(tests/golden/make_golden.py), the reference's artificial-GT test and exact
Fourier-shift self checks of the restated scikit-image functions."""

import warnings

import numpy as np
import pytest

import cases
from oracle import registration as oreg
from oracle import skimage_restated as sk


@pytest.mark.parametrize("name", sorted(cases.registration_cases().keys()))
def test_oracle_matches_reference_candidate_loop(name, registration_golden):
    f, m, _ = cases.registration_cases()[name]
    res = oreg.phase_correlation_registration(f, m)
    assert np.array_equal(res["affine_matrix"], registration_golden[name + "/affine"])
    assert res["quality"] == float(registration_golden[name + "/quality"])


def test_reference_artificial_gt_within_0p1px():
    """_tests/test_registration.py:262-336: recovered affine within 0.1 of GT
    (seed 0 -> translation (0.48813504, 2.15189366))."""
    f, m, tr = cases.registration_cases()["blocks_100"]
    res = oreg.phase_correlation_registration(f, m)
    A = np.eye(3)
    A[:2, 2] = tr
    assert np.allclose(res["affine_matrix"], A, atol=0.1)
    assert np.allclose(res["affine_matrix"][:2, 2], [0.5, 2.2], atol=1e-6)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("normalization", ["phase", None])
def test_exact_fourier_shift_recovered(dtype, normalization):
    rng = np.random.default_rng(5)
    shape = (128, 77)
    im = rng.random(shape)
    shift = (3.3, -2.6)
    F = np.fft.fftn(im)
    ky = np.fft.fftfreq(shape[0])[:, None]
    kx = np.fft.fftfreq(shape[1])[None, :]
    moved = np.fft.ifftn(F * np.exp(-2j * np.pi * (ky * shift[0] + kx * shift[1]))).real
    s, _, _ = sk.phase_cross_correlation(
        im.astype(dtype), moved.astype(dtype), upsample_factor=10, normalization=normalization
    )
    assert np.allclose(s, (-3.3, 2.6), atol=1e-5)
    assert s.dtype == dtype


def test_upsampled_dft_equals_zero_padded_ifft():
    rng = np.random.default_rng(6)
    a = rng.random((16, 12)) + 1j * rng.random((16, 12))
    up = 4
    got = sk._upsampled_dft(a, (16 * up, 12 * up), up, (0, 0))
    # direct evaluation
    yy = np.arange(16 * up)[:, None] * np.fft.fftfreq(16, up)[None, :]
    xx = np.arange(12 * up)[:, None] * np.fft.fftfreq(12, up)[None, :]
    ref = np.exp(-2j * np.pi * yy) @ a @ np.exp(-2j * np.pi * xx).T
    assert np.allclose(got, ref)


def test_masked_inverted_masks_give_zero_shift():
    """registration.py:433-443 passes isnan masks (True = invalid); the
    published masked algorithm then degenerates to a zero shift."""
    f, m, _ = cases.registration_cases()["strip_nan"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        s = sk.phase_cross_correlation(
            f, m, reference_mask=np.isnan(f), moving_mask=np.isnan(m)
        )[0]
    assert np.all(s == 0)


def test_ssim_identity_and_range():
    rng = np.random.default_rng(7)
    a = rng.random((40, 50)).astype(np.float32)
    assert sk.structural_similarity(a, a, data_range=np.float32(1.0)) == pytest.approx(1.0)
    b = rng.random((40, 50)).astype(np.float32)
    v = sk.structural_similarity(a, b, data_range=np.float32(1.0))
    assert -1 < v < 0.2


def test_rescale_intensity_float32():
    a = np.array([2.0, 4.0, 6.0, np.nan], np.float32)
    r = sk.rescale_intensity(a, in_range=(np.nanmin(a), np.nanmax(a)), out_range=(0, 1))
    assert r.dtype == np.float32
    assert np.array_equal(r[:3], np.array([0, 0.5, 1], np.float32)) and np.isnan(r[3])


def test_constant_guard_warns_and_returns_identity():
    """_tests/test_registration.py:682-708 / registration.py:1504-1530."""
    a = np.zeros((20, 20), np.float32)
    b = np.random.default_rng(0).random((20, 20)).astype(np.float32)
    with pytest.warns(UserWarning):
        res = oreg.dispatch_pairwise_reg_func(oreg.phase_correlation_registration, a, b)
    assert np.array_equal(res["affine_matrix"], np.eye(3)) and np.isnan(res["quality"])
