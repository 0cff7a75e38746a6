"""Seeded input definitions shared by ``make_golden.py`` (which runs the
reference on them) and the tests (which run the oracle and the CUDA engine on
them).  Pure numpy/scipy; no reference, no oracle imports."""

from __future__ import annotations

import numpy as np
from scipy import ndimage

DIMS = ["z", "y", "x"]


def _smooth(rng, shape, sigma=1.5):
    im = ndimage.gaussian_filter(rng.random(shape), sigma)
    im = (im - im.min()) / (im.max() - im.min())
    return im


def _view(data, origin, spacing):
    dims = DIMS[-data.ndim :]
    return {
        "data": data,
        "origin": dict(zip(dims, map(float, origin))),
        "spacing": dict(zip(dims, map(float, spacing))),
    }


def _translation(t):
    n = len(t)
    p = np.eye(n + 1)
    p[:n, n] = t
    return p


def _rot2d(angle, scale=(1.0, 1.0), t=(0.0, 0.0), center=(0.0, 0.0)):
    c, s = np.cos(angle), np.sin(angle)
    L = np.array([[c, -s], [s, c]]) @ np.diag(scale)
    p = np.eye(3)
    p[:2, :2] = L
    center = np.asarray(center, float)
    p[:2, 2] = np.asarray(t, float) + center - L @ center
    return p


def _rot3d(rotvec, scale=(1.0, 1.0, 1.0), t=(0, 0, 0), center=(0, 0, 0)):
    from scipy.spatial.transform import Rotation

    L = Rotation.from_rotvec(rotvec).as_matrix() @ np.diag(scale)
    p = np.eye(4)
    p[:3, :3] = L
    center = np.asarray(center, float)
    p[:3, 3] = np.asarray(t, float) + center - L @ center
    return p


def fusion_cases():
    """name -> dict(views, params, kwargs).  ``kwargs`` are fuse_np keyword
    arguments expressed by NAME for the function-valued ones
    (``fusion_func`` / ``weights_func``) so every implementation can map them
    to its own callables."""
    cases = {}

    # --- 2-D, two uint16 tiles, fractional translation, default blending ---
    rng = np.random.default_rng(1)
    gt = (_smooth(rng, (48, 90)) * 4000).astype(np.uint16)
    views = [
        _view(gt[:, :50].copy(), (0, 0), (1, 1)),
        _view(gt[:, 38:88].copy(), (0, 38), (1, 1)),
    ]
    params = [_translation((0, 0)), _translation((0.3, 1.7))]
    cases["2d_u16_pair_lin"] = dict(
        views=views, params=params, kwargs=dict(interpolation_order=1)
    )
    cases["2d_u16_pair_nn_max"] = dict(
        views=views,
        params=[_translation((0, 0)), _translation((1.0, 2.0))],
        kwargs=dict(interpolation_order=0, fusion_func="max_fusion"),
    )
    cases["2d_u16_pair_nn_wavg"] = dict(
        views=views,
        params=[_translation((0, 0)), _translation((1.0, 2.0))],
        kwargs=dict(interpolation_order=0),
    )

    # --- 2-D, four float32 tiles meeting in a corner, spacing 0.5 ---
    rng = np.random.default_rng(2)
    gt = _smooth(rng, (80, 84)).astype(np.float32)
    views, params = [], []
    for iy, ix in np.ndindex(2, 2):
        y0, x0 = iy * 34, ix * 36
        views.append(
            _view(gt[y0 : y0 + 46, x0 : x0 + 48].copy(), (y0 * 0.5, x0 * 0.5), (0.5, 0.5))
        )
        params.append(_translation(rng.uniform(-0.8, 0.8, 2)))
    cases["2d_f32_quad_lin"] = dict(
        views=views,
        params=params,
        kwargs=dict(interpolation_order=1, blending_widths={"y": 4, "x": 6}),
    )
    cases["2d_f32_quad_mean"] = dict(
        views=views,
        params=params,
        kwargs=dict(interpolation_order=1, fusion_func="simple_average_fusion"),
    )

    # --- 2-D general affines (rotation + anisotropic scale) ---
    rng = np.random.default_rng(3)
    views = [
        _view(_smooth(rng, (40, 56)).astype(np.float32), (0, 0), (1, 1)),
        _view(_smooth(rng, (44, 50)).astype(np.float32), (3, 20), (1.2, 0.9)),
        _view(_smooth(rng, (36, 40)).astype(np.float32), (25, 5), (1, 1)),
    ]
    params = [
        _rot2d(0.05, (1.02, 0.97), (0.4, -0.7), (20, 28)),
        _rot2d(-0.3, (1.0, 1.1), (2.2, 3.1), (25, 40)),
        _rot2d(np.pi / 2, (1.0, 1.0), (1.0, 2.0), (40, 25)),
    ]
    cases["2d_f32_affine_lin"] = dict(
        views=views, params=params, kwargs=dict(interpolation_order=1)
    )
    cases["2d_f32_affine_nn_max"] = dict(
        views=views,
        params=params,
        kwargs=dict(interpolation_order=0, fusion_func="max_fusion"),
    )

    # --- 3-D, two uint16 tiles, anisotropic spacing ---
    rng = np.random.default_rng(4)
    gt = (_smooth(rng, (14, 30, 52)) * 3000).astype(np.uint16)
    views = [
        _view(gt[:, :, :30].copy(), (0, 0, 0), (2, 0.5, 0.5)),
        _view(gt[:, :, 22:52].copy(), (0, 0, 11), (2, 0.5, 0.5)),
    ]
    params = [_translation((0, 0, 0)), _translation((0.6, -0.2, 0.35))]
    cases["3d_u16_pair_lin"] = dict(
        views=views, params=params, kwargs=dict(interpolation_order=1)
    )
    cases["3d_u16_pair_nn_max"] = dict(
        views=views,
        params=[_translation((0, 0, 0)), _translation((2.0, 0.5, -0.5))],
        kwargs=dict(interpolation_order=0, fusion_func="max_fusion"),
    )

    # --- 3-D general affines (multi-view) ---
    rng = np.random.default_rng(5)
    views = [
        _view(_smooth(rng, (12, 26, 30)).astype(np.float32), (0, 0, 0), (2, 1, 1)),
        _view(_smooth(rng, (14, 24, 28)).astype(np.float32), (1, 2, 3), (2, 1, 1)),
        _view(_smooth(rng, (12, 26, 30)).astype(np.float32), (0, 0, 0), (1.5, 1, 1.1)),
    ]
    params = [
        _rot3d((0.0, 0.0, 0.0)),
        _rot3d((0.04, -0.1, 0.07), (1.01, 0.98, 1.0), (0.5, -1.2, 2.3), (12, 13, 15)),
        _rot3d((0.0, np.pi / 2, 0.0), (1, 1, 1), (1.0, 0.0, 2.0), (10, 13, 15)),
    ]
    cases["3d_f32_affine_lin"] = dict(
        views=views, params=params, kwargs=dict(interpolation_order=1)
    )

    # --- content-based weights (Preibisch), 2-D and 3-D ---
    rng = np.random.default_rng(6)
    gt = _smooth(rng, (60, 70), 1.0) * 1000 + rng.random((60, 70)) * 50
    views = [
        _view(gt[:, :44].astype(np.float32), (0, 0), (1, 1)),
        _view(
            ndimage.gaussian_filter(gt, 1.5)[:, 26:70].astype(np.float32),
            (0, 26),
            (1, 1),
        ),
    ]
    params = [_translation((0, 0)), _translation((0.25, -0.4))]
    cases["2d_f32_content"] = dict(
        views=views,
        params=params,
        kwargs=dict(
            interpolation_order=1,
            weights_func="content_based",
            weights_func_kwargs=dict(sigma_1=1, sigma_2=2),
        ),
    )
    rng = np.random.default_rng(7)
    gt = _smooth(rng, (16, 30, 40), 1.0) * 1000 + rng.random((16, 30, 40)) * 50
    views = [
        _view(gt[:, :, :26].astype(np.uint16), (0, 0, 0), (1, 1, 1)),
        _view(
            ndimage.gaussian_filter(gt, 1.0)[:, :, 14:40].astype(np.uint16),
            (0, 0, 14),
            (1, 1, 1),
        ),
    ]
    params = [_translation((0, 0, 0)), _translation((0.0, 0.5, 0.25))]
    cases["3d_u16_content"] = dict(
        views=views,
        params=params,
        kwargs=dict(
            interpolation_order=1,
            weights_func="content_based",
            weights_func_kwargs=dict(sigma_1=1, sigma_2=2),
        ),
    )
    return cases


def blocks_image():
    """The reference's artificial-ground-truth image,
    _tests/test_registration.py:272-281."""
    im = np.zeros((100, 100), dtype=float)
    im[10:40, 20:40] = 1
    im[70:80, 60:90] = 1
    im[20:40, 60:70] = 1
    im[60:80, 20:40] = 1
    return ndimage.gaussian_filter(im, 3)


def registration_cases():
    """name -> (fixed float32, moving float32, expected translation or None)."""
    cases = {}
    # reference artificial GT (seed 0 -> (0.48813504, 2.15189366))
    im = blocks_image()
    np.random.seed(0)
    tr = np.random.random(2) * 10 - 10 / 2
    A = _translation(tr)
    imt = ndimage.affine_transform(im, np.linalg.inv(A))
    cases["blocks_100"] = (im.astype(np.float32), imt.astype(np.float32), tr)

    # textured overlap strips, integer + fractional shifts, non-pow2 sizes
    rng = np.random.default_rng(11)
    big = _smooth(rng, (160, 140), 1.2)
    f = big[10:138, 20:97]  # 128 x 77
    m = ndimage.shift(big, (-3.3, 2.6), order=3)[10:138, 20:97]
    cases["strip_128x77"] = (f.astype(np.float32), m.astype(np.float32), (3.3, -2.6))

    # with NaNs in the moving image (pre-transformed border)
    m2 = m.astype(np.float32).copy()
    m2[:, :4] = np.nan
    m2[-3:, :] = np.nan
    cases["strip_nan"] = (f.astype(np.float32), m2, (3.3, -2.6))

    # 3-D
    rng = np.random.default_rng(12)
    big = _smooth(rng, (30, 80, 60), 1.0)
    f = big[2:26, 5:69, 4:55]  # 24 x 64 x 51
    m = ndimage.shift(big, (1.0, -2.5, 1.5), order=3)[2:26, 5:69, 4:55]
    cases["vol_24x64x51"] = (
        f.astype(np.float32),
        m.astype(np.float32),
        (-1.0, 2.5, -1.5),
    )
    return cases


def pair_cases(extra=False):
    """name -> dict(views=[fixed, moving], affines=[A1, A2], kwargs) for pair
    preparation (``register_pair_of_msims``): two tiles cut from one smooth
    ground truth with a hidden jitter; ``affines`` are the stage transforms
    (view physical -> world) the overlap is computed in.  ``extra=True`` adds
    the geometry-heavy cases (mixed spacings, a rotated 3-D view, negative origins
    with a per-axis tolerance)."""
    cases = {}

    def cut(gt, start, shape, dtype):
        sl = tuple(slice(int(s), int(s) + int(n)) for s, n in zip(start, shape))
        t = gt[sl]
        if dtype == np.uint16:
            return np.round(t * 4000).astype(np.uint16)
        return t.astype(np.float32)

    # 2-D, pixel-aligned stage positions, neighbours along x, uint16
    rng = np.random.default_rng(21)
    gt = _smooth(rng, (140, 260), 1.3)
    a = cut(gt, (10, 8), (96, 128), np.uint16)
    b = cut(gt, (12, 8 + 102 - 3), (96, 128), np.uint16)  # true offset (2, 99), stage says (0, 102)
    cases["grid2d_x_u16"] = {
        "views": [_view(a, (0, 0), (1, 1)), _view(b, (0, 0), (1, 1))],
        "affines": [_translation((0, 0)), _translation((0, 102))],
        "kwargs": {"registration_binning": {"y": 1, "x": 1}},
    }

    # 2-D, neighbours along y, float32, spacing 0.5 and a stage position off the pixel grid
    rng = np.random.default_rng(22)
    gt = _smooth(rng, (230, 150), 1.3)
    a = cut(gt, (6, 10), (110, 120), np.float32)
    b = cut(gt, (6 + 88 + 2, 9), (110, 120), np.float32)
    cases["grid2d_y_f32_subpixel"] = {
        "views": [_view(a, (3.0, -2.0), (0.5, 0.5)), _view(b, (3.0, -2.0), (0.5, 0.5))],
        "affines": [_translation((0, 0)), _translation((88 * 0.5 + 0.15, 0.2))],
        "kwargs": {"registration_binning": {"y": 1, "x": 1}},
    }

    # 2-D, second view rotated by 3 degrees: polytope overlap, general pre-transform
    rng = np.random.default_rng(23)
    gt = _smooth(rng, (150, 150), 1.5)
    a = cut(gt, (5, 5), (100, 100), np.float32)
    b = cut(gt, (20, 40), (100, 100), np.float32)
    cases["rot2d_f32"] = {
        "views": [_view(a, (0, 0), (1, 1)), _view(b, (0, 0), (1, 1))],
        "affines": [_translation((0, 0)), _rot2d(np.deg2rad(3.0), t=(15.0, 35.0), center=(50, 50))],
        "kwargs": {"registration_binning": {"y": 1, "x": 1}},
    }

    # 3-D, anisotropic spacing, neighbours along x, uint16, overlap tolerance
    rng = np.random.default_rng(24)
    gt = _smooth(rng, (30, 60, 100), 1.0)
    a = cut(gt, (2, 4, 3), (24, 48, 56), np.uint16)
    b = cut(gt, (3, 5, 3 + 40 - 2), (24, 48, 56), np.uint16)
    cases["grid3d_x_u16_tol"] = {
        "views": [_view(a, (0, 0, 0), (2, 1, 1)), _view(b, (0, 0, 0), (2, 1, 1))],
        "affines": [_translation((0, 0, 0)), _translation((0, 0, 40))],
        "kwargs": {"registration_binning": {"z": 1, "y": 1, "x": 1}, "overlap_tolerance": 2.0},
    }

    # 2-D, binned 2 x 2 (uint16 means are truncated), odd tile size (trimmed)
    rng = np.random.default_rng(25)
    gt = _smooth(rng, (260, 420), 2.5)
    a = cut(gt, (8, 8), (201, 223), np.uint16)
    b = cut(gt, (10, 8 + 180 + 4), (201, 223), np.uint16)
    cases["grid2d_binned_u16"] = {
        "views": [_view(a, (0, 0), (1, 1)), _view(b, (0, 0), (1, 1))],
        "affines": [_translation((0, 0)), _translation((0, 180))],
        "kwargs": {"registration_binning": {"y": 2, "x": 2}},
    }
    if not extra:
        return cases

    # moving view sampled twice as finely as the fixed one: common grid at the coarser spacing
    rng = np.random.default_rng(26)
    gt = _smooth(rng, (200, 300), 2.0)
    a = cut(gt, (10, 10), (120, 140), np.float32)
    fine = ndimage.zoom(gt, 2, order=1)
    b = fine[2 * 14 : 2 * 14 + 240, 2 * 118 : 2 * 118 + 280].astype(np.float32)
    cases["mixed_spacing2d_f32"] = {
        "views": [_view(a, (0, 0), (1, 1)), _view(b, (0.25, -0.25), (0.5, 0.5))],
        "affines": [_translation((0, 0)), _translation((3.0, 110.0))],
        "kwargs": {"registration_binning": {"y": 1, "x": 1}},
    }

    # 3-D, anisotropic, second view rotated about z and slightly tilted
    rng = np.random.default_rng(27)
    gt = _smooth(rng, (30, 90, 120), 1.0)
    a = cut(gt, (2, 5, 5), (24, 64, 72), np.uint16)
    b = cut(gt, (3, 12, 50), (24, 64, 60), np.uint16)
    cases["rot3d_u16"] = {
        "views": [_view(a, (0, 0, 0), (2, 1, 1)), _view(b, (1.0, -3.0, 2.0), (2, 1, 1))],
        "affines": [
            _translation((0, 0, 0)),
            _rot3d(np.deg2rad([4.0, 0.5, -0.3]), t=(1.0, 9.0, 44.0), center=(24, 32, 30)),
        ],
        "kwargs": {"registration_binning": {"z": 1, "y": 1, "x": 1}},
    }

    # negative origins, spacing 0.65, tolerance on one axis only
    rng = np.random.default_rng(28)
    gt = _smooth(rng, (160, 260), 1.5)
    a = cut(gt, (8, 8), (100, 128), np.uint16)
    b = cut(gt, (20, 8 + 96), (100, 128), np.uint16)
    cases["neg_origin2d_u16_tol_x"] = {
        "views": [_view(a, (-40.3, -100.0), (0.65, 0.65)), _view(b, (-40.3, -100.0), (0.65, 0.65))],
        "affines": [_translation((0.0, 0.0)), _translation((12 * 0.65 + 0.1, 96 * 0.65))],
        "kwargs": {"registration_binning": {"y": 1, "x": 1}, "overlap_tolerance": {"x": 3.0}},
    }
    return cases
