"""The oracle's fusion restatement against (a) fixtures produced by the
reference's own fuse_np (tests/golden/make_golden.py) and (b) the reference's
known-answer tests (_tests/test_fusion.py), replayed through the oracle."""

import numpy as np
import pytest

import cases
from oracle import fusion as of

DIMS = ["z", "y", "x"]


def _funcs(kwargs):
    kwargs = dict(kwargs)
    if "fusion_func" in kwargs:
        kwargs["fusion_func"] = getattr(of, kwargs["fusion_func"])
    if "weights_func" in kwargs:
        kwargs["weights_func"] = getattr(of, kwargs["weights_func"])
    return kwargs


@pytest.mark.parametrize("name", sorted(cases.fusion_cases().keys()))
def test_oracle_matches_reference_fuse_np(name, fusion_golden):
    case = cases.fusion_cases()[name]
    views, params = case["views"], case["params"]
    ndim = views[0]["data"].ndim
    dims = DIMS[-ndim:]
    bbs = [of.view_bb(v) for v in views]
    osp = of.calc_stack_properties(bbs, params, views[0]["spacing"])
    # output stack geometry is the reference's
    assert np.array_equal(
        [osp["shape"][d] for d in dims], fusion_golden[name + "/shape"]
    )
    assert np.allclose(
        [osp["origin"][d] for d in dims], fusion_golden[name + "/origin"], rtol=0, atol=0
    )
    fused, tv, bw, fw = of.fuse_np(
        views, params, osp, full_view_bbs=bbs, return_intermediates=True,
        **_funcs(case["kwargs"])
    )
    ref = fusion_golden[name + "/fused"]
    assert fused.dtype == ref.dtype
    # same library calls in the same order: bit-identical
    assert np.array_equal(fused, ref)
    if name + "/views" in fusion_golden:
        assert np.array_equal(tv, fusion_golden[name + "/views"], equal_nan=True)
        assert np.array_equal(bw, fusion_golden[name + "/bw"], equal_nan=True)
    if name + "/fw" in fusion_golden:
        assert np.array_equal(fw, fusion_golden[name + "/fw"], equal_nan=True)


@pytest.mark.parametrize("name", ["2d_f32_content", "3d_u16_content"])
def test_oracle_halo_trim_matches_reference(name, fusion_golden):
    case = cases.fusion_cases()[name]
    views, params = case["views"], case["params"]
    dims = DIMS[-views[0]["data"].ndim :]
    bbs = [of.view_bb(v) for v in views]
    osp = of.calc_stack_properties(bbs, params, views[0]["spacing"])
    ov = int(fusion_golden[name + "/sub_overlap"])
    start = fusion_golden[name + "/sub_start"]
    shape = fusion_golden[name + "/sub_shape"]
    hbb = {
        "origin": {
            d: osp["origin"][d] + (start[i] - ov) * osp["spacing"][d]
            for i, d in enumerate(dims)
        },
        "spacing": osp["spacing"],
        "shape": {d: int(shape[i]) + 2 * ov for i, d in enumerate(dims)},
    }
    sub = of.fuse_np(
        views, params, hbb, full_view_bbs=bbs, trim_overlap_in_pixels=ov,
        **_funcs(case["kwargs"])
    )
    assert np.array_equal(sub, fusion_golden[name + "/sub_fused"])


# ---- reference KATs (_tests/test_fusion.py) replayed through oracle.fuse ----


def _v(data, origin, spacing):
    dims = DIMS[-data.ndim :]
    return {"data": data, "origin": dict(zip(dims, origin)), "spacing": dict(zip(dims, spacing))}


def test_kat_max_fusion_two_tiles():
    """_tests/test_fusion.py:204-237."""
    views = [
        _v(np.ones((8, 8)) * value, (0.0, x0), (1.0, 1.0))
        for value, x0 in [(1, 0.0), (2, 6.0)]
    ]
    fused, osp = of.fuse(
        views, [np.eye(3)] * 2, fusion_func=of.max_fusion, output_chunksize={"y": 4, "x": 4}
    )
    assert fused.shape == (8, 14)
    assert np.all(fused[:, :6] == 1) and np.all(fused[:, 6:] == 2)


def test_kat_nn_singleton_spacing():
    """_tests/test_fusion.py:480-530."""
    view = _v(np.ones((2, 20), dtype=np.uint16), (0.0, 0.0), (0.3, 0.3))
    osp = {
        "origin": {"y": 0.0, "x": -2.7},
        "spacing": {"y": 0.3, "x": 0.3},
        "shape": {"y": 2, "x": 29},
    }
    fused, _ = of.fuse(
        [view], [np.eye(3)], output_stack_properties=osp, fusion_func=of.max_fusion,
        interpolation_order=0, output_chunksize={"y": 2, "x": 10},
    )
    expect = np.tile(np.concatenate([np.zeros(9, np.uint16), np.ones(20, np.uint16)]), (2, 1))
    assert np.array_equal(fused, expect)


def test_kat_nn_large_origin_roundoff():
    """_tests/test_fusion.py:533-573."""
    origin = 861.5120670572916
    scale = 0.13810709635416665
    view = _v(np.ones((2, 4084), dtype=np.uint16), (0.0, origin), (scale, scale))
    osp = {
        "origin": {"y": 0.0, "x": origin - 9 * scale},
        "spacing": {"y": scale, "x": scale},
        "shape": {"y": 2, "x": 4093},
    }
    fused, _ = of.fuse(
        [view], [np.eye(3)], output_stack_properties=osp, fusion_func=of.max_fusion,
        interpolation_order=0, output_chunksize={"y": 2, "x": 4084},
    )
    expect = np.tile(np.concatenate([np.zeros(9, np.uint16), np.ones(4084, np.uint16)]), (2, 1))
    assert np.array_equal(fused, expect)


def test_kat_identity_single_view_bit_exact():
    """_tests/test_fusion.py:357-360."""
    rng = np.random.default_rng(0)
    data = rng.integers(0, 60000, (33, 47)).astype(np.uint16)
    fused, _ = of.fuse([_v(data, (0, 0), (1, 1))], [np.eye(3)])
    assert np.array_equal(fused, data)


def test_kat_fractional_translation_shape():
    """_tests/test_fusion.py:756-810: four 10x10 tiles on an 8.5 px pitch."""
    views = []
    for i, (iy, ix) in enumerate(np.ndindex(2, 2)):
        views.append(_v(np.full((10, 10), i + 1, np.uint16), (iy * 8.5, ix * 8.5), (1, 1)))
    fused, osp = of.fuse(views, [np.eye(3)] * 4)
    assert fused.shape == (18, 18)
    assert fused.max() == 4 and fused.min() > 0


def test_blending_weights_cover_views():
    """_tests/test_weights.py:43-133 property: normalised weights sum to 0 or
    1 and are positive wherever a view is valid (2-D, 3-D, random affines)."""
    for ndim in (2, 3):
        case = cases.fusion_cases()["2d_f32_affine_lin" if ndim == 2 else "3d_f32_affine_lin"]
        views, params = case["views"], case["params"]
        bbs = [of.view_bb(v) for v in views]
        osp = of.calc_stack_properties(bbs, params, views[0]["spacing"])
        _, tv, bw, _ = of.fuse_np(views, params, osp, full_view_bbs=bbs, return_intermediates=True)
        s = np.nansum(bw, axis=0)
        assert np.all((np.abs(s) < 1e-5) | (np.abs(s - 1) < 1e-5))
        assert np.all(bw[~np.isnan(tv)] > 0)


def test_kat_fused_field_slice_is_aligned():
    """_tests/test_fusion.py:932-987 (arithmetic half): a one-plane output stack placed on the
    view's z index 1 (anisotropic spacing, translated view, order 1) reproduces the input value
    exactly -- the slice is aligned with the input grid, nothing is interpolated across planes."""
    spacing = {"z": 3.5, "y": 2.5, "x": 4.5}
    trans = {"z": 1.3, "y": 1.0, "x": 2.0}
    data = np.zeros((5, 50, 100), dtype=np.float32)
    data[1] = 1.0  # only the plane the output slice coincides with carries the value
    view = {"data": data, "origin": {d: 0.0 for d in DIMS}, "spacing": spacing}
    param = np.eye(4)
    param[:3, 3] = [trans[d] for d in DIMS]
    osp = {"spacing": spacing, "origin": {d: trans[d] + 1 * spacing[d] for d in DIMS},
           "shape": {"z": 1, "y": 40, "x": 70}}
    fused, _ = of.fuse([view], [param], output_stack_properties=osp, interpolation_order=1)
    assert fused.shape == (1, 40, 70)
    assert not np.any(fused - 1.0)
