"""Parity of the post-resample hooks (fusion_func / weights_func level) and of
content-weighted fusion with the reference-generated fixtures and the oracle."""

import numpy as np
import pytest
from scipy import ndimage

import cases
from oracle import fusion as of

pytestmark = pytest.mark.gpu

DIMS = ["z", "y", "x"]


@pytest.fixture(scope="module")
def eng():
    from multiview_stitcher_b200 import fusion

    return fusion


@pytest.fixture(scope="module")
def hooks():
    from multiview_stitcher_b200 import hooks

    return hooks


def _close(got, ref, exact=False):
    assert got.shape == ref.shape and got.dtype == ref.dtype
    if exact:
        assert np.array_equal(got, ref, equal_nan=True)
    elif ref.dtype.kind == "u":
        d = np.abs(got.astype(np.int64) - ref.astype(np.int64))
        assert d.max() <= 1, f"max LSB diff {d.max()}"
    else:
        tol = 1e-4 * np.abs(ref) + 1e-6 * np.nanmax(np.abs(ref))
        err = np.abs(got.astype(np.float64) - ref.astype(np.float64))
        assert np.array_equal(np.isnan(got), np.isnan(ref))
        assert np.all((err <= tol) | np.isnan(ref)), f"max err {np.nanmax(err)}"


@pytest.mark.parametrize("sigma", [1.0, 2.5, 5.0])
@pytest.mark.parametrize("shape", [(3, 40, 57), (2, 9, 33, 21)])
def test_gaussian_filter_matches_scipy_bitwise(hooks, shape, sigma):
    rng = np.random.default_rng(int(sigma * 10) + len(shape))
    a = rng.random(shape).astype(np.float32)
    got = hooks.gaussian_filter(a, sigma)
    ref = np.stack([ndimage.gaussian_filter(x, sigma, mode="reflect") for x in a])
    assert np.array_equal(got, ref)


def test_normalize_weights_matches_oracle(hooks):
    rng = np.random.default_rng(0)
    w = rng.random((4, 30, 41)).astype(np.float32)
    w[0, :5] = 0
    w[:, 10:12, 3:9] = 0
    w[1, 20:, :4] = np.nan
    assert np.array_equal(hooks.normalize_weights(w), of.normalize_weights(w), equal_nan=True)


@pytest.mark.parametrize("name", ["2d_f32_content", "3d_u16_content", "2d_f32_quad_lin", "3d_f32_affine_lin"])
def test_stack_hooks_match_reference_intermediates(hooks, name, fusion_golden):
    """fusion_func / weights_func hooks on the reference's own (V, *chunk) stacks."""
    tv, bw = fusion_golden[name + "/views"], fusion_golden[name + "/bw"]
    fw = fusion_golden[name + "/fw"] if name + "/fw" in fusion_golden else None
    got = hooks.weighted_average_fusion(tv, bw, fw)
    ref = of.weighted_average_fusion(tv, bw, fw)
    assert np.array_equal(got, ref, equal_nan=True)
    with np.errstate(all="ignore"):
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            assert np.array_equal(hooks.max_fusion(tv), of.max_fusion(tv), equal_nan=True)
            assert np.array_equal(hooks.simple_average_fusion(tv), of.simple_average_fusion(tv), equal_nan=True)
    if fw is not None:
        got_w = hooks.content_based(tv, bw, sigma_1=1, sigma_2=2)
        _close(got_w, fw)
        # where both agree the Gaussian pipeline is bit-faithful to scipy
        assert np.mean(got_w == fw) > 0.9 or np.nanmax(np.abs(got_w - fw)) < 1e-6


@pytest.mark.parametrize("name", ["2d_f32_content", "3d_u16_content"])
def test_content_weighted_fuse_matches_reference_golden(eng, name, fusion_golden):
    case = cases.fusion_cases()[name]
    kw = dict(case["kwargs"])
    views, params = case["views"], case["params"]
    bbs = [of.view_bb(v) for v in views]
    osp = of.calc_stack_properties(bbs, params, views[0]["spacing"])
    # the fixture is the reference's fuse_np on the whole stack (no halo)
    fused = eng.fuse_np(
        views, params, osp, weights_func=eng.content_based, weights_func_kwargs=kw["weights_func_kwargs"],
        interpolation_order=kw["interpolation_order"], full_view_bbs=bbs,
    )
    _close(fused, fusion_golden[name + "/fused"])


@pytest.mark.parametrize("name", ["2d_f32_content", "3d_u16_content"])
def test_content_weighted_fuse_np_halo_trim(eng, name, fusion_golden):
    case = cases.fusion_cases()[name]
    views, params = case["views"], case["params"]
    dims = DIMS[-views[0]["data"].ndim:]
    bbs = [of.view_bb(v) for v in views]
    osp = of.calc_stack_properties(bbs, params, views[0]["spacing"])
    ov = int(fusion_golden[name + "/sub_overlap"])
    start, shape = fusion_golden[name + "/sub_start"], fusion_golden[name + "/sub_shape"]
    hbb = {
        "origin": {d: osp["origin"][d] + (start[i] - ov) * osp["spacing"][d] for i, d in enumerate(dims)},
        "spacing": osp["spacing"],
        "shape": {d: int(shape[i]) + 2 * ov for i, d in enumerate(dims)},
    }
    got = eng.fuse_np(
        views, params, hbb, weights_func=eng.content_based,
        weights_func_kwargs=case["kwargs"]["weights_func_kwargs"], full_view_bbs=bbs,
        trim_overlap_in_pixels=ov, interpolation_order=1,
    )
    _close(got, fusion_golden[name + "/sub_fused"])


def test_chunked_content_fusion_matches_oracle(eng):
    """Chunks + halo (required_overlap = 2*sigma_2) reproduce the oracle's chunked run."""
    case = cases.fusion_cases()["2d_f32_content"]
    kw = case["kwargs"]["weights_func_kwargs"]
    cs = {"y": 24, "x": 32}
    got, _ = eng.fuse(case["views"], case["params"], weights_func=eng.content_based, weights_func_kwargs=kw, output_chunksize=cs)
    ref, _ = of.fuse(case["views"], case["params"], weights_func=of.content_based, weights_func_kwargs=kw, output_chunksize=cs)
    _close(got, ref)


def test_foreign_fusion_func_is_called_with_numpy(eng):
    """A user fusion_func (not a built-in) receives host stacks like in the reference."""
    case = cases.fusion_cases()["2d_f32_quad_lin"]
    seen = {}

    def my_fusion(transformed_views, blending_weights):
        seen["types"] = (type(transformed_views), type(blending_weights))
        return np.nansum(transformed_views * blending_weights, axis=0)

    got, _ = eng.fuse(case["views"], case["params"], fusion_func=my_fusion, output_chunksize={"y": 10_000, "x": 10_000},
                      blending_widths={"y": 4, "x": 6})
    ref, _ = eng.fuse(case["views"], case["params"], output_chunksize={"y": 10_000, "x": 10_000}, blending_widths={"y": 4, "x": 6})
    assert seen["types"] == (np.ndarray, np.ndarray)
    _close(got, ref)
