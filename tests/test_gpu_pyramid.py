"""Output pyramid on the GPU (multiview_stitcher_b200.pyramid) against the
oracle's ``np.mean(...).astype(dtype)`` coarsening: integer levels bit-exact, float32 levels
within 1e-6 relative (float64 accumulation here, float32 pairwise sums in numpy), NaNs
propagate; level geometry (origin / spacing) exact."""

import numpy as np
import pytest

from oracle import pyramid as opyr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.float32])
@pytest.mark.parametrize("shape", [(450, 517), (70, 230, 333)])
def test_pyramid_levels_match_oracle(dtype, shape):
    from multiview_stitcher_b200 import pyramid

    rng = np.random.default_rng(len(shape))
    if dtype == np.float32:
        data = rng.random(shape).astype(np.float32)
        data[3, 5:9] = np.nan
    else:
        data = rng.integers(0, np.iinfo(dtype).max, shape, endpoint=True).astype(dtype)
    dims = ["z", "y", "x"][-len(shape):]
    view = {"data": data, "origin": dict(zip(dims, (1.5, -2.0, 0.25)[-len(shape):])),
            "spacing": dict(zip(dims, (2.0, 0.5, 0.5)[-len(shape):]))}
    want = opyr.build_pyramid(view, min_shape=30)
    got = pyramid.build_pyramid(view, min_shape=30)
    assert len(got) == len(want) >= 3
    for g, w in zip(got, want):
        arr = g.tensor.cpu().numpy()
        assert arr.shape == w["data"].shape and arr.dtype == w["data"].dtype
        assert g.origin == w["origin"] and g.spacing == w["spacing"]
        if dtype == np.float32:
            np.testing.assert_array_equal(np.isnan(arr), np.isnan(w["data"]))
            np.testing.assert_allclose(arr, w["data"], rtol=1e-6, atol=0, equal_nan=True)
        else:
            np.testing.assert_array_equal(arr, w["data"])


def test_pyramid_of_a_fused_stack_default_levels():
    """Default rule (halve while the result stays above 100 px) on a fused 2-D grid."""
    from multiview_stitcher_b200 import fusion, geometry, pyramid, synthetic

    views, stage, true = synthetic.make_grid((2, 2), (300, 300), (40, 40), np.uint16, jitter=2, seed=5)
    osp = geometry.union_stack_props([v.bb() for v in views], true, views[0].spacing)
    fused, _ = fusion.fuse(views, true, osp, output_on_backend=True)
    dims = ["y", "x"]
    fv = fusion.DeviceView(fused, osp["origin"], osp["spacing"])
    levels = pyramid.build_pyramid(fv)
    host = {"data": fused.cpu().numpy(), "origin": osp["origin"], "spacing": osp["spacing"]}
    want = opyr.build_pyramid(host)
    assert [tuple(l.shape) for l in levels] == [w["data"].shape for w in want] and len(levels) == 3
    for g, w in zip(levels, want):
        np.testing.assert_array_equal(g.tensor.cpu().numpy(), w["data"])
        assert g.spacing == w["spacing"] and g.origin == w["origin"]
