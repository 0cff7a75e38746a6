"""Parity of the CUDA fused resample-blend path (through the C ABI) with the
reference-generated golden fixtures and the oracle.

Bars (BASELINE.json north_star): float32 fused voxels within 1e-4 relative
(|got-ref| <= 1e-4*|ref| + 1e-6*max|ref|); uint16 nearest-neighbour bit-exact
under max_fusion and on single-view voxels; <= 1 LSB on blended uint16 voxels
(hard part 3)."""

import numpy as np
import pytest

import cases
from oracle import fusion as of

pytestmark = pytest.mark.gpu

DIMS = ["z", "y", "x"]


@pytest.fixture(scope="module")
def eng():
    from multiview_stitcher_b200 import fusion

    return fusion


def _ofuncs(kwargs):
    kwargs = dict(kwargs)
    if "fusion_func" in kwargs:
        kwargs["fusion_func"] = getattr(of, kwargs["fusion_func"])
    return kwargs


def _efuncs(eng, kwargs):
    kwargs = dict(kwargs)
    if "fusion_func" in kwargs:
        kwargs["fusion_func"] = getattr(eng, kwargs["fusion_func"])
    return kwargs


def assert_fused_close(got, ref, exact=False):
    assert got.shape == ref.shape and got.dtype == ref.dtype
    if exact:
        assert np.array_equal(got, ref)
    elif ref.dtype.kind == "u":
        d = np.abs(got.astype(np.int64) - ref.astype(np.int64))
        assert d.max() <= 1, f"max LSB diff {d.max()}"
    else:
        tol = 1e-4 * np.abs(ref) + 1e-6 * np.abs(ref).max()
        err = np.abs(got.astype(np.float64) - ref.astype(np.float64))
        assert np.all(err <= tol), f"max err {err.max()} at {np.unravel_index(err.argmax(), err.shape)}"


PLAIN = [n for n in sorted(cases.fusion_cases()) if "content" not in n]


@pytest.mark.parametrize("name", PLAIN)
def test_fuse_matches_reference_golden(eng, name, fusion_golden):
    case = cases.fusion_cases()[name]
    fused, osp = eng.fuse(case["views"], case["params"], **_efuncs(eng, case["kwargs"]))
    ref = fusion_golden[name + "/fused"]
    dims = DIMS[-ref.ndim :]
    assert np.array_equal([osp["shape"][d] for d in dims], fusion_golden[name + "/shape"])
    assert_fused_close(fused, ref, exact="nn_max" in name)


@pytest.mark.parametrize("name", PLAIN)
@pytest.mark.parametrize("chunk", [16, 37])
def test_chunked_equals_golden(eng, name, chunk, fusion_golden):
    case = cases.fusion_cases()[name]
    ndim = case["views"][0]["data"].ndim
    cs = {d: (chunk if d != "z" else 5) for d in DIMS[-ndim:]}
    fused, _ = eng.fuse(case["views"], case["params"], output_chunksize=cs, **_efuncs(eng, case["kwargs"]))
    assert_fused_close(fused, fusion_golden[name + "/fused"], exact="nn_max" in name)


def test_single_view_voxels_bit_exact_uint16(eng, fusion_golden):
    name = "2d_u16_pair_nn_wavg"
    case = cases.fusion_cases()[name]
    fused, _ = eng.fuse(case["views"], case["params"], **_efuncs(eng, case["kwargs"]))
    ref = fusion_golden[name + "/fused"]
    tv = fusion_golden[name + "/views"]
    single = (~np.isnan(tv)).sum(0) == 1
    assert single.sum() > 100
    assert np.array_equal(fused[single], ref[single])


@pytest.mark.parametrize("name", ["2d_u16_pair_lin", "3d_u16_pair_lin", "2d_f32_affine_lin"])
def test_fuse_np_halo_trim_matches_oracle(eng, name):
    case = cases.fusion_cases()[name]
    views, params = case["views"], case["params"]
    dims = DIMS[-views[0]["data"].ndim :]
    bbs = [of.view_bb(v) for v in views]
    osp = of.calc_stack_properties(bbs, params, views[0]["spacing"])
    ov = 3
    hbb = {
        "origin": {d: osp["origin"][d] + (4 - ov) * osp["spacing"][d] for d in dims},
        "spacing": osp["spacing"],
        "shape": {d: min(12, osp["shape"][d] - 6) + 2 * ov for d in dims},
    }
    ref = of.fuse_np(views, params, hbb, full_view_bbs=bbs, trim_overlap_in_pixels=ov, **_ofuncs(case["kwargs"]))
    got = eng.fuse_np(views, params, hbb, full_view_bbs=bbs, trim_overlap_in_pixels=ov, **_efuncs(eng, case["kwargs"]))
    assert_fused_close(got, ref)


def test_fuse_np_on_view_slices_matches_oracle(eng):
    """fuse_np is handed SLICES of the views by the reference's planner
    (fusion/_core.py:1348-1462); slice origin != view origin."""
    case = cases.fusion_cases()["2d_f32_quad_lin"]
    views, params = case["views"], case["params"]
    bbs = [of.view_bb(v) for v in views]
    osp = of.calc_stack_properties(bbs, params, views[0]["spacing"])
    chunk = {
        "origin": {d: osp["origin"][d] + 20 * osp["spacing"][d] for d in "yx"},
        "spacing": osp["spacing"],
        "shape": {"y": 30, "x": 40},
    }
    slices = []
    for v in views:
        sl = (slice(3, 44), slice(2, 46))
        slices.append(
            {
                "data": v["data"][sl],
                "origin": {"y": v["origin"]["y"] + 3 * 0.5, "x": v["origin"]["x"] + 2 * 0.5},
                "spacing": v["spacing"],
            }
        )
    kw = dict(interpolation_order=1, blending_widths={"y": 4, "x": 6})
    ref = of.fuse_np(slices, params, chunk, full_view_bbs=bbs, **kw)
    got = eng.fuse_np(slices, params, chunk, full_view_bbs=bbs, **kw)
    assert_fused_close(got, ref)


# ---- reference KATs (_tests/test_fusion.py) through the engine ----


def _v(data, origin, spacing):
    dims = DIMS[-data.ndim :]
    return {"data": data, "origin": dict(zip(dims, origin)), "spacing": dict(zip(dims, spacing))}


def test_kat_max_fusion_two_tiles(eng):
    views = [_v(np.full((8, 8), value, np.float32), (0.0, x0), (1.0, 1.0)) for value, x0 in [(1, 0.0), (2, 6.0)]]
    fused, _ = eng.fuse(views, [np.eye(3)] * 2, fusion_func=eng.max_fusion, output_chunksize={"y": 4, "x": 4})
    assert fused.shape == (8, 14)
    assert np.all(fused[:, :6] == 1) and np.all(fused[:, 6:] == 2)


def test_kat_nn_singleton_spacing(eng):
    view = _v(np.ones((2, 20), dtype=np.uint16), (0.0, 0.0), (0.3, 0.3))
    osp = {"origin": {"y": 0.0, "x": -2.7}, "spacing": {"y": 0.3, "x": 0.3}, "shape": {"y": 2, "x": 29}}
    fused, _ = eng.fuse([view], [np.eye(3)], output_stack_properties=osp, fusion_func=eng.max_fusion,
                        interpolation_order=0, output_chunksize={"y": 2, "x": 10})
    expect = np.tile(np.concatenate([np.zeros(9, np.uint16), np.ones(20, np.uint16)]), (2, 1))
    assert np.array_equal(fused, expect)


def test_kat_nn_large_origin_roundoff(eng):
    origin = 861.5120670572916
    scale = 0.13810709635416665
    view = _v(np.ones((2, 4084), dtype=np.uint16), (0.0, origin), (scale, scale))
    osp = {"origin": {"y": 0.0, "x": origin - 9 * scale}, "spacing": {"y": scale, "x": scale}, "shape": {"y": 2, "x": 4093}}
    fused, _ = eng.fuse([view], [np.eye(3)], output_stack_properties=osp, fusion_func=eng.max_fusion,
                        interpolation_order=0, output_chunksize={"y": 2, "x": 4084})
    expect = np.tile(np.concatenate([np.zeros(9, np.uint16), np.ones(4084, np.uint16)]), (2, 1))
    assert np.array_equal(fused, expect)


def test_kat_identity_single_view_bit_exact(eng):
    rng = np.random.default_rng(0)
    data = rng.integers(0, 60000, (33, 47)).astype(np.uint16)
    fused, _ = eng.fuse([_v(data, (0, 0), (1, 1))], [np.eye(3)])
    assert np.array_equal(fused, data)


def test_kat_fractional_translation_shape(eng):
    views = [_v(np.full((10, 10), i + 1, np.uint16), (iy * 8.5, ix * 8.5), (1, 1)) for i, (iy, ix) in enumerate(np.ndindex(2, 2))]
    fused, _ = eng.fuse(views, [np.eye(3)] * 4)
    assert fused.shape == (18, 18)
    assert fused.max() == 4 and fused.min() > 0


def test_random_affine_nn_bit_exact_vs_scipy(eng):
    """Order-0 picks and the outside predicate are bit-identical to scipy for
    random general affines (float64 coordinate path)."""
    rng = np.random.default_rng(42)
    for ndim in (2, 3):
        shape = (31, 45) if ndim == 2 else (9, 21, 25)
        data = rng.integers(1, 60000, shape).astype(np.uint16)
        for _ in range(4):
            p = np.eye(ndim + 1)
            p[:ndim, :ndim] += rng.uniform(-0.3, 0.3, (ndim, ndim))
            p[:ndim, ndim] = rng.uniform(-5, 5, ndim)
            view = _v(data, rng.uniform(-3, 3, ndim), rng.uniform(0.5, 1.5, ndim))
            bbs = [of.view_bb(view)]
            osp = of.calc_stack_properties(bbs, [p], view["spacing"])
            ref = of.fuse_np([view], [p], osp, full_view_bbs=bbs, fusion_func=of.max_fusion, interpolation_order=0)
            got = eng.fuse_np([view], [p], osp, full_view_bbs=bbs, fusion_func=eng.max_fusion, interpolation_order=0)
            assert np.array_equal(got, ref)


def test_synthetic_grid_full_size_properties(eng):
    """Size-independent properties at a larger size than the oracle handles
    quickly: registered synthetic tiles cut from one ground truth fuse back to
    the ground truth exactly under max fusion (order 0, integer shifts) and to
    within 1 LSB under blending."""
    from multiview_stitcher_b200 import synthetic

    views, stage, true = synthetic.make_grid((3, 3), (512, 512), (77, 77), np.uint16, jitter=2, seed=3)
    fused, osp = eng.fuse(views, true, fusion_func=eng.max_fusion, interpolation_order=0, output_on_backend=True)
    org = [int(round(osp["origin"][d])) for d in "yx"]
    gt = synthetic.make_tile(tuple(fused.shape), org, np.uint16, seed=3)
    f, g = fused.cpu().numpy(), gt.cpu().numpy()
    covered = f > 0
    assert covered.mean() > 0.98
    assert np.array_equal(f[covered], g[covered])
    fused2, _ = eng.fuse(views, true, interpolation_order=1)
    d = np.abs(fused2.astype(np.int64) - g.astype(np.int64))[covered]
    assert d.max() <= 1


# ---- translation fast path (bulk-copy stencil kernel) vs general kernel / oracle ----


def _grid_views(rng, ndim, dtype, tile, grid, ov, jitter):
    import itertools

    views, params = [], []
    pitch = [t - o for t, o in zip(tile, ov)]
    full = [p * (g - 1) + t + 8 for p, g, t in zip(pitch, grid, tile)]
    gt = ndimage_smooth(rng, full)
    if np.dtype(dtype).kind == "u":
        gt = (gt * (250 if dtype == np.uint8 else 4000)).astype(dtype)
    else:
        gt = gt.astype(dtype)
    for idx in itertools.product(*[range(g) for g in grid]):
        org = [4 + i * p for i, p in zip(idx, pitch)]
        sl = tuple(slice(o, o + t) for o, t in zip(org, tile))
        views.append(_v(np.ascontiguousarray(gt[sl]), [float(o) for o in org], [1.0] * ndim))
        p = np.eye(ndim + 1)
        p[:ndim, ndim] = jitter(rng, ndim)
        params.append(p)
    return views, params


def ndimage_smooth(rng, shape):
    from scipy import ndimage

    im = ndimage.gaussian_filter(rng.random(shape), 1.2)
    return (im - im.min()) / (im.max() - im.min())


STENCIL_CASES = [
    (2, np.float32, (72, 136), (2, 3), (20, 24)),
    (2, np.uint16, (64, 144), (3, 2), (16, 40)),
    (2, np.uint8, (48, 160), (2, 2), (12, 32)),
    (3, np.uint16, (12, 20, 136), (2, 2, 2), (4, 6, 24)),
    (3, np.float32, (10, 24, 132), (1, 2, 2), (2, 8, 20)),
]


@pytest.mark.parametrize("ndim,dtype,tile,grid,ov", STENCIL_CASES)
@pytest.mark.parametrize("order,func", [(1, "weighted_average_fusion"), (0, "weighted_average_fusion"), (0, "max_fusion"), (1, "simple_average_fusion")])
@pytest.mark.parametrize("integer_shift", [False, True])
def test_stencil_path_matches_general_and_oracle(eng, monkeypatch, ndim, dtype, tile, grid, ov, order, func, integer_shift):
    rng = np.random.default_rng(ndim * 100 + np.dtype(dtype).itemsize + order)
    jit = (lambda r, n: r.integers(-2, 3, n).astype(float)) if integer_shift else (lambda r, n: np.round(r.uniform(-2, 2, n), 3))
    views, params = _grid_views(rng, ndim, dtype, tile, grid, ov, jit)
    cs = {d: c for d, c in zip(DIMS[-ndim:], (7, 40, 200)[-ndim:])}
    kw = dict(interpolation_order=order, output_chunksize=cs)
    monkeypatch.delenv("MVS_FUSE_GENERIC", raising=False)
    fast, _ = eng.fuse(views, params, fusion_func=getattr(eng, func), **kw)
    monkeypatch.setenv("MVS_FUSE_GENERIC", "1")
    slow, _ = eng.fuse(views, params, fusion_func=getattr(eng, func), **kw)
    monkeypatch.delenv("MVS_FUSE_GENERIC", raising=False)
    if order == 0 and func == "max_fusion":
        assert np.array_equal(fast, slow)
    else:
        assert_fused_close(fast, slow)
    ref, _ = of.fuse(views, params, fusion_func=getattr(of, func), interpolation_order=order)
    assert_fused_close(fast, ref, exact=(order == 0 and func == "max_fusion"))


def test_stencil_path_is_taken(eng):
    """The plan schedules translation-only, aligned views on the stencil kernel."""
    import ctypes

    from multiview_stitcher_b200 import synthetic

    views, stage, true = synthetic.make_grid((2, 2), (256, 256), (40, 40), np.float32, jitter=2, seed=9)
    osp = of.calc_stack_properties([v.bb() for v in views], true, views[0].spacing)
    plan = eng.FusionPlan(views, true, osp)
    assert plan.launches_per_run == 1
    plan.close()


def test_pipelined_host_fusion_equals_plain(eng):
    """fuse(out_host=pinned) overlaps H2D / kernels / D2H band by band and must
    return exactly what the plain path returns."""
    import torch

    rng = np.random.default_rng(8)
    views, params = _grid_views(rng, 2, np.uint16, (96, 160), (3, 2), (20, 32), lambda r, n: np.round(r.uniform(-2, 2, n), 2))
    cs = {"y": 64, "x": 128}
    ref, osp = eng.fuse(views, params, output_chunksize=cs)
    out_host = torch.empty(ref.shape, dtype=torch.uint16).pin_memory()
    got, osp2 = eng.fuse(views, params, output_chunksize=cs, out_host=out_host)
    assert osp2 == osp
    assert np.array_equal(got.numpy(), ref)  # out_host (a pinned tensor) is returned
    got2, _ = eng.fuse(views, params, output_chunksize=cs, out_host=np.empty_like(ref), fusion_func=eng.max_fusion)
    ref2, _ = eng.fuse(views, params, output_chunksize=cs, fusion_func=eng.max_fusion)
    assert np.array_equal(got2, ref2)


def test_host_fuser_reuse(eng):
    """A HostFuser fuses changing data of one geometry; results equal fuse()."""
    import torch

    rng = np.random.default_rng(9)
    views, params = _grid_views(rng, 2, np.float32, (64, 132), (2, 3), (12, 28), lambda r, n: np.round(r.uniform(-2, 2, n), 2))
    fuser = eng.HostFuser(views, params, output_chunksize={"y": 48, "x": 128})
    out = torch.empty(fuser.out_shape, dtype=torch.float32).pin_memory()
    for rep in range(2):
        for v in views:
            v["data"] = (v["data"] * (1.0 + rep)).astype(np.float32)
        ref, _ = eng.fuse(views, params, output_chunksize={"y": 48, "x": 128})
        got = fuser(views, out).numpy()
        assert np.array_equal(got, ref)
    fuser.close()


@pytest.mark.parametrize("kind", ["translation", "affine"])
@pytest.mark.parametrize("func", ["weighted_average_fusion", "max_fusion", "simple_average_fusion"])
def test_nan_data_inside_float_views_is_masked_per_voxel(kind, func):
    """float32 views carrying NaN DATA (not just NaN outside): the reference zeroes the blending
    weight where the transformed view is NaN (fusion/_core.py:1648) and its fusion functions are
    nan-aware, so the other view's data survives in the overlap and a lone NaN voxel becomes 0."""
    from multiview_stitcher_b200 import fusion
    from oracle import fusion as of

    rng = np.random.default_rng(3)
    views, params = [], []
    for k in range(2):
        data = (rng.random((48, 64)) * 100 + 10).astype(np.float32)
        views.append({"data": data, "origin": {"y": 0.0, "x": 40.0 * k}, "spacing": {"y": 1.0, "x": 1.0}})
        p = np.eye(3)
        p[:2, 2] = (0.3 * k, -0.45 * k)
        if kind == "affine":
            a = np.deg2rad(3.0 * (k + 1))
            p[:2, :2] = [[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]]
        params.append(p)
    views[0]["data"][10:20, 45:60] = np.nan   # inside the overlap: view 1 must fill it
    views[0]["data"][30:34, 5:9] = np.nan     # where view 0 is alone: fused 0
    views[1]["data"][25, 3] = np.nan
    ofunc = getattr(of, func)
    efunc = getattr(fusion, func)
    ref, osp = of.fuse(views, params, fusion_func=ofunc)
    got, _ = fusion.fuse(views, params, output_stack_properties=osp, fusion_func=efunc)
    assert not np.isnan(got).any()
    tol = 1e-4 * np.abs(ref) + 1e-6 * np.abs(ref).max()
    bad = np.abs(got - ref) > tol
    assert not bad.any(), (int(bad.sum()), float(np.abs(got - ref).max()))
    assert (ref[12:18, 48:56] > 0).all()  # the hole in view 0 was filled by view 1
