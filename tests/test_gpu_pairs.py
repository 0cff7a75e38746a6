"""Pair preparation + batched registration on the GPU (multiview_stitcher_b200.pairs,
SURVEY.md 8f-1 / hook A) against fixtures produced by the reference's own
``_get_overlap_bboxes`` / ``sims_to_intrinsic_coord_system`` /
``phase_correlation_registration`` / ``get_affine_from_intrinsic_affine`` and against
the oracle on a tile grid.

Bars: crop geometry (boxes, grid, NaN mask) exact; crop values exact where the sample
positions are integers, else float32 interpolation vs scipy's float64 taps (<= 2e-6 of
the value range); shifts within 0.1 px (north_star); physical transform / quality equal
once the same 1/u-px bin is chosen."""

import warnings

import numpy as np
import pytest

import cases
from oracle import pairs as opairs

pytestmark = pytest.mark.gpu

CASES = cases.pair_cases(extra=True)
TOL_PX = 0.1 + 1e-4


@pytest.fixture(scope="module")
def gold():
    import os

    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pairs_golden.npz"))


@pytest.fixture(scope="module")
def epairs():
    from multiview_stitcher_b200 import pairs

    return pairs


def _close_crop(got, ref):
    assert got.shape == ref.shape and got.dtype == np.float32
    np.testing.assert_array_equal(np.isnan(got), np.isnan(ref))
    scale = float(np.nanmax(np.abs(ref)))
    np.testing.assert_allclose(np.nan_to_num(got), np.nan_to_num(ref), rtol=0, atol=2e-6 * scale)


@pytest.mark.parametrize("name", sorted(CASES))
def test_prepare_pairs_matches_reference(epairs, gold, name):
    c = CASES[name]
    prep = epairs.prepare_pairs(c["views"], c["affines"], [(0, 1)], **c["kwargs"])
    np.testing.assert_array_equal(np.array(prep.lowers[0]), gold[name + "/lowers"])
    np.testing.assert_array_equal(np.array(prep.uppers[0]), gold[name + "/uppers"])
    np.testing.assert_array_equal(prep.grid[0]["origin"], gold[name + "/grid_origin"])
    f, m = prep.fixed[0].cpu().numpy(), prep.moving[0].cpu().numpy()
    _close_crop(f, gold[name + "/fixed"])
    _close_crop(m, gold[name + "/moving"])
    if name in ("grid2d_x_u16", "grid3d_x_u16_tol", "grid2d_binned_u16"):  # integer sample positions
        np.testing.assert_array_equal(f, gold[name + "/fixed"])
        np.testing.assert_array_equal(m, gold[name + "/moving"])


@pytest.mark.parametrize("name", sorted(CASES))
def test_register_views_matches_reference(epairs, gold, name):
    c = CASES[name]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = epairs.register_views(c["views"], c["affines"], [(0, 1)], **c["kwargs"])[0]
    ndim = c["views"][0]["data"].ndim
    ref_a = gold[name + "/affine_matrix"]
    assert np.max(np.abs(res["affine_matrix"][:ndim, ndim] - ref_a[:ndim, ndim])) <= TOL_PX
    np.testing.assert_array_equal(res["affine_matrix"][:ndim, :ndim], np.eye(ndim))
    np.testing.assert_array_equal(res["bbox"], gold[name + "/bbox"])
    if np.allclose(res["affine_matrix"], ref_a, rtol=0, atol=1e-6):
        np.testing.assert_allclose(res["transform"], gold[name + "/transform"], rtol=0, atol=1e-5)
        assert abs(float(res["quality"]) - float(gold[name + "/quality"])) < 1e-4


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.float32])
@pytest.mark.parametrize("shape,b", [((37, 50), (2, 3)), ((9, 20, 33), (2, 2, 2)), ((8, 31, 64), (1, 2, 4))])
def test_bin_mean(epairs, dtype, shape, b):
    import torch

    from multiview_stitcher_b200.fusion import DeviceView

    rng = np.random.default_rng(5)
    if dtype == np.float32:
        a = rng.random(shape).astype(np.float32)
        a[rng.random(shape) < 0.05] = np.nan
    else:
        a = rng.integers(0, np.iinfo(dtype).max, shape, endpoint=True).astype(dtype)
    dims = opairs.SPATIAL_DIMS[-len(shape):]
    view = {"data": a, "origin": dict(zip(dims, [0.0] * len(shape))), "spacing": dict(zip(dims, [1.0] * len(shape)))}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = opairs.bin_view(opairs.with_coords(view), dict(zip(dims, b)))["data"]
    t = torch.from_numpy(a).cuda()
    # also through a strided window of a larger tensor
    big = torch.zeros(tuple(s + 3 for s in shape), dtype=t.dtype, device="cuda")
    sl = tuple(slice(2, 2 + s) for s in shape)
    big[sl] = t
    for src in (t, big[sl]):
        got = epairs.bin_view(DeviceView(src, view["origin"], view["spacing"]), dict(zip(dims, b))).cpu().numpy()
        assert got.shape == want.shape and got.dtype == want.dtype
        if dtype == np.float32:
            np.testing.assert_allclose(got, want, rtol=1e-6, atol=0, equal_nan=True)
        else:
            np.testing.assert_array_equal(got, want)


def _grid_dataset(ny, nx, tile, ov, seed, dtype=np.uint16):
    """Tiles cut from one smooth ground truth at stage position + hidden jitter."""
    from scipy import ndimage

    rng = np.random.default_rng(seed)
    step = tile - ov
    H, W = step * (ny - 1) + tile + 8, step * (nx - 1) + tile + 8
    gt = ndimage.gaussian_filter(rng.random((H, W)), 1.5)
    gt = (gt - gt.min()) / (gt.max() - gt.min())
    views, affines, jit = [], [], []
    for iy in range(ny):
        for ix in range(nx):
            j = rng.integers(-2, 3, 2)
            y0, x0 = 4 + iy * step + j[0], 4 + ix * step + j[1]
            t = gt[y0:y0 + tile, x0:x0 + tile]
            t = np.round(t * 4000).astype(np.uint16) if dtype == np.uint16 else t.astype(np.float32)
            views.append({"data": t, "origin": {"y": 0.0, "x": 0.0}, "spacing": {"y": 1.0, "x": 1.0}})
            a = np.eye(3)
            a[:2, 2] = (iy * step, ix * step)
            affines.append(a)
            jit.append(j)
    pairs = []
    for iy in range(ny):
        for ix in range(nx):
            k = iy * nx + ix
            if ix + 1 < nx:
                pairs.append((k, k + 1))
            if iy + 1 < ny:
                pairs.append((k, k + nx))
    return views, affines, pairs, jit


def test_grid_all_pairs_batched_vs_oracle(epairs):
    """A 2 x 3 grid: 7 pairs in two crop-shape groups, one resample launch per group; every
    pair agrees with the oracle's register_pair and with the hidden jitter."""
    views, affines, pairs, jit = _grid_dataset(2, 3, 160, 32, seed=3)
    binning = {"y": 1, "x": 1}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res, prep = epairs.register_views(views, affines, pairs, registration_binning=binning, return_prepared=True)
    assert prep.launches == 2
    for k, (i, j) in enumerate(pairs):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref = opairs.register_pair(views[i], views[j], affines[i], affines[j], registration_binning=binning)
        np.testing.assert_array_equal(prep.fixed[k].cpu().numpy(), ref["prepared"]["fixed"])
        np.testing.assert_array_equal(prep.moving[k].cpu().numpy(), ref["prepared"]["moving"])
        assert np.max(np.abs(res[k]["transform"][:2, 2] - ref["transform"][:2, 2])) <= TOL_PX
        np.testing.assert_array_equal(res[k]["bbox"], ref["bbox"])
        # moving content sits at stage + jitter: fixed world -> moving world shift = jit_i - jit_j
        want = np.asarray(jit[i] - jit[j], dtype=float)
        assert np.max(np.abs(res[k]["transform"][:2, 2] - want)) <= 0.35, (k, res[k]["transform"][:2, 2], want)


def test_pairwise_executor_hook(epairs):
    """Hook A (registration.py:2649-2655) on dict 'msims' carrying their transforms."""
    views, affines, pairs, _ = _grid_dataset(1, 3, 128, 30, seed=4, dtype=np.float32)
    msims = [dict(v, transforms={"stage": a}) for v, a in zip(views, affines)]
    kw = {"transform_key": "stage", "registration_binning": {"y": 1, "x": 1}, "overlap_tolerance": None,
          "pairwise_reg_func_kwargs": None}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = epairs.pairwise_executor(msims, pairs, kw)
        direct = epairs.register_views(views, affines, pairs, registration_binning={"y": 1, "x": 1})
    assert len(out) == len(pairs)
    for o, d in zip(out, direct):
        assert np.asarray(o["transform"]).shape == (1, 3, 3)
        assert np.asarray(o["quality"]).shape == (1,)
        assert np.asarray(o["bbox"]).shape == (1, 2, 2)
        np.testing.assert_array_equal(np.asarray(o["transform"])[0], d["transform"])
    from multiview_stitcher_b200._lib import EngineError

    with pytest.raises(EngineError):
        epairs.pairwise_executor(msims, pairs, dict(kw, pairwise_reg_func=lambda fixed_data, moving_data: None))


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.float32])
def test_synthetic_kernel_equals_host_mirror(dtype):
    """csrc/synth.cu and its numpy mirror generate the same tiles (negative origins too), so
    the oracle can be run on the bench's inputs without a GPU."""
    from multiview_stitcher_b200 import synthetic

    for shape, origin in (((5, 33, 70), (-7, 1000, -20)), ((64, 96), (123456, -5))):
        got = synthetic.make_tile(shape, origin, dtype, seed=4).cpu().numpy()
        np.testing.assert_array_equal(got, synthetic.ground_truth(shape, origin, dtype, seed=4))


def test_register_views_sharded_single_rank(epairs):
    from multiview_stitcher_b200 import distributed

    views, affines, pairs, _ = _grid_dataset(1, 3, 128, 30, seed=6)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = distributed.register_views_sharded(views, affines, pairs, registration_binning={"y": 1, "x": 1})
        b = epairs.register_views(views, affines, pairs, registration_binning={"y": 1, "x": 1})
    for x, y in zip(a, b):
        np.testing.assert_array_equal(x["transform"], y["transform"])


class _XA:
    """xarray.DataArray stand-in: dims / sizes / coords[d].values / data / isel(**kw)."""

    class _C:
        def __init__(self, v):
            self.values = np.asarray(v)

    def __init__(self, data, dims, coords=None):
        self.data, self.dims = data, tuple(dims)
        self.coords = coords or {}
        self.sizes = dict(zip(self.dims, data.shape))

    def isel(self, indexers=None, **kw):
        sel = dict(indexers or {}, **kw)
        idx = tuple(sel.get(d, slice(None)) for d in self.dims)
        dims = [d for d in self.dims if d not in sel]
        return _XA(self.data[idx], dims, {d: c for d, c in self.coords.items() if d in dims})


def test_pairwise_executor_on_timelapse_msims(epairs):
    """msim-like mappings with a "t" axis (two time points, per-time-point transforms):
    one result per edge with a leading t axis, each time point equal to register_views."""
    rng = np.random.default_rng(8)
    per_t = [_grid_dataset(1, 3, 128, 30, seed=10 + t, dtype=np.float32) for t in range(2)]
    pairs = per_t[0][2]
    msims = []
    for v in range(3):
        data = np.stack([per_t[t][0][v]["data"] for t in range(2)])
        aff = np.stack([per_t[t][1][v] for t in range(2)])
        aff[1, 0, 2] += 0.25 * v  # the stage drifts between time points -> a second plan
        coords = {"y": _XA._C(np.arange(128, dtype=float)), "x": _XA._C(np.arange(128, dtype=float))}
        msims.append({"scale0/image": _XA(data, ("t", "y", "x"), coords), "scale0": {"stage": _XA(aff, ("t", "x_in", "x_out"))}})
    kw = {"transform_key": "stage", "registration_binning": {"y": 1, "x": 1}, "overlap_tolerance": None,
          "pairwise_reg_func_kwargs": None, "points_key": "beads", "prefilter_markers": False, "reg_res_level": None}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = epairs.pairwise_executor(msims, pairs, kw)
    assert len(out) == len(pairs)
    for t in range(2):
        views = [{"data": m["scale0/image"].data[t], "origin": {"y": 0.0, "x": 0.0}, "spacing": {"y": 1.0, "x": 1.0}} for m in msims]
        affs = [m["scale0"]["stage"].data[t] for m in msims]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want = epairs.register_views(views, affs, pairs, registration_binning={"y": 1, "x": 1})
        for o, w in zip(out, want):
            assert np.asarray(o["transform"]).shape == (2, 3, 3) and np.asarray(o["bbox"]).shape == (2, 2, 2)
            np.testing.assert_array_equal(np.asarray(o["transform"])[t], w["transform"])
            np.testing.assert_array_equal(np.asarray(o["bbox"])[t], w["bbox"])
            assert float(np.asarray(o["quality"])[t]) == float(w["quality"])
