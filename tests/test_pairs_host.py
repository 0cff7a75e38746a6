"""Host geometry of pair preparation (multiview_stitcher_b200.pairs) -- no GPU needed.

The engine's plan for every fixture pair must carry exactly the reference's numbers:
overlap boxes, crop windows, the common grid, and pixel matrices / offsets which, fed
to scipy on the cropped data, reproduce the reference's crops bit for bit."""

import os
import sys
import types

import numpy as np
import pytest
from scipy import ndimage

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

import cases  # noqa: E402
from multiview_stitcher_b200 import pairs as epairs  # noqa: E402
from oracle import pairs as opairs  # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "pairs_golden.npz"))
CASES = cases.pair_cases(extra=True)


def _axes(view):
    dims = opairs.SPATIAL_DIMS[-view["data"].ndim:]
    dv = types.SimpleNamespace(dims=dims, origin=view["origin"], spacing=view["spacing"], shape=view["data"].shape)
    return epairs._Axes.of_view(dv)


@pytest.mark.parametrize("name", sorted(CASES))
def test_plan_pair_reproduces_reference_crops(name):
    c = CASES[name]
    dims = opairs.SPATIAL_DIMS[-c["views"][0]["data"].ndim:]
    b = tuple(c["kwargs"]["registration_binning"][d] for d in dims)
    views = [opairs.with_coords(v) for v in c["views"]]
    axes = [_axes(v) for v in c["views"]]
    if max(b) > 1:
        views = [opairs.bin_view(v, dict(zip(dims, b))) for v in views]
        axes = [a.binned(b) for a in axes]
        for a, v in zip(axes, views):
            for ca, d in zip(a.coords, dims):
                np.testing.assert_array_equal(ca, v["coords"][d])
    tol = epairs._tolerance(c["kwargs"].get("overlap_tolerance"), dims)
    pl = epairs.plan_pair(axes[0], axes[1], np.asarray(c["affines"][0], float), np.asarray(c["affines"][1], float), tol)
    np.testing.assert_array_equal(np.array(pl["lowers"]), GOLD[name + "/lowers"])
    np.testing.assert_array_equal(np.array(pl["uppers"]), GOLD[name + "/uppers"])
    np.testing.assert_array_equal(pl["origin"], GOLD[name + "/grid_origin"])
    np.testing.assert_array_equal(pl["spacing"], GOLD[name + "/grid_spacing"])
    assert pl["shape"] == GOLD[name + "/fixed"].shape
    for side, key in ((0, "fixed"), (1, "moving")):
        win = views[side]["data"][tuple(slice(i0, i1) for i0, i1 in pl["ranges"][side])].astype(np.float32)
        m, off = pl["xforms"][side]
        got = ndimage.affine_transform(win, matrix=m, offset=off, output_shape=pl["shape"], mode="constant",
                                       cval=np.nan, order=1)
        np.testing.assert_array_equal(got.astype(np.float32), GOLD[name + "/" + key])


@pytest.mark.parametrize("name", sorted(CASES))
def test_physical_transform_and_bbox(name):
    c = CASES[name]
    dims = opairs.SPATIAL_DIMS[-c["views"][0]["data"].ndim:]
    grid = {"origin": GOLD[name + "/grid_origin"], "spacing": GOLD[name + "/grid_spacing"]}
    t = epairs.physical_transform(GOLD[name + "/affine_matrix"], grid, np.asarray(c["affines"][0], float))
    np.testing.assert_array_equal(t, GOLD[name + "/transform"])
    tol = epairs._tolerance(c["kwargs"].get("overlap_tolerance"), dims)
    lo, hi = epairs.overlap_bboxes(_axes(c["views"][0]), _axes(c["views"][1]), np.asarray(c["affines"][0], float),
                                   np.asarray(c["affines"][1], float), tol, intrinsic=False)
    np.testing.assert_array_equal(np.array([lo[0], hi[0]]), GOLD[name + "/bbox"])


def test_binning_heuristic_matches_oracle():
    for shape, sp in (((256, 512, 512), (1, 1, 1)), ((256, 512, 512), (2, 1, 1)), ((100, 100, 100), (1, 1, 1)),
                      ((2048, 2048), (1, 1)), ((9000, 9000), (0.5, 0.5))):
        dims = opairs.SPATIAL_DIMS[-len(shape):]
        v = opairs.with_coords({"data": np.broadcast_to(np.zeros((1,) * len(shape), np.uint8), shape),
                                "origin": dict(zip(dims, (0,) * len(shape))), "spacing": dict(zip(dims, sp))})
        want = opairs.optimal_registration_binning(v, v)
        got = epairs.optimal_registration_binning(shape, shape, sp, sp, dims)
        assert got == want


def test_disjoint_views_raise():
    from multiview_stitcher_b200._lib import EngineError

    v = {"data": np.zeros((10, 10), np.float32), "origin": {"y": 0.0, "x": 0.0}, "spacing": {"y": 1.0, "x": 1.0}}
    a1, a2 = np.eye(3), np.eye(3)
    a2[:2, 2] = (0, 50)
    with pytest.raises(EngineError):
        epairs.plan_pair(_axes(v), _axes(v), a1, a2, {"y": 0.0, "x": 0.0})


def test_pair_plan_from_xarray_likes_and_view_subset():
    """PairPlan reads only coordinates from xarray-like views (no data access) and a plan
    over a subset of the pairs names exactly the views it needs (sharded runs)."""

    class Coord:
        def __init__(self, v):
            self.values = np.asarray(v)

    class XView:
        def __init__(self, view):
            dims = opairs.SPATIAL_DIMS[-view["data"].ndim:]
            self.dims = tuple(dims)
            self.coords = {d: Coord(view["origin"][d] + view["spacing"][d] * np.arange(n, dtype=float))
                           for d, n in zip(dims, view["data"].shape)}

        @property
        def data(self):
            raise AssertionError("planning must not touch the voxels")

    c = CASES["grid2d_y_f32_subpixel"]
    views3 = [c["views"][0], c["views"][1], c["views"][1]]
    aff3 = [c["affines"][0], c["affines"][1], c["affines"][1]]
    kw = dict(overlap_tolerance=None, registration_binning=c["kwargs"]["registration_binning"])
    p_dict = epairs.PairPlan(views3, aff3, [(0, 1), (0, 2)], **kw)
    p_x = epairs.PairPlan([XView(v) for v in views3], aff3, [(0, 2)], **kw)
    assert p_x.used_views == [0, 2] and p_dict.used_views == [0, 1, 2]
    a, b = p_dict.items[1], p_x.items[0]
    assert a["shape"] == b["shape"] and a["ranges"] == b["ranges"]
    for s in (0, 1):
        np.testing.assert_array_equal(a["xforms"][s][0], b["xforms"][s][0])
        np.testing.assert_array_equal(a["xforms"][s][1], b["xforms"][s][1])
    np.testing.assert_array_equal(p_dict.bbox[0], GOLD["grid2d_y_f32_subpixel/bbox"])


def test_synthetic_host_mirror_overlaps_agree():
    """Tiles of the synthetic ground truth agree exactly where they overlap (the property
    the registration ground truth rests on); values stay below the documented bound."""
    from multiview_stitcher_b200 import synthetic

    a = synthetic.ground_truth((6, 40, 48), (-3, 10, 5), np.uint16, seed=3)
    b = synthetic.ground_truth((6, 40, 48), (-1, 25, 20), np.uint16, seed=3)
    np.testing.assert_array_equal(a[2:, 15:, 15:], b[:4, :25, :33])
    assert a.max() < 5376 and a.std() > 100
    f = synthetic.ground_truth((40, 48), (10, 5), np.float32, seed=3)
    np.testing.assert_array_equal(f, synthetic.ground_truth((1, 40, 48), (0, 10, 5), np.uint16, seed=3)[0].astype(np.float32) / np.float32(8192))


def test_plan_pair_fuzz_vs_oracle():
    """Random pair geometries (2-D / 3-D, anisotropic and unequal spacings, off-grid origins,
    small rotations / shears, tolerance, binning): the engine's plan carries exactly the
    oracle's overlap boxes, grid and crop windows, and -- fed to scipy -- its pixel affines
    reproduce the oracle's crops bit for bit."""
    rng = np.random.default_rng(1234)
    n_done = 0
    for trial in range(60):
        ndim = 2 if trial % 3 else 3
        dims = opairs.SPATIAL_DIMS[-ndim:]
        shape = tuple(int(s) for s in rng.integers(12, 40, ndim))
        sp1 = rng.choice([0.5, 0.65, 1.0, 2.0], ndim)
        sp2 = sp1 if trial % 4 else sp1 * rng.choice([0.5, 1.0, 2.0], ndim)
        views = []
        for sp in (sp1, sp2):
            data = rng.integers(0, 4000, shape).astype(np.uint16 if trial % 2 else np.float32)
            views.append({"data": data, "origin": dict(zip(dims, rng.normal(0, 5, ndim).round(2))),
                          "spacing": dict(zip(dims, map(float, sp)))})
        ext = (np.array(shape) - 1) * sp1
        a1 = np.eye(ndim + 1)
        a2 = np.eye(ndim + 1)
        if trial % 5 == 0:
            a2[:ndim, :ndim] += rng.normal(0, 0.02, (ndim, ndim))
        ax = int(rng.integers(0, ndim))
        a2[ax, ndim] = ext[ax] * rng.uniform(0.55, 0.8)
        a2[:ndim, ndim] += rng.normal(0, 0.7, ndim)
        binning = dict(zip(dims, (2,) * ndim)) if trial % 7 == 0 else dict(zip(dims, (1,) * ndim))
        tolerance = None if trial % 6 else float(rng.uniform(0.5, 3.0))
        try:
            want = opairs.prepare_pair(views[0], views[1], a1, a2, tolerance, binning)
        except Exception:
            continue  # degenerate overlap: the reference's path fails as well
        axes = [_axes(v) for v in views]
        b = tuple(binning[d] for d in dims)
        bviews = [opairs.with_coords(v) for v in views]
        if max(b) > 1:
            axes = [a.binned(b) for a in axes]
            bviews = [opairs.bin_view(v, binning) for v in bviews]
        pl = epairs.plan_pair(axes[0], axes[1], a1, a2, epairs._tolerance(tolerance, dims))
        np.testing.assert_array_equal(np.array(pl["lowers"]), np.array(want["lowers"]))
        np.testing.assert_array_equal(np.array(pl["uppers"]), np.array(want["uppers"]))
        assert pl["shape"] == want["fixed"].shape
        np.testing.assert_array_equal(pl["origin"], [want["grid"]["origin"][d] for d in dims])
        np.testing.assert_array_equal(pl["spacing"], [want["grid"]["spacing"][d] for d in dims])
        for side, key in ((0, "fixed"), (1, "moving")):
            win = bviews[side]["data"][tuple(slice(i0, i1) for i0, i1 in pl["ranges"][side])].astype(np.float32)
            m, off = pl["xforms"][side]
            got = ndimage.affine_transform(win, matrix=m, offset=off, output_shape=pl["shape"], mode="constant",
                                           cval=np.nan, order=1)
            np.testing.assert_array_equal(got.astype(np.float32), want[key])
        n_done += 1
    assert n_done >= 45


def test_pairwise_executor_assigns_time_coordinates(monkeypatch):
    """Hook A's results carry the "t" coordinate VALUES of the input (registration.py:2091):
    the reference's groupwise resolution looks time points up by label
    (param_resolution/utils.py:23-39).  xarray is not installed here, so a minimal stand-in
    records what the executor constructs; the device work is stubbed."""
    import sys
    import types

    from multiview_stitcher_b200 import pairs as P

    class DataArray:
        def __init__(self, data, dims=None, coords=None):
            self.data, self.dims, self.coords = np.asarray(data), tuple(dims), dict(coords or {})

    monkeypatch.setitem(sys.modules, "xarray", types.SimpleNamespace(DataArray=DataArray))

    class Coord:
        def __init__(self, v):
            self.values = np.asarray(v)

    class Sim:
        def __init__(self, nt):
            self.dims = ("t", "y", "x")
            self.sizes = {"t": nt, "y": 8, "x": 8}
            self.coords = {"t": Coord([10.0, 20.5, 31.0][:nt])}

        def isel(self, t=0):
            return {"data": np.zeros((8, 8), np.float32), "origin": {"y": 0.0, "x": 0.0}, "spacing": {"y": 1.0, "x": 1.0}}

    msims = [{"scale0/image": Sim(3), "scale0": {"reg": np.eye(3)}} for _ in range(2)]

    class FakePlan:
        def __init__(self, *a, **k):
            pass

    def fake_register_views(views, plan=None, **kw):
        return [{"transform": np.eye(3), "quality": 0.5, "bbox": np.zeros((2, 2))}]

    monkeypatch.setattr(P, "PairPlan", FakePlan)
    monkeypatch.setattr(P, "register_views", fake_register_views)
    out = P.pairwise_executor(msims, [(0, 1)], {"transform_key": "reg"})
    assert len(out) == 1
    for key, dims in (("transform", ("t", "x_in", "x_out")), ("quality", ("t",)), ("bbox", ("t", "point_index", "dim"))):
        da = out[0][key]
        assert da.dims == dims and da.data.shape[0] == 3
        np.testing.assert_array_equal(da.coords["t"], [10.0, 20.5, 31.0])
    assert out[0]["transform"].coords["x_in"] == ["y", "x", "1"]
