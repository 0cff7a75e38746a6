"""CPU tests of the product's host side: geometry identical to the oracle's
restatement of the reference, the C-ABI library loads and exports every symbol
the header declares, and the product refuses to run without a GPU."""

import ctypes
import os
import re

import numpy as np
import pytest
from scipy import ndimage

import cases
from multiview_stitcher_b200 import _lib, geometry
from oracle import fusion as of

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DIMS = ["z", "y", "x"]


def _ensure_built():
    from multiview_stitcher_b200 import build

    build.build()


def test_library_exports_every_declared_symbol():
    _ensure_built()
    header = open(os.path.join(ROOT, "include", "mvs_b200.h")).read()
    declared = set(re.findall(r"\b(mvs_[a-z0-9_]+)\s*\(", header))
    declared -= {"mvs_status"}
    assert len(declared) >= 20
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [name for name in sorted(declared) if not hasattr(lib, name)]
    assert not missing, f"not exported: {missing}"
    # the Python binding covers the same set
    assert declared == set(_lib.exported_symbols())


def test_struct_layouts_match_the_compiled_library():
    _ensure_built()
    lib = _lib.load()
    a, b = ctypes.c_int(), ctypes.c_int()
    assert lib.mvs_struct_sizes(ctypes.byref(a), ctypes.byref(b)) == 0
    assert (a.value, b.value) == (_lib.VIEW_XFORM_DTYPE.itemsize, _lib.CHUNK_DTYPE.itemsize) == (248, 88)
    assert lib.mvs_abi_version() == 1


def test_no_cpu_fallback_without_device():
    """Without a CUDA device the product raises instead of computing on the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    _ensure_built()
    from multiview_stitcher_b200 import fusion, registration

    case = cases.fusion_cases()["2d_u16_pair_lin"]
    with pytest.raises((_lib.EngineUnavailable, RuntimeError, AssertionError)):
        fusion.fuse(case["views"], case["params"])
    f, m, _ = cases.registration_cases()["blocks_100"]
    with pytest.raises((_lib.EngineUnavailable, RuntimeError, AssertionError)):
        registration.phase_correlation_registration(f, m)


def test_argument_validation_without_gpu():
    """C-ABI argument checks run before any device work."""
    _ensure_built()
    lib = _lib.load()
    plan = ctypes.c_void_p()
    rc = lib.mvs_fuse_plan_create(ctypes.byref(plan), None, 0, None, 0, None, 0, 4, 1, 0, None)
    assert rc == -1 and b"ndim" in lib.mvs_last_error()
    rc = lib.mvs_fuse_plan_create(ctypes.byref(plan), None, 0, None, 0, None, 0, 2, 3, 0, None)
    assert rc == -3 and b"order" in lib.mvs_last_error()
    shp = (ctypes.c_int32 * 3)(1, 16, 16)
    pc = ctypes.c_void_p()
    assert lib.mvs_pc_plan_create(ctypes.byref(pc), 5, shp, 1, 10) == -1
    assert lib.mvs_pc_plan_create(ctypes.byref(pc), 2, shp, 0, 10) == -1


@pytest.mark.parametrize("name", sorted(cases.fusion_cases().keys()))
def test_pixel_affines_equal_the_oracle(name):
    """geometry.pixel_affine reproduces transformation.py:31-83 (via the oracle,
    itself bit-identical to the reference fixtures) for views and weight tables."""
    case = cases.fusion_cases()[name]
    views, params = case["views"], case["params"]
    ndim = views[0]["data"].ndim
    dims = DIMS[-ndim:]
    bbs = [of.view_bb(v) for v in views]
    osp = of.calc_stack_properties(bbs, params, views[0]["spacing"])
    assert geometry.union_stack_props(bbs, params, views[0]["spacing"]) == osp
    o_org, o_sp, _ = geometry.bb_arrays(osp, dims)
    for v, p, bb in zip(views, params, bbs):
        inv = np.linalg.inv(p)
        m_ref, off_ref = of.pixel_affine(inv, osp, v["origin"], v["spacing"])
        m, off = geometry.pixel_affine(inv, o_org, o_sp, [v["origin"][d] for d in dims], [v["spacing"][d] for d in dims])
        assert np.array_equal(m, m_ref) and np.array_equal(off, off_ref)
        # blending support table: closed form == scipy's EDT on the 5^ndim mask
        t_ref, org_ref, sp_ref = of.blending_support(bb, case["kwargs"].get("blending_widths"))
        t, org, sp = geometry.blending_table(bb, case["kwargs"].get("blending_widths"))
        assert np.array_equal(t, t_ref)
        assert np.array_equal(org, [org_ref[d] for d in dims]) and np.array_equal(sp, [sp_ref[d] for d in dims])


@pytest.mark.parametrize("shape,spacing,widths", [
    ((2048, 2048), (1.0, 1.0), None),
    ((256, 512, 512), (2.0, 0.5, 0.5), {"z": 3, "y": 10, "x": 10}),
    ((37, 211), (0.3, 1.7), {"y": 4.5, "x": 0.9}),
    ((9, 100, 31), (1.0, 1.0, 1.0), {"z": 100.0, "y": 1.0, "x": 7.0}),
])
def test_blending_table_closed_form_equals_edt(shape, spacing, widths):
    dims = DIMS[-len(shape):]
    bb = {"origin": dict(zip(dims, [0.0] * len(shape))), "spacing": dict(zip(dims, spacing)), "shape": dict(zip(dims, shape))}
    t_ref, _, _ = of.blending_support(bb, widths)
    t, _, _ = geometry.blending_table(bb, widths)
    assert np.array_equal(t, t_ref)
    # shrink_distance variant (weights.py:348-388)
    t_ref, o_ref, _ = of.blending_support(bb, widths, shrink_distance=1.5)
    t, o, _ = geometry.blending_table(bb, widths, shrink_distance=1.5)
    assert np.array_equal(t, t_ref) and np.array_equal(o, [o_ref[d] for d in dims])


def test_chunk_grid_matches_oracle_chunks():
    bb = {"origin": {"y": 3.0, "x": -2.0}, "spacing": {"y": 0.5, "x": 2.0}, "shape": {"y": 101, "x": 64}}
    cs = {"y": 32, "x": 48}
    got = geometry.chunk_grid(bb, cs)
    ref = of.chunk_bbs(bb, cs)
    assert [g[0] for g in got] == [r[1] for r in ref]
    assert [g[1] for g in got] == [tuple(r[0]["shape"][d] for d in "yx") for r in ref]


def test_gaussian_kernel_is_scipys():
    from multiview_stitcher_b200 import hooks

    for sigma in (1.0, 2.0, 5.0, 11.0):
        w, r = hooks.gaussian_kernel1d(sigma)
        delta = np.zeros(2 * r + 1)
        delta[r] = 1.0
        full = ndimage.gaussian_filter1d(delta, sigma, mode="constant")
        assert r == int(4 * sigma + 0.5)
        assert np.allclose(w, full[r:], rtol=0, atol=1e-17)


def test_candidate_expansion_matches_oracle():
    from multiview_stitcher_b200 import registration as reg
    from oracle import registration as oreg

    rng = np.random.default_rng(0)
    for shape in [(64, 77), (12, 40, 33)]:
        for _ in range(5):
            sc = [rng.uniform(-5, 5, len(shape)).astype(np.float32), np.zeros(len(shape), np.float32),
                  np.array([0.0] + list(rng.uniform(-3, 3, len(shape) - 1)), dtype=np.float32)]
            a = reg._expand_candidates(sc, shape, max(shape))
            b = oreg.expand_candidates(sc, shape, max(shape))
            assert len(a) == len(b) and all(np.array_equal(x, y) for x, y in zip(a, b))


# ---- registration host glue (closed-form candidate statistics, candidate expansion) ----


def test_valid_range_matches_scipy_outside_predicate():
    """0 <= x + t <= n - 1 evaluated in float64, the predicate of
    scipy.ndimage.affine_transform(mode="constant") that decides where the shifted
    moving image is NaN (registration.py:494-505)."""
    from multiview_stitcher_b200.registration import _valid_range

    rng = np.random.default_rng(0)
    for _ in range(500):
        n = int(rng.integers(1, 60))
        t = np.array([rng.uniform(-70, 70), float(rng.integers(-70, 70)), float(np.float32(rng.uniform(-3, 3))), 0.0])
        lo, hi = _valid_range(n, t)
        for tt, l, h in zip(t, lo, hi):
            x = np.arange(n, dtype=np.float64) + tt
            v = ~((x < 0) | (x > n - 1))
            if v.any():
                assert l == np.argmax(v) and h == n - 1 - np.argmax(v[::-1]) and v.sum() == h - l + 1
            else:
                assert h < l


def test_box_candidate_stats_matches_affine_transform():
    """For NaN-free pairs the candidate statistics (mask count, bbox of the valid
    region of im1t) follow from the shift alone; check against scipy on a real image."""
    from scipy import ndimage

    from multiview_stitcher_b200.registration import _box_candidate_stats
    from oracle import registration as oreg

    rng = np.random.default_rng(1)
    im = rng.random((23, 31)).astype(np.float32)
    ts = [[0.0, 0.0], [1.5, -2.25], [-22.0, 3.0], [40.0, 0.0], [-0.1, 30.0], [22.0, -30.0]]
    got = _box_candidate_stats(im.shape, ts)
    for t, g in zip(ts, got):
        im1t = ndimage.affine_transform(im, oreg.affine_from_translation(list(t)), order=1, mode="constant", cval=np.nan)
        valid = ~np.isnan(im1t)
        assert g[0] == valid.sum() and g[1] == valid.sum()
        if valid.any():
            bb = oreg.get_bb_from_nanmask(valid)
            assert list(g[3:5]) == [b[0] for b in bb] and list(g[6:8]) == [b[1] for b in bb]


def test_expand_candidates_matches_oracle():
    """registration.py:461-477 in float32 arithmetic, np.ndindex order."""
    from multiview_stitcher_b200.registration import _expand_candidates
    from oracle import registration as oreg

    rng = np.random.default_rng(2)
    for _ in range(50):
        ndim = int(rng.integers(2, 4))
        shape = tuple(int(x) for x in rng.integers(5, 400, ndim))
        scs = [np.round(rng.uniform(-6, 6, ndim) * 10).astype(np.float32) / np.float32(10) for _ in range(2)]
        scs[1][0] = 0
        a = oreg.expand_candidates(scs, shape, max(shape))
        b = _expand_candidates(scs, shape, max(shape))
        assert len(a) == len(b)
        for x, y in zip(a, b):
            assert all(float(p) == q for p, q in zip(x, y))


def test_batch_block_geometry_is_the_regular_chunk_grid():
    """Hook C: block ids -> offsets / shapes as dask's ``normalize_chunks`` lays them out
    for ``_fuse_chunk_to_zarr`` (fusion/_core.py:2065-2090): full chunks, remainder last."""
    from multiview_stitcher_b200.batch import block_geometry

    osp = {"origin": {"z": 0.0, "y": 1.0, "x": 2.0}, "spacing": {"z": 2.0, "y": 1.0, "x": 1.0},
           "shape": {"z": 5, "y": 70, "x": 64}}
    cs = {"z": 4, "y": 32, "x": 32}
    table = block_geometry(osp, cs)
    assert sorted(table) == [(a, b, c) for a in range(2) for b in range(3) for c in range(2)]
    lin, start, shape = table[(1, 2, 1)]
    assert start == (4, 64, 32) and shape == (1, 6, 32)
    assert [table[k][0] for k in sorted(table)] == list(range(12))
    covered = np.zeros((5, 70, 64), int)
    for _, st, sh in table.values():
        covered[tuple(slice(a, a + n) for a, n in zip(st, sh))] += 1
    assert (covered == 1).all()


def test_hook_argument_errors_are_raised_before_any_device_work():
    """Error behaviour of hooks A and C that does not need a GPU: unsupported registration
    functions / register kwargs / backends are refused up front, an empty edge list is fine."""
    import functools

    from multiview_stitcher_b200 import pairs, registration
    from multiview_stitcher_b200._lib import EngineError
    from multiview_stitcher_b200.batch import BatchFuser

    assert pairs.pairwise_executor([], [], {"transform_key": "stage"}) == []
    assert registration.pairwise_executor([], [], {"transform_key": "stage"}) == []
    with pytest.raises(EngineError, match="phase_correlation_registration only"):
        pairs.pairwise_executor([], [(0, 1)], {"transform_key": "k", "pairwise_reg_func": lambda fixed_data, moving_data: None})
    with pytest.raises(EngineError, match="reg_res_level"):
        pairs.pairwise_executor([], [(0, 1)], {"transform_key": "k", "reg_res_level": 2})
    with pytest.raises(EngineError, match="unsupported register kwargs"):
        pairs.pairwise_executor([], [(0, 1)], {"transform_key": "k", "made_up_option": 1})
    with pytest.raises(KeyError):
        pairs.pairwise_executor([], [(0, 1)], {})  # transform_key is mandatory, like in register()

    osp = {"origin": {"y": 0.0, "x": 0.0}, "spacing": {"y": 1.0, "x": 1.0}, "shape": {"y": 8, "x": 8}}
    part = functools.partial(lambda block_id, **kw: None, output_stack_properties=osp, ns_shape={}, nsdims=[],
                             fuse_kwargs={"images": [], "transform_key": "k", "backend": "jax"},
                             output_chunksize={"y": 8, "x": 8}, output_zarr_array=np.zeros((8, 8)))
    # backend="numpy" / "cupy" are accepted (host arrays are staged, device arrays used in place)
    with pytest.raises(EngineError, match="unknown backend"):
        BatchFuser()(part, [(0, 0)])


def test_pitched_decomposition_of_array_windows():
    """_lib._pitched: how a numpy / tensor window maps onto the staged 3-D copies
    (planes x rows x contiguous width, outer loop over the remaining axes)."""
    from multiview_stitcher_b200 import _lib

    a = np.zeros((5, 7, 11, 13), dtype=np.uint16)
    outer, planes, rows, width, pitch, plane = _lib._pitched(a.shape, a.strides, a.itemsize)
    assert (len(outer), planes, rows, width, pitch, plane) == (5, 7, 11, 26, 26, 11 * 26)
    w = a[1:4, 2:6, 3:9, 4:12]  # a window: pitches stay those of the parent
    outer, planes, rows, width, pitch, plane = _lib._pitched(w.shape, w.strides, w.itemsize)
    assert (len(outer), planes, rows, width, pitch, plane) == (3, 4, 6, 16, 26, 11 * 26)
    b = np.zeros((9, 20), dtype=np.float32)[:, 3:10]
    assert _lib._pitched(b.shape, b.strides, 4) == ([()], 1, 9, 28, 80, 720)
    c = np.zeros(17, dtype=np.uint8)
    assert _lib._pitched(c.shape, c.strides, 1) == ([()], 1, 1, 17, 17, 17)
    with pytest.raises(_lib.EngineError):
        _lib._pitched((4, 6), (4, 24), 4)  # non-contiguous last axis
