"""oracle.pairs against fixtures produced by the reference's own pair-preparation
code (tests/golden/make_golden_pairs.py -> pairs_golden.npz)."""

import os
import sys
import warnings

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

import cases  # noqa: E402
from oracle import pairs as opairs  # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "pairs_golden.npz"))
CASES = cases.pair_cases(extra=True)


@pytest.mark.parametrize("name", sorted(CASES))
def test_prepare_pair_matches_reference(name):
    c = CASES[name]
    prep = opairs.prepare_pair(c["views"][0], c["views"][1], c["affines"][0], c["affines"][1], **c["kwargs"])
    np.testing.assert_array_equal(np.array(prep["lowers"]), GOLD[name + "/lowers"])
    np.testing.assert_array_equal(np.array(prep["uppers"]), GOLD[name + "/uppers"])
    dims = opairs.SPATIAL_DIMS[-c["views"][0]["data"].ndim:]
    np.testing.assert_array_equal([prep["grid"]["origin"][d] for d in dims], GOLD[name + "/grid_origin"])
    np.testing.assert_array_equal([prep["grid"]["spacing"][d] for d in dims], GOLD[name + "/grid_spacing"])
    for k in ("fixed", "moving"):
        assert prep[k].dtype == np.float32
        np.testing.assert_array_equal(prep[k], GOLD[name + "/" + k])  # NaNs compare equal


@pytest.mark.parametrize("name", sorted(CASES))
def test_register_pair_matches_reference(name):
    c = CASES[name]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = opairs.register_pair(c["views"][0], c["views"][1], c["affines"][0], c["affines"][1], **c["kwargs"])
    np.testing.assert_array_equal(res["affine_matrix"], GOLD[name + "/affine_matrix"])
    np.testing.assert_array_equal(res["transform"], GOLD[name + "/transform"])
    np.testing.assert_array_equal(res["bbox"], GOLD[name + "/bbox"])
    assert float(res["quality"]) == float(GOLD[name + "/quality"])


def test_optimal_binning_heuristic():
    # registration.py:114-191: bins the finest-spaced axes until the tile has < 400^3 voxels
    v = lambda shape, sp: opairs.with_coords(
        {"data": np.zeros(shape, np.uint8), "origin": dict(zip("zyx", (0, 0, 0))), "spacing": dict(zip("zyx", sp))})
    assert opairs.optimal_registration_binning(v((256, 512, 512), (1, 1, 1)), v((256, 512, 512), (1, 1, 1))) == {"z": 2, "y": 1, "x": 1}
    assert opairs.optimal_registration_binning(v((256, 512, 512), (2, 1, 1)), v((256, 512, 512), (2, 1, 1))) == {"z": 1, "y": 2, "x": 2}
    assert opairs.optimal_registration_binning(v((100, 100, 100), (1, 1, 1)), v((100, 100, 100), (1, 1, 1))) == {"z": 1, "y": 1, "x": 1}


def test_binning_heuristic_matches_reference_source():
    """Runs the reference's own ``get_optimal_registration_binning`` body (registration.py:114-191,
    extracted by name; only needs shapes and spacings) next to the oracle and the engine."""
    import ast
    import types

    path = "/root/reference/src/multiview_stitcher/registration.py"
    if not os.path.exists(path):
        pytest.skip("reference tree not present")
    tree = ast.parse(open(path).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "get_optimal_registration_binning")
    si = types.SimpleNamespace(
        get_spatial_dims_from_sim=lambda sim: list(sim.dims),
        get_spacing_from_sim=lambda sim, asarray=False: dict(sim.spacing),
    )
    ns = {"np": np, "spatial_image_utils": si}
    exec(compile(ast.Module([fn], []), path, "exec"), ns)
    from multiview_stitcher_b200 import pairs as epairs

    for shape, sp in (((256, 512, 512), (1, 1, 1)), ((256, 512, 512), (2, 1, 1)), ((512, 2048, 2048), (1, 0.5, 0.5)),
                      ((2048, 2048), (1, 1)), ((9000, 9000), (0.5, 0.5)), ((100, 100, 100), (1, 1, 1)),
                      ((800, 400, 400), (0.3, 1, 1))):
        dims = opairs.SPATIAL_DIMS[-len(shape):]
        sim = types.SimpleNamespace(dims=dims, shape=shape, spacing=dict(zip(dims, map(float, sp))))
        want = ns["get_optimal_registration_binning"](sim, sim)
        v = opairs.with_coords({"data": np.broadcast_to(np.zeros((1,) * len(shape), np.uint8), shape),
                                "origin": dict(zip(dims, (0.0,) * len(shape))), "spacing": dict(zip(dims, map(float, sp)))})
        assert opairs.optimal_registration_binning(v, v) == want
        assert epairs.optimal_registration_binning(shape, shape, sp, sp, dims) == want
