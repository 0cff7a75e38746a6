"""Resolution-level arithmetic of the output pyramid: engine mirror == oracle == (when the
reference tree is present) the reference's own ``msi_utils.calc_resolution_levels``."""

import ast
import os

import numpy as np
import pytest

from multiview_stitcher_b200 import pyramid as epyr
from oracle import pyramid as opyr

SHAPES = [{"y": 9012, "x": 9012}, {"z": 490, "y": 1899, "x": 1897}, {"z": 64, "y": 300, "x": 201}, {"y": 100, "x": 100},
          {"z": 516, "y": 3895, "x": 7579}]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("factors", [None, {"z": 1, "y": 2, "x": 2}, {"z": 2, "y": 3, "x": 3}])
def test_levels_match_oracle(shape, factors):
    f = None if factors is None else {d: factors[d] for d in shape}
    assert epyr.calc_resolution_levels(shape, f) == opyr.calc_resolution_levels(shape, f)
    assert epyr.calc_resolution_levels(shape, f, min_shape=30) == opyr.calc_resolution_levels(shape, f, min_shape=30)


def test_oracle_levels_match_reference_source():
    """Executes the reference's own function body (pure dict arithmetic, no imports needed)."""
    path = "/root/reference/src/multiview_stitcher/msi_utils.py"
    if not os.path.exists(path):
        pytest.skip("reference tree not present")
    tree = ast.parse(open(path).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "calc_resolution_levels")
    ns = {}
    exec(compile(ast.Module([fn], []), path, "exec"), ns)
    for shape in SHAPES:
        for f in (None, {d: 2 for d in shape}, {d: (1 if d == "z" else 2) for d in shape}):
            assert ns["calc_resolution_levels"](shape, f) == opyr.calc_resolution_levels(shape, f)


def test_oracle_coarsen_is_windowed_mean_truncated():
    a = np.arange(7 * 10, dtype=np.uint16).reshape(7, 10) * 3
    got = opyr.coarsen(a, [2, 3])
    assert got.shape == (3, 3) and got.dtype == np.uint16
    assert got[1, 2] == int(a[2:4, 6:9].mean())
