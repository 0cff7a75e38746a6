"""The Zarr v2 / OME-Zarr oracle (oracle/ngff.py) against fixtures produced by the reference's
own VirtualOMEZarr and calc_ngff_coordinate_transformations_and_axes
(tests/golden/make_golden_ngff.py), and the engine's host-side ZarrArray against the oracle."""
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden_ngff import ngff_cases, transform_cases  # noqa: E402

from oracle import ngff as ongff  # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "ngff_golden.npz"))


@pytest.mark.parametrize("name", sorted(ngff_cases()))
def test_chunks_and_metadata_equal_reference(name):
    case = ngff_cases()[name]
    n_checked = 0
    for lvl, (data, _, _) in enumerate(case["levels"]):
        chunks = tuple(min(c, n) for c, n in zip(case["chunks"], case["levels"][0][0].shape))
        # the virtual store clamps per level (ngff_utils.py:142-185)
        chunks = tuple(min(c, n) for c, n in zip(chunks, data.shape))
        assert ongff.array_zarray(data.shape, chunks, data.dtype) == json.loads(str(GOLD[f"{name}/{lvl}/zarray"]))
        for key, raw in ongff.encode_array(data, chunks).items():
            assert raw == GOLD[f"{name}/{lvl}/chunk/{key}"].tobytes(), (name, lvl, key)
            n_checked += 1
        back = ongff.decode_array(ongff.encode_array(data, chunks), data.shape, chunks, data.dtype)
        assert np.array_equal(back, data)
    assert n_checked == len([k for k in GOLD.files if k.startswith(name + "/") and "/chunk/" in k])
    zattrs = ongff.virtual_root_zattrs([(o, s) for _, o, s in case["levels"]], case["dims"], name="image")
    assert zattrs == json.loads(str(GOLD[f"{name}/zattrs"]))
    assert json.loads(str(GOLD[f"{name}/zgroup"])) == {"zarr_format": 2}


@pytest.mark.parametrize("name", sorted(transform_cases()))
def test_level_transforms_equal_reference(name):
    kw = transform_cases()[name]
    coordtfs, axes = ongff.calc_ngff_coordinate_transformations_and_axes(**kw)
    ref = json.loads(str(GOLD[f"tf/{name}"]))
    assert coordtfs == ref["coordtfs"] and axes == ref["axes"]
    from multiview_stitcher_b200 import ngff_io

    coordtfs2, axes2 = ngff_io.calc_ngff_coordinate_transformations_and_axes(**kw)
    assert coordtfs2 == ref["coordtfs"] and axes2 == ref["axes"]


@pytest.mark.parametrize("compressor", [None, {"id": "zlib", "level": 1}, {"id": "gzip", "level": 1}])
def test_host_zarr_array_matches_oracle(tmp_path, compressor):
    """ZarrArray's numpy-style region writes (what hook C does with the destination array,
    fusion/_core.py:2130-2150) leave the oracle's chunk files; reads give the array back."""
    from multiview_stitcher_b200 import ngff_io

    rng = np.random.default_rng(0)
    data = rng.integers(0, 60000, (2, 9, 21, 30)).astype(np.uint16)
    chunks = (1, 4, 8, 16)
    arr = ngff_io.ZarrArray.create(tmp_path / "a", data.shape, chunks, data.dtype, compressor=compressor)
    # block-wise like _fuse_chunk_to_zarr, plus one unaligned write (read-modify-write)
    for c in range(2):
        for z0 in range(0, 9, 4):
            arr[c : c + 1, z0 : z0 + 4] = data[c : c + 1, z0 : z0 + 4]
    arr[1, 2:7, 3:11, 5:29] = data[1, 2:7, 3:11, 5:29]
    with open(tmp_path / "a" / ".zarray") as f:
        assert json.load(f) == ongff.array_zarray(data.shape, chunks, data.dtype, compressor)
    expect = ongff.encode_array(data, chunks)
    dec = ngff_io._codec(compressor)
    for key, raw in expect.items():
        with open(tmp_path / "a" / key, "rb") as f:
            got = f.read()
        assert (dec[1](got) if dec else got) == raw, key
    again = ngff_io.ZarrArray.open(tmp_path / "a")
    assert np.array_equal(again[...], data)
    assert np.array_equal(again[1, 2:7, :, 5], data[1, 2:7, :, 5])
    assert np.array_equal(np.asarray(again), data)


def test_unsupported_store_features_raise(tmp_path):
    from multiview_stitcher_b200 import ngff_io
    from multiview_stitcher_b200._lib import EngineError

    with pytest.raises(EngineError):
        ngff_io.ZarrArray.create(tmp_path / "b", (4, 4), (2, 2), np.uint8, compressor={"id": "blosc"})
    with pytest.raises(EngineError):
        ngff_io.multiscales_zattrs([], [], ngff_version="0.5")
    with pytest.raises(EngineError):
        ngff_io.ZarrArray.open(tmp_path / "missing")
