"""Output / input side on the device: chunk pack / unpack kernels, the
device -> chunk-file store and its inverse, the OME-Zarr writer (pyramid + metadata) and hook C
writing into the engine's Zarr array -- all against oracle/ngff.py, which is bit-identical to
the reference's own VirtualOMEZarr encoding (tests/test_oracle_ngff.py).  Byte-exact."""
import ctypes
import functools
import json
import os

import numpy as np
import pytest

import cases
from oracle import fusion as of
from oracle import ngff as ongff

pytestmark = pytest.mark.gpu


def _pack(t, chunks):
    import torch

    from multiview_stitcher_b200 import _lib, ngff_io

    lib = _lib.load(require_device=True)
    grid = [-(-n // c) for n, c in zip(t.shape, chunks)]
    nbytes = int(np.prod(grid)) * int(np.prod(chunks)) * t.element_size()
    packed = torch.full((nbytes,), 0xAB, dtype=torch.uint8, device="cuda")
    sh, st, ch = ngff_io._triples(tuple(t.shape), tuple(t.stride()), chunks)
    _lib.check(lib.mvs_chunks_pack(ctypes.c_void_p(t.data_ptr()), t.element_size(), sh, st, ch,
                                   ctypes.c_void_p(packed.data_ptr()), _lib.current_stream_ptr()), "pack")
    return packed


@pytest.mark.parametrize("shape,chunks,dtype,window", [
    ((37, 50), (16, 64), np.float32, None),          # chunk wider than the array
    ((9, 21, 30), (4, 8, 16), np.uint16, None),      # 16-byte rows (vector path), ragged edges
    ((9, 21, 30), (4, 8, 7), np.uint16, None),       # odd chunk rows (element path)
    ((5, 33, 48), (5, 16, 16), np.uint8, None),
    ((40, 64, 96), (16, 32, 32), np.float32, (slice(3, 35), slice(8, 60), slice(16, 80))),  # strided window
    ((64, 70), (32, 32), np.float32, (slice(1, 60), slice(3, 67))),                          # unaligned window
])
def test_pack_unpack_equal_oracle(shape, chunks, dtype, window):
    import torch

    from multiview_stitcher_b200 import _lib, ngff_io

    rng = np.random.default_rng(1)
    full = (rng.random(shape) * 250).astype(dtype)
    t = torch.from_numpy(full).cuda()
    if window is not None:
        t, full = t[window], full[window]
    packed = _pack(t, chunks).cpu().numpy().tobytes()
    enc = ongff.encode_array(np.ascontiguousarray(full), chunks)
    expect = b"".join(enc[k] for k in sorted(enc, key=lambda s: tuple(int(i) for i in s.split("/"))))
    assert packed == expect
    # inverse
    lib = _lib.load(require_device=True)
    back = torch.zeros_like(t)
    sh, st, ch = ngff_io._triples(tuple(back.shape), tuple(back.stride()), chunks)
    pk = torch.from_numpy(np.frombuffer(expect, dtype=np.uint8).copy()).cuda()
    _lib.check(lib.mvs_chunks_unpack(ctypes.c_void_p(pk.data_ptr()), back.element_size(), sh, st, ch,
                                     ctypes.c_void_p(back.data_ptr()), _lib.current_stream_ptr()), "unpack")
    assert np.array_equal(back.cpu().numpy(), full)


@pytest.mark.parametrize("compressor", [None, {"id": "zlib", "level": 1}])
def test_device_store_and_load(tmp_path, compressor):
    import torch

    from multiview_stitcher_b200 import ngff_io

    rng = np.random.default_rng(2)
    data = rng.integers(0, 65535, (2, 70, 100, 130)).astype(np.uint16)
    chunks = (1, 32, 32, 64)
    arr = ngff_io.ZarrArray.create(tmp_path / "s", data.shape, chunks, data.dtype, compressor=compressor)
    t = torch.from_numpy(data).cuda()
    arr.write_device(t[0], lead=(0,))
    # second channel in two chunk-aligned boxes (what hook C does block by block)
    arr.write_device(t[1][:64], lead=(1,), start=(0, 0, 0))
    arr.write_device(t[1][64:], lead=(1,), start=(64, 0, 0))
    expect = ongff.encode_array(data, chunks)
    dec = ngff_io._codec(compressor)
    assert len(expect) == 2 * 3 * 4 * 3
    for key, raw in expect.items():
        with open(tmp_path / "s" / key, "rb") as f:
            got = f.read()
        assert (dec[1](got) if dec else got) == raw, key
    again = ngff_io.ZarrArray.open(tmp_path / "s")
    assert np.array_equal(again.read_device((1,)).cpu().numpy(), data[1])
    assert np.array_equal(again.read_device((0,), start=(32, 64, 64), shape=(38, 36, 66)).cpu().numpy(), data[0, 32:, 64:, 64:])
    assert np.array_equal(again[0, 5:40, 7, 3:99], data[0, 5:40, 7, 3:99])  # host reader sees the same store
    os.remove(again.chunk_path((0, 1, 1, 1)))  # a missing chunk reads as the fill value
    holed = data[0].copy()
    holed[32:64, 32:64, 64:128] = 0
    assert np.array_equal(again.read_device((0,)).cpu().numpy(), holed)
    from multiview_stitcher_b200._lib import EngineError

    with pytest.raises(EngineError):
        again.write_device(t[0][:, 3:], lead=(0,), start=(0, 3, 0))  # not chunk aligned


def _files(root):
    out = {}
    for d, _, fs in os.walk(root):
        for f in fs:
            p = os.path.join(d, f)
            out[os.path.relpath(p, root)] = p
    return out


def test_write_sim_to_ome_zarr_equals_oracle(tmp_path):
    import torch

    from multiview_stitcher_b200 import ngff_io

    rng = np.random.default_rng(3)
    data = rng.integers(0, 4000, (2, 40, 90, 120)).astype(np.uint16)  # (c, z, y, x)
    origin, spacing = {"z": -3.0, "y": 10.5, "x": 2.25}, {"z": 2.0, "y": 0.5, "x": 0.25}
    chunks = {"z": 16, "y": 32, "x": 64}
    img = {"data": torch.from_numpy(data).cuda(), "dims": ["c", "z", "y", "x"], "origin": origin, "spacing": spacing,
           "c_coords": ["dapi", "gfp"]}
    res = ngff_io.write_sim_to_ome_zarr(img, tmp_path / "o.zarr", overwrite=True, chunks=chunks, min_shape=20,
                                        zarr_array_creation_kwargs={"compressor": None})
    store = ongff.write_sim_to_ome_zarr(data, ["c", "z", "y", "x"], origin, spacing, chunks, c_coords=["dapi", "gfp"], min_shape=20)
    files = _files(tmp_path / "o.zarr")
    assert set(files) == set(store)
    assert len(res["arrays"]) == 3 and any(k.startswith("2/") for k in store)
    for key, val in store.items():
        with open(files[key], "rb") as f:
            got = f.read()
        if isinstance(val, bytes):
            assert got == val, key
        else:
            assert json.loads(got) == val, key
    # reading a level back: data + NGFF placement
    dv = ngff_io.read_sim_from_ome_zarr(tmp_path / "o.zarr", resolution_level=1, lead=(1,))
    from oracle import pyramid as opyr

    lvl1 = opyr.build_pyramid({"data": data[1], "origin": origin, "spacing": spacing}, None, 20)[1]
    assert np.array_equal(dv.tensor.cpu().numpy(), lvl1["data"])
    assert dv.spacing == lvl1["spacing"] and dv.origin == lvl1["origin"]
    # overwrite=False keeps the existing level 0 and completes the rest from it (fusion/_core.py:1160-1168)
    import shutil

    shutil.rmtree(tmp_path / "o.zarr" / "1")
    shutil.rmtree(tmp_path / "o.zarr" / "2")
    blank = dict(img, data=torch.zeros_like(img["data"]))
    ngff_io.write_sim_to_ome_zarr(blank, tmp_path / "o.zarr", overwrite=False, chunks=chunks, min_shape=20)
    for key, val in store.items():
        if isinstance(val, bytes):
            with open(_files(tmp_path / "o.zarr")[key], "rb") as f:
                assert f.read() == val, key


def test_fuse_to_ome_zarr_and_hook_c_into_engine_array(tmp_path):
    """fuse(output_zarr_url=..., zarr_options={"ome_zarr": True}) and hook C with the engine's
    ZarrArray as destination: the chunk files hold the fused stack (device-encoded)."""
    from multiview_stitcher_b200 import fusion, ngff_io
    from multiview_stitcher_b200.batch import BatchFuser, block_geometry

    case = cases.fusion_cases()["3d_u16_pair_lin"]
    views, params = case["views"], case["params"]
    kwargs = {k: v for k, v in case["kwargs"].items() if k in ("interpolation_order", "blending_widths")}
    _, osp = of.fuse(views, params, **kwargs)
    dims = ["z", "y", "x"]
    chunksize = {d: max(8, int(osp["shape"][d]) // 2 + 1) for d in dims}
    ref, _ = of.fuse(views, params, output_stack_properties=osp, output_chunksize=chunksize, **kwargs)
    fused, osp2 = fusion.fuse(views, params, output_chunksize=chunksize, output_zarr_url=tmp_path / "f.zarr",
                              zarr_options={"ome_zarr": True}, **kwargs)
    assert np.abs(fused.astype(np.int64) - ref.astype(np.int64)).max() <= 1
    lvl0 = ngff_io.ZarrArray.open(tmp_path / "f.zarr" / "0")
    assert np.array_equal(lvl0[...], fused) and lvl0.chunks == tuple(min(chunksize[d], fused.shape[i]) for i, d in enumerate(dims))
    with open(tmp_path / "f.zarr" / ".zattrs") as f:
        ms = json.load(f)["multiscales"][0]
    assert ms["version"] == "0.4" and [a["name"] for a in ms["axes"]] == dims
    assert ms["datasets"][0]["coordinateTransformations"][1]["translation"] == [osp2["origin"][d] for d in dims]

    # hook C: destination = engine ZarrArray (raw chunks on the output chunk grid)
    dest = ngff_io.ZarrArray.create(tmp_path / "c.zarr", ref.shape, [chunksize[d] for d in dims], ref.dtype)
    msims = [dict(v, transforms={"reg": p}) for v, p in zip(views, params)]
    fk = {"images": msims, "transform_key": "reg", "fusion_func": None, "weights_func": None, "backend": None,
          "output_chunksize": chunksize, **kwargs}

    def never(block_id, **kw):
        raise AssertionError("fuse_chunk called")

    fuse_chunk = functools.partial(never, output_stack_properties=osp, ns_shape={}, nsdims=[], fuse_kwargs=fk,
                                   output_chunksize=chunksize, output_zarr_array=dest)
    ids = sorted(block_geometry(osp, chunksize))
    bf = BatchFuser()
    for i in range(0, len(ids), 3):
        bf(fuse_chunk, ids[i:i + 3])
    assert bf.blocks_written == len(ids) and dest.bytes_written == len(ids) * dest.chunk_bytes
    assert np.array_equal(ngff_io.ZarrArray.open(tmp_path / "c.zarr")[...], fused)


def test_hook_c_content_weighted_into_engine_array(tmp_path):
    """The multi-pass path of hook C (a weights_func: one output stack per block, blocks alternating
    between two streams) writes the engine's ZarrArray from the device and equals the numpy
    destination of the same call sequence."""
    from multiview_stitcher_b200 import fusion as efusion, ngff_io
    from multiview_stitcher_b200.batch import BatchFuser, block_geometry

    rng = np.random.default_rng(10)
    base = cases._smooth(rng, (90, 150), 1.2).astype(np.float32)
    views = [cases._view(base[:, :90].copy(), (0, 0), (1, 1)), cases._view(base[:, 60:].copy(), (0, 0), (1, 1))]
    params = [cases._translation((0, 0)), cases._translation((0.3, 60.2))]
    wkw = {"sigma_1": 2, "sigma_2": 3}
    osp = of.calc_stack_properties([of.view_bb(v) for v in views], params, views[0]["spacing"])
    chunksize = {"y": 48, "x": 64}
    ref, _ = of.fuse(views, params, output_stack_properties=osp, output_chunksize=chunksize,
                     weights_func=of.content_based, weights_func_kwargs=wkw)
    msims = [dict(v, transforms={"reg": p}) for v, p in zip(views, params)]
    ids = sorted(block_geometry(osp, chunksize))

    def run(dest):
        fk = {"images": msims, "transform_key": "reg", "fusion_func": efusion.weighted_average_fusion,
              "weights_func": efusion.content_based, "weights_func_kwargs": wkw, "interpolation_order": 1,
              "blending_widths": None, "backend": None, "output_chunksize": chunksize}

        def never(block_id, **kw):
            raise AssertionError("fuse_chunk called")

        part = functools.partial(never, output_stack_properties=osp, ns_shape={}, nsdims=[], fuse_kwargs=fk,
                                 output_chunksize=chunksize, output_zarr_array=dest)
        bf = BatchFuser()
        for i in range(0, len(ids), 4):
            bf(part, ids[i:i + 4])
        assert bf.blocks_written == len(ids)

    host = np.zeros(ref.shape, np.float32)
    run(host)
    tol = 1e-4 * np.abs(ref) + 1e-6 * np.abs(ref).max()
    assert np.all(np.abs(host - ref) <= tol)
    dest = ngff_io.ZarrArray.create(tmp_path / "cw.zarr", ref.shape, [chunksize[d] for d in "yx"], np.float32)
    run(dest)
    assert dest.bytes_written == len(ids) * dest.chunk_bytes
    assert np.array_equal(ngff_io.ZarrArray.open(tmp_path / "cw.zarr")[...], host)
