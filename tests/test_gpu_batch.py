"""Hook C (``batch_options={"batch_func": ...}``, fusion/_core.py:1133-1141): the engine's
``BatchFuser`` driven exactly as the reference drives a batch_func -- a
``functools.partial`` carrying ``_fuse_chunk_to_zarr``'s keywords (_core.py:2044-2053) and
lists of block ids -- must fill the destination array with what the oracle's chunked
``fuse`` computes (float32 within 1e-4 relative, uint16 within 1 LSB / exact for
max_fusion)."""

import functools

import numpy as np
import pytest

import cases
from oracle import fusion as of

pytestmark = pytest.mark.gpu


def _never(block_id, **kw):  # the batch_func must not fall back to the per-block CPU path
    raise AssertionError("fuse_chunk called")


class XSim:
    """xarray-like stand-in (dims / coords / data / attrs / isel) for views with t, c."""

    def __init__(self, data, dims, coords, attrs):
        self.data, self.dims, self.coords, self.attrs = data, tuple(dims), coords, attrs

    def isel(self, sel):
        idx = tuple(sel.get(d, slice(None)) for d in self.dims)
        dims = [d for d in self.dims if d not in sel]
        return XSim(self.data[idx], dims, {d: self.coords[d] for d in dims if d in self.coords}, self.attrs)


class XCoord:
    def __init__(self, values):
        self.values = np.asarray(values)


def _partial(views_or_sims, osp, chunksize, out, nsdims=(), ns_shape=None, **fuse_kwargs):
    fk = {"images": views_or_sims, "transform_key": "reg", "fusion_func": None, "weights_func": None,
          "interpolation_order": 1, "blending_widths": None, "backend": None, "output_chunksize": chunksize}
    fk.update(fuse_kwargs)
    return functools.partial(_never, output_stack_properties=osp, ns_shape=ns_shape or {}, nsdims=list(nsdims),
                             fuse_kwargs=fk, output_chunksize=chunksize, output_zarr_array=out)


def _close(got, ref):
    if ref.dtype.kind == "u":
        assert np.abs(got.astype(np.int64) - ref.astype(np.int64)).max() <= 1
    else:
        tol = 1e-4 * np.abs(ref) + 1e-6 * np.abs(ref).max()
        assert np.all(np.abs(got.astype(np.float64) - ref.astype(np.float64)) <= tol)


@pytest.mark.parametrize("name", ["2d_f32_quad_lin", "3d_u16_pair_lin", "2d_f32_affine_lin"])
def test_batches_fill_the_store_like_the_oracle(name):
    from multiview_stitcher_b200.batch import BatchFuser, block_geometry

    case = cases.fusion_cases()[name]
    views, params = case["views"], case["params"]
    ndim = views[0]["data"].ndim
    dims = ["z", "y", "x"][-ndim:]
    kwargs = {k: v for k, v in case["kwargs"].items() if k in ("interpolation_order", "blending_widths")}
    ref_full, osp = of.fuse(views, params, **kwargs)
    chunksize = {d: max(8, int(osp["shape"][d]) // 3 + 1) for d in dims}
    ref, _ = of.fuse(views, params, output_stack_properties=osp, output_chunksize=chunksize, **kwargs)
    msims = [dict(v, transforms={"reg": p}) for v, p in zip(views, params)]
    out = np.zeros(ref.shape, dtype=ref.dtype)
    fuse_chunk = _partial(msims, osp, chunksize, out, **kwargs)
    ids = sorted(block_geometry(osp, chunksize))
    bf = BatchFuser()
    for i in range(0, len(ids), 4):  # n_batch = 4
        bf(fuse_chunk, ids[i:i + 4])
    assert bf.blocks_written == len(ids)
    assert out.dtype == views[0]["data"].dtype
    _close(out, ref)


def test_nonspatial_dims_and_max_fusion():
    """(t, c, y, x) sims with per-time-point transforms; blocks of different (t, c) slices
    arrive in one batch."""
    from multiview_stitcher_b200 import fusion as efusion
    from multiview_stitcher_b200.batch import BatchFuser, block_geometry

    rng = np.random.default_rng(9)
    T, C, H, W = 2, 2, 40, 56
    sims, per_t = [], []
    for v in range(2):
        data = rng.integers(0, 4000, (T, C, H, W)).astype(np.uint16)
        tr = np.stack([np.eye(3)] * T)
        tr[:, :2, 2] = [(0.0, v * 40.0), (1.0, v * 41.0)]
        coords = {"t": XCoord([0, 1]), "c": XCoord(["a", "b"]), "y": XCoord(np.arange(H, dtype=float)),
                  "x": XCoord(np.arange(W, dtype=float))}
        sims.append(XSim(data, ("t", "c", "y", "x"), coords, {"transforms": {"reg": XSim(tr, ("t", "x_in", "x_out"), {}, {})}}))
        per_t.append(tr)
    osp = {"origin": {"y": 0.0, "x": 0.0}, "spacing": {"y": 1.0, "x": 1.0}, "shape": {"y": 42, "x": 100}}
    chunksize = {"y": 32, "x": 32}
    out = np.zeros((T, C, 42, 100), np.uint16)
    fuse_chunk = _partial(sims, osp, chunksize, out, nsdims=("t", "c"), ns_shape={"t": T, "c": C},
                          fusion_func=efusion.max_fusion, interpolation_order=0)
    spatial = sorted(block_geometry(osp, chunksize))
    ids = [(t, c) + s for t in range(T) for c in range(C) for s in spatial]
    bf = BatchFuser()
    for i in range(0, len(ids), 5):
        bf(fuse_chunk, ids[i:i + 5])
    for t in range(T):
        for c in range(C):
            views = [{"data": s.data[t, c], "origin": {"y": 0.0, "x": 0.0}, "spacing": {"y": 1.0, "x": 1.0}} for s in sims]
            ref, _ = of.fuse(views, [p[t] for p in per_t], output_stack_properties=osp, output_chunksize=chunksize,
                             fusion_func=of.max_fusion, interpolation_order=0)
            np.testing.assert_array_equal(out[t, c], ref)


def test_weights_func_blocks_match_per_chunk_fuse():
    """content_based through hook C: every block is its own output stack with the halo the
    hook asks for, like the reference's per-chunk fuse() (_core.py:2118-2128)."""
    from multiview_stitcher_b200 import fusion as efusion
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
    out = np.zeros(ref.shape, np.float32)
    msims = [dict(v, transforms={"reg": p}) for v, p in zip(views, params)]
    fuse_chunk = _partial(msims, osp, chunksize, out, weights_func=efusion.content_based, weights_func_kwargs=wkw,
                          fusion_func=efusion.weighted_average_fusion)
    BatchFuser()(fuse_chunk, sorted(block_geometry(osp, chunksize)))
    _close(out, ref)


class _CudaArray:
    """CuPy-like device array: only ``__cuda_array_interface__``, shape, dtype, ndim (what the
    views of ``fuse(backend="cupy")`` expose, fusion/_core.py:1579-1587)."""

    def __init__(self, t):
        self._t = t
        self.__cuda_array_interface__ = t.__cuda_array_interface__
        self.shape, self.ndim = tuple(t.shape), t.ndim
        self.dtype = np.dtype(str(t.dtype).replace("torch.", ""))


def test_device_array_inputs_at_the_hooks():
    """backend="cupy": hook C and fuse_np take device arrays of another library in place."""
    import torch

    from multiview_stitcher_b200 import fusion as efusion
    from multiview_stitcher_b200.batch import BatchFuser, block_geometry

    case = cases.fusion_cases()["2d_f32_quad_lin"]
    views, params = case["views"], case["params"]
    ref, osp = of.fuse(views, params)
    dev = [dict(v, data=_CudaArray(torch.from_numpy(np.ascontiguousarray(v["data"])).cuda())) for v in views]
    chunksize = {d: max(8, int(osp["shape"][d]) // 2 + 1) for d in "yx"}
    msims = [dict(v, transforms={"reg": p}) for v, p in zip(dev, params)]
    out = np.zeros(ref.shape, dtype=ref.dtype)
    bf = BatchFuser()
    bf(_partial(msims, osp, chunksize, out, backend="cupy"), sorted(block_geometry(osp, chunksize)))
    _close(out, ref)
    assert bf.h2d_bytes == 0  # nothing was uploaded: the device arrays were used in place
    # fuse_np on device slices
    cprops = osp
    got = efusion.fuse_np(dev, params, cprops, full_view_bbs=[of.view_bb(v) for v in views])
    _close(got, ref)
