"""Generates tests/golden/ngff_golden.npz by running the REFERENCE's own statement of the
Zarr v2 / OME-Zarr 0.4 encoding -- ``ngff_utils.VirtualOMEZarr`` (array_zarray, read_chunk,
root_zattrs; ngff_utils.py:196-450) and ``calc_ngff_coordinate_transformations_and_axes``
(:1493-1561), loaded through _ref_loader -- on seeded multiscale images.  Run in this
container only (needs /root/reference):

    python tests/golden/make_golden_ngff.py
"""
import importlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import _ref_loader  # noqa: E402

SP = ("z", "y", "x")


class Sim5:
    """The few xarray attributes VirtualOMEZarr touches on a (t, c, z, y, x) image."""

    def __init__(self, data, dims, origin, spacing, chunks):
        self.data, self.dims = data, tuple(dims)
        self.origin, self.spacing = origin, spacing
        self.encoding = {"preferred_chunks": dict(zip(dims, chunks))}
        self.attrs = {}

    shape = property(lambda s: s.data.shape)
    dtype = property(lambda s: s.data.dtype)
    ndim = property(lambda s: s.data.ndim)
    sizes = property(lambda s: dict(zip(s.dims, s.data.shape)))

    def isel(self, indexers):
        sl = tuple(indexers.get(d, slice(None)) for d in self.dims)
        return Sim5(self.data[sl], self.dims, self.origin, self.spacing, [1] * len(self.dims))


class Msim(dict):
    attrs = {}


def ngff_cases():
    """name -> dict(levels=[(data, origin, spacing)], dims, chunks)."""
    rng = np.random.default_rng(5)
    out = {}
    a = rng.integers(0, 4000, (2, 2, 9, 21, 30)).astype(np.uint16)
    out["tczyx_u16"] = dict(dims=("t", "c", "z", "y", "x"), chunks=(1, 1, 4, 8, 16),
                            levels=[(a, {"z": -3.0, "y": 10.5, "x": 2.25}, {"z": 2.0, "y": 0.5, "x": 0.5}),
                                    (a[:, :, ::2, ::2, ::2].copy(), {"z": -2.0, "y": 10.75, "x": 2.5}, {"z": 4.0, "y": 1.0, "x": 1.0})])
    b = rng.random((37, 50)).astype(np.float32)
    out["yx_f32"] = dict(dims=("y", "x"), chunks=(16, 64), levels=[(b, {"y": 0.0, "x": -7.0}, {"y": 1.0, "x": 1.3})])
    c = rng.integers(0, 255, (1, 5, 6, 7)).astype(np.uint8)
    out["czyx_u8_exact"] = dict(dims=("c", "z", "y", "x"), chunks=(1, 5, 3, 7),
                                levels=[(c, {"z": 0.0, "y": 0.0, "x": 0.0}, {"z": 1.0, "y": 1.0, "x": 1.0})])
    return out


def transform_cases():
    return {
        "3d_tc": dict(stack_properties_res0={"spacing": {"z": 2.0, "y": 0.5, "x": 0.25}, "origin": {"z": -3.0, "y": 10.5, "x": 2.25},
                                             "shape": {"z": 9, "y": 500, "x": 800}},
                      res_abs_factors=[{"z": 1, "y": 1, "x": 1}, {"z": 1, "y": 2, "x": 2}, {"z": 1, "y": 4, "x": 4}],
                      nsdims=["t", "c"], time_transform={"scale": 2.5, "translation": 1.0, "unit": "second"}),
        "2d_plain": dict(stack_properties_res0={"spacing": {"y": 1.0, "x": 1.3}, "origin": {"y": 0.1, "x": -7.0}, "shape": {"y": 300, "x": 300}},
                         res_abs_factors=[{"y": 1, "x": 1}, {"y": 2, "x": 2}], nsdims=[], time_transform=None),
    }


if __name__ == "__main__":
    _ref_loader.load_reference()
    si = sys.modules["multiview_stitcher.spatial_image_utils"]
    si.get_spatial_dims_from_sim = lambda sim: [d for d in sim.dims if d in SP]
    si.get_spacing_from_sim = lambda sim: dict(sim.spacing)
    si.get_origin_from_sim = lambda sim: dict(sim.origin)
    si._get_backend_data = lambda sim: sim.data
    msi = sys.modules["multiview_stitcher.msi_utils"]
    msi.is_msim = lambda m: isinstance(m, Msim)
    msi.get_sorted_scale_keys = lambda m: sorted(m.keys())
    msi.get_sim_from_msim = lambda m, scale="scale0": m[scale]
    sys.modules.pop("multiview_stitcher.ngff_utils", None)
    ngff = importlib.import_module("multiview_stitcher.ngff_utils")
    Msim.__getitem__ = lambda self, k: dict.__getitem__(self, k.split("/")[0])

    arrays = {}
    for name, case in ngff_cases().items():
        msim = Msim({f"scale{i}": Sim5(d, case["dims"], o, s, case["chunks"]) for i, (d, o, s) in enumerate(case["levels"])})
        v = ngff.VirtualOMEZarr(msim, name="image")
        arrays[f"{name}/zattrs"] = np.array(json.dumps(v.root_zattrs()))
        arrays[f"{name}/zgroup"] = np.array(json.dumps(v.root_zgroup()))
        for lvl, (d, _, _) in enumerate(case["levels"]):
            za = v.array_zarray(str(lvl))
            arrays[f"{name}/{lvl}/zarray"] = np.array(json.dumps(za))
            grid = [-(-n // c) for n, c in zip(za["shape"], za["chunks"])]
            for idx in np.ndindex(*grid):
                key = "/".join(str(i) for i in idx)
                arrays[f"{name}/{lvl}/chunk/{key}"] = np.frombuffer(v.read_chunk(str(lvl), key), dtype=np.uint8)
        print(name, len([k for k in arrays if k.startswith(name)]), "entries")
    for name, kw in transform_cases().items():
        coordtfs, axes = ngff.calc_ngff_coordinate_transformations_and_axes(**kw)
        arrays[f"tf/{name}"] = np.array(json.dumps({"coordtfs": coordtfs, "axes": axes}))
    np.savez_compressed(os.path.join(HERE, "ngff_golden.npz"), **arrays)
    print("saved", len(arrays), "arrays")
