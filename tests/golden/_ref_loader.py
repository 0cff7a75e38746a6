"""Import the reference's OWN hot-path modules in the build container.

Only usable where ``/root/reference`` exists (the build container) -- never
on the GPU box.  ``make_golden.py`` uses it to run the reference's real
``fuse_np`` / ``transform_sim`` / ``get_blending_weights`` / ``content_based``
/ ``phase_correlation_registration`` code and store the results as fixtures.

The reference's data-model dependencies (dask, xarray, zarr, dask-image,
scikit-image ...) are not installed in this image.  They are replaced by
inert placeholder modules, and ``multiview_stitcher.spatial_image_utils``
(the xarray data model, out of scope -- SURVEY.md section 2) by a ~60-line
fake that carries ``data / dims / origin / spacing``.  The arithmetic that
runs is the reference's, unmodified.  scikit-image's three functions are
bound to ``oracle.skimage_restated`` (so registration fixtures pin the
reference's candidate loop, not scikit-image itself).
"""

from __future__ import annotations

import importlib
import importlib.abc
import importlib.machinery
import inspect
import os
import sys
import types

import numpy as np

REF_SRC = "/root/reference/src"
_PLACEHOLDER_ROOTS = {
    "dask",
    "dask_image",
    "xarray",
    "zarr",
    "skimage",
    "spatial_image",
    "multiscale_spatial_image",
    "ngff_zarr",
    "tifffile",
    "ome_zarr",
    "fsspec",
    "aiohttp",
    "numcodecs",
    "matplotlib",
    "ants",
    "itk",
}


class _Meta(type):
    """Placeholder classes: any attribute is another placeholder class, any
    call returns one too, so import-time decorators / annotations work."""

    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Meta(name, (object,), {})

    def __call__(cls, *a, **k):
        return _Meta("called_" + cls.__name__, (object,), {})

    def __iter__(cls):
        return iter(())


class _PlaceholderModule(types.ModuleType):
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Meta(name, (object,), {})


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _PLACEHOLDER_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _PlaceholderModule(spec.name)

    def exec_module(self, module):
        pass


class FakeSim:
    """Minimal stand-in for the reference's xarray 'sim'."""

    def __init__(self, data, dims, origin, spacing, attrs=None, coords=None):
        self.data = data
        self.dims = tuple(dims)
        self.origin = {d: float(origin[d]) for d in dims}
        self.spacing = {d: float(spacing[d]) for d in dims}
        self.attrs = attrs if attrs is not None else {}
        self.coord_arrays = coords  # explicit coordinates (after sel / coarsen), else None

    @property
    def dtype(self):
        return self.data.dtype

    @property
    def shape(self):
        return self.data.shape

    @property
    def ndim(self):
        return self.data.ndim

    def astype(self, dtype):
        return FakeSim(self.data.astype(dtype), self.dims, self.origin, self.spacing, attrs=dict(self.attrs),
                       coords=self.coord_arrays)

    def copy(self, data=None):
        return FakeSim(
            self.data.copy() if data is None else data,
            self.dims,
            self.origin,
            self.spacing,
        )


def _fake_si_utils():
    m = types.ModuleType("multiview_stitcher.spatial_image_utils")
    m.SPATIAL_DIMS = ["z", "y", "x"]
    m.DEFAULT_TRANSFORM_KEY = "affine_metadata"
    m.DEFAULT_SPATIAL_CHUNKSIZES_3D = {"z": 256, "y": 256, "x": 256}
    m.DEFAULT_SPATIAL_CHUNKSIZES_2D = {"y": 2048, "x": 2048}
    m.get_ndim_from_sim = lambda sim: len(sim.dims)
    m.get_spatial_dims_from_sim = lambda sim: list(sim.dims)

    def _get(attr):
        def f(sim, asarray=False):
            d = getattr(sim, attr)
            return np.array([d[k] for k in sim.dims]) if asarray else dict(d)

        return f

    m.get_spacing_from_sim = _get("spacing")
    m.get_origin_from_sim = _get("origin")

    def get_shape_from_sim(sim, asarray=False):
        if asarray:
            return np.array(sim.data.shape)
        return dict(zip(sim.dims, sim.data.shape))

    m.get_shape_from_sim = get_shape_from_sim
    m._get_backend_data = lambda sim: sim.data
    m.is_dask_backed_dataarray = lambda sim: False

    def to_spatial_image(data, dims=None, scale=None, translation=None, **kw):
        dims = list(dims)
        return FakeSim(data, dims, translation, scale)

    m.to_spatial_image = to_spatial_image

    def __getattr__(name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Meta(name, (object,), {})

    m.__getattr__ = __getattr__
    return m


def _has_keyword(func, keyword):
    try:
        return keyword in inspect.signature(func).parameters
    except Exception:
        return False


_loaded = None


def load_reference():
    """Returns a namespace with the reference's real modules:
    ``transformation, weights, fusion_core, registration, param_utils``."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not os.path.isdir(REF_SRC):
        raise RuntimeError("reference tree not present (only in the build container)")
    repo_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    if repo_root not in sys.path:
        sys.path.insert(0, repo_root)
    from oracle import skimage_restated as sk

    sys.meta_path.insert(0, _Finder())
    # functional pieces of the placeholders
    import dask.utils  # placeholder

    dask.utils.has_keyword = _has_keyword
    import skimage.exposure
    import skimage.metrics
    import skimage.registration

    skimage.exposure.rescale_intensity = sk.rescale_intensity
    skimage.metrics.structural_similarity = sk.structural_similarity
    skimage.registration.phase_cross_correlation = sk.phase_cross_correlation

    pkg = types.ModuleType("multiview_stitcher")
    pkg.__path__ = [os.path.join(REF_SRC, "multiview_stitcher")]
    sys.modules["multiview_stitcher"] = pkg
    sys.modules["multiview_stitcher.spatial_image_utils"] = _fake_si_utils()
    pkg.spatial_image_utils = sys.modules["multiview_stitcher.spatial_image_utils"]
    for name in (
        "msi_utils",
        "mv_graph",
        "ngff_utils",
        "zarr_utils",
        "param_resolution",
        "transforms",
        "_zarr_compat",
    ):
        mod = _PlaceholderModule("multiview_stitcher." + name)
        sys.modules["multiview_stitcher." + name] = mod
        setattr(pkg, name, mod)

    ns = types.SimpleNamespace()
    ns.param_utils = importlib.import_module("multiview_stitcher.param_utils")
    ns.transformation = importlib.import_module("multiview_stitcher.transformation")
    ns.weights = importlib.import_module("multiview_stitcher.weights")
    ns.fusion_core = importlib.import_module("multiview_stitcher.fusion._core")
    ns.registration = importlib.import_module("multiview_stitcher.registration")
    ns.FakeSim = FakeSim
    _loaded = ns
    return ns


class FakeAffine(np.ndarray):
    """ndarray with the three xarray attributes the pair-preparation code touches
    (``.data``, ``.dims``, ``.squeeze()``)."""

    dims = ("x_in", "x_out")

    @property
    def data(self):
        return np.asarray(self)


def fake_affine(a):
    return np.asarray(a, dtype=float).view(FakeAffine)


def _coords_of(sim, d):
    if getattr(sim, "coord_arrays", None) is not None:
        return sim.coord_arrays[d]
    n = sim.data.shape[sim.dims.index(d)]
    # spatial_image_utils.py:316-317
    return sim.origin[d] + sim.spacing[d] * np.arange(n, dtype=float)


def _extend_si_utils_for_pairs(m):
    """Adds the coordinate-level helpers ``register_pair_of_msims`` calls.  These
    restate xarray behaviour (label-based ``sel``, origin/spacing read back from
    the coordinates); everything that consumes them is the reference's code."""

    # origin / spacing read back from the coordinate arrays, like the real getters
    # (spatial_image_utils.py:554-589): spacing = c[1] - c[0] carries the rounding of
    # origin + spacing * 1.0
    def get_origin_from_sim(sim, asarray=False):
        d = {k: float(_coords_of(sim, k)[0]) for k in sim.dims}
        return np.array([d[k] for k in sim.dims]) if asarray else d

    def get_spacing_from_sim(sim, asarray=False):
        d = {}
        for k in sim.dims:
            c = _coords_of(sim, k)
            d[k] = float(c[1] - c[0]) if len(c) > 1 else 1.0
        return np.array([d[k] for k in sim.dims]) if asarray else d

    m.get_origin_from_sim = get_origin_from_sim
    m.get_spacing_from_sim = get_spacing_from_sim

    def get_affine_from_sim(sim, transform_key):
        return sim.attrs["transforms"][transform_key]

    def set_sim_affine(sim, xaffine, transform_key, base_transform_key=None):
        sim.attrs.setdefault("transforms", {})[transform_key] = xaffine

    def get_stack_properties_from_sim(sim, transform_key=None, asarray=False):
        sp = {
            "shape": m.get_shape_from_sim(sim, asarray=asarray),
            "spacing": m.get_spacing_from_sim(sim, asarray=asarray),
            "origin": m.get_origin_from_sim(sim, asarray=asarray),
        }
        if transform_key is not None:
            sp["transform"] = get_affine_from_sim(sim, transform_key)
        return sp

    def extend_stack_props(stack_props, extend_by):
        # spatial_image_utils.py:889-913
        sdims = [d for d in m.SPATIAL_DIMS if d in stack_props["spacing"]]
        if not isinstance(extend_by, dict):
            extend_by = {d: extend_by for d in sdims}
        for d, val in extend_by.items():
            stack_props["shape"][d] += int(np.ceil(2 * val / stack_props["spacing"][d]))
            stack_props["origin"][d] -= val
        return stack_props

    def sim_sel_coords(sim, sel_dict):
        sl, origin, spacing, cs = [], {}, {}, {}
        for d in sim.dims:
            c = _coords_of(sim, d)
            s = sel_dict[d]
            i0 = int(np.searchsorted(c, s.start, side="left"))
            i1 = int(np.searchsorted(c, s.stop, side="right"))
            sl.append(slice(i0, i1))
            cc = c[i0:i1]
            cs[d] = cc
            origin[d] = cc[0]
            spacing[d] = cc[1] - cc[0] if len(cc) > 1 else 1.0
        return FakeSim(sim.data[tuple(sl)], sim.dims, origin, spacing, attrs=dict(sim.attrs), coords=cs)

    m.get_affine_from_sim = get_affine_from_sim
    m.set_sim_affine = set_sim_affine
    m.get_stack_properties_from_sim = get_stack_properties_from_sim
    m.extend_stack_props = extend_stack_props
    m.sim_sel_coords = sim_sel_coords


def load_reference_pairs():
    """``load_reference()`` plus the reference's real ``mv_graph`` (half-space
    geometry) wired into its ``registration`` module."""
    ns = load_reference()
    if getattr(ns, "mv_graph", None) is not None:
        return ns
    _extend_si_utils_for_pairs(sys.modules["multiview_stitcher.spatial_image_utils"])
    sys.modules.pop("multiview_stitcher.mv_graph", None)
    real = importlib.import_module("multiview_stitcher.mv_graph")
    sys.modules["multiview_stitcher"].mv_graph = real
    ns.registration.mv_graph = real
    ns.mv_graph = real
    # param_utils.py:124-150 wraps np.eye in an xr.DataArray (placeholder here)
    ns.param_utils.identity_transform = lambda ndim, t_coords=None: fake_affine(np.eye(ndim + 1))
    ns.fake_affine = fake_affine
    return ns


if __name__ == "__main__":
    ns = load_reference()
    print("loaded:", ns.fusion_core.fuse_np, ns.registration.phase_correlation_registration)
