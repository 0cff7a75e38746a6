"""Pair preparation and batched pairwise registration with the views resident in HBM
(hook A ``pairwise_executor``, section 8b).

What ``register_pair_of_msims`` (registration.py:1547-2058) does per pair on the
CPU -- mean-binning, overlap boxes, crop with one pixel of margin, resampling of
both crops onto the fixed view's pixel grid (``sims_to_intrinsic_coord_system``,
:280-350, through dask-image), the ``pairwise_reg_func`` hook, and the conversion
of the pixel translation into a physical transform (:1382-1474) -- restructured for
the GPU:

* the geometry of ALL pairs is worked out on the host first (a few float64
  operations per pair: half-space intersection of the two transformed boxes,
  ``mv_graph.py:183-338``; coordinate selection, ``spatial_image_utils.py:1278``);
* views are uploaded once and stay on the device; binning is one kernel per view
  (``mvs_bin_mean``), crops are strided windows of the resident tensors (no copy);
* every group of pairs with the same crop shape is resampled by ONE launch of the
  fusion path's resampler (``mvs_resample_views``: 2 x n_pairs "views", each with
  its own pixel matrix / offset, NaN outside) straight into the (2P, *crop) stack
  that ``registration.register_pairs`` consumes.

The host geometry uses ``scipy.optimize.linprog`` and
``scipy.spatial.HalfspaceIntersection`` -- the same library calls the reference
makes (``mv_graph.py:320-330``) -- because the crop shape is ``floor()`` of a
difference of polytope vertices (registration.py:311) and only the same qhull
arithmetic reproduces the reference's shapes on grid-aligned tiles.
"""

from __future__ import annotations

import ctypes

import numpy as np

from . import _lib, geometry
from ._lib import EngineError

# --- coordinates ----------------------------------------------------------------------


class _Axes:
    """Per-dimension coordinate arrays of a view: the reference reads origin and spacing
    back from them after every selection / coarsening (spatial_image_utils.py:554-589)."""

    def __init__(self, dims, coords):
        self.dims = list(dims)
        self.coords = [np.asarray(c, dtype=np.float64) for c in coords]

    @classmethod
    def of_view(cls, dv):
        # spatial_image_utils.py:316-317
        return cls(dv.dims, [dv.origin[d] + dv.spacing[d] * np.arange(n, dtype=float) for d, n in zip(dv.dims, dv.shape)])

    @property
    def origin(self):
        return np.array([c[0] for c in self.coords])

    @property
    def spacing(self):
        return np.array([(c[1] - c[0]) if len(c) > 1 else 1.0 for c in self.coords])

    @property
    def shape(self):
        return np.array([len(c) for c in self.coords])

    def binned(self, b):
        # coarsen(boundary="trim").mean() of the coordinates (registration.py:1732-1743)
        out = []
        for c, k in zip(self.coords, b):
            n = len(c) // k
            out.append(c[: n * k].reshape(n, k).mean(axis=1))
        return _Axes(self.dims, out)

    def select(self, lo, hi):
        """Label-based, inclusive ``sel(slice(lo, hi))`` (pandas ``slice_locs`` on an
        increasing index): index ranges and the selected axes."""
        rng, out = [], []
        for c, a, b in zip(self.coords, lo, hi):
            i0 = int(np.searchsorted(c, a, side="left"))
            i1 = int(np.searchsorted(c, b, side="right"))
            rng.append((i0, i1))
            out.append(c[i0:i1])
        return rng, _Axes(self.dims, out)


def _tolerance(overlap_tolerance, dims):
    # registration.py:1624-1637
    if overlap_tolerance is None:
        return {d: 0.0 for d in dims}
    if isinstance(overlap_tolerance, (int, float)):
        return {d: float(overlap_tolerance) for d in dims}
    return {d: float(overlap_tolerance.get(d, 0.0)) for d in dims}


def optimal_registration_binning(shape1, shape2, spacing1, spacing2, dims, max_total_pixels_per_stack=400**3):
    """registration.py:114-191: increment the binning of the finest-spaced axis
    (x and y tied) until the larger tile has fewer than 400^3 voxels."""
    ndim = len(dims)
    ext = [max(a, b) for a, b in zip(shape1, shape2)]
    binning = [1] * ndim
    sp = [list(spacing1), list(spacing2)]
    cur = [list(spacing1), list(spacing2)]
    while np.prod([e / b for e, b in zip(ext, binning)]) >= max_total_pixels_per_stack:
        k = int(np.argmin([min(cur[0][i], cur[1][i]) for i in range(ndim)]))
        if ndim == 3 and k == 0:
            binning[0] += 1
        else:
            binning[-1] += 1
            binning[-2] += 1
        cur = [[s[i] * binning[i] for i in range(ndim)] for s in sp]
    return dict(zip(dims, binning))


# --- overlap polytope (mv_graph.py:183-338) -------------------------------------------


def _box_halfspaces(origin, spacing, shape, affine):
    """Rows ``[n, c]`` (``n.x + c <= 0`` inside) of the pixel-centre box of a view, mapped
    by ``affine``; face / normal / sign conventions of mv_graph.py:183-218, 386-420."""
    ndim = len(origin)
    affine = np.asarray(affine, dtype=np.float64)
    unit = np.array(list(np.ndindex(*([2] * ndim))))
    extent = (np.asarray(shape) - 1) * np.asarray(spacing)
    center = np.asarray(origin) + np.asarray(spacing) * (np.asarray(shape) - 1) / 2
    center = np.matmul(affine, np.concatenate([center, np.ones(1)]))[:ndim]
    rows = []
    for ax in range(ndim):
        for side in (0, 1):
            f = unit[np.where(unit[:, ax] == side)[0]] * extent + np.asarray(origin)
            f = np.dot(affine, np.hstack([f, np.ones((f.shape[0], 1))]).T).T[:, :-1]
            if ndim == 2:
                n = np.array([-(f[1][1] - f[0][1]), f[1][0] - f[0][0]])
            else:
                n = np.cross(f[1] - f[0], f[2] - f[0])
            n = n / np.linalg.norm(n)
            if np.dot(n, center) + -np.dot(n, f[0]) > 0:
                n = -n
            rows.append(np.concatenate([n, [-np.dot(n, f[0])]]))
    return np.array(rows)


def overlap_vertices(box1, box2):
    """Vertices of the intersection of two transformed boxes ``(origin, spacing, shape,
    affine)``: Chebyshev centre as interior point, then qhull (mv_graph.py:301-338).
    Raises ``EngineError`` when the boxes do not overlap."""
    try:
        from scipy.optimize import linprog
        from scipy.spatial import HalfspaceIntersection, QhullError
    except ImportError as e:  # pragma: no cover
        raise EngineError("pair preparation needs scipy (linprog, HalfspaceIntersection) for the host geometry") from e
    eqs = np.concatenate([_box_halfspaces(*box1), _box_halfspaces(*box2)])
    norms = np.reshape(np.linalg.norm(eqs[:, :-1], axis=1), (eqs.shape[0], 1))
    cost = np.zeros((eqs.shape[1],))
    cost[-1] = -1
    lp = linprog(cost, A_ub=np.hstack((eqs[:, :-1], norms)), b_ub=-eqs[:, -1:], bounds=(None, None))
    if lp.x is None:
        raise EngineError("views do not overlap")
    try:
        return HalfspaceIntersection(eqs, lp.x[:-1]).intersections
    except QhullError as e:
        raise EngineError("views do not overlap") from e


def _to_intrinsic(affine, pts):
    inv = np.linalg.inv(affine)
    h = np.concatenate([pts, np.ones((pts.shape[0], 1))], axis=1)
    return np.array([np.dot(inv, p) for p in h])[:, :-1]  # transformation.py:151-161


def overlap_bboxes(axes1, axes2, affine1, affine2, tol, intrinsic=True, return_vertices=False):
    """``_get_overlap_bboxes`` (registration.py:194-277): ``(lowers, uppers)`` of the
    overlap polytope, per view in its intrinsic physical coordinates, or in the
    world system."""
    boxes = []
    for ax, A in ((axes1, affine1), (axes2, affine2)):
        sp = ax.spacing
        shape = ax.shape + np.array([int(np.ceil(2 * tol[d] / s)) for d, s in zip(ax.dims, sp)])
        origin = ax.origin - np.array([tol[d] for d in ax.dims])
        boxes.append((origin, sp, shape, A))
    verts = overlap_vertices(*boxes)
    per_view = [_to_intrinsic(A, verts) for A in (affine1, affine2)] if intrinsic else [verts, verts]
    if return_vertices:
        return [v.min(axis=0) for v in per_view], [v.max(axis=0) for v in per_view], verts
    return [v.min(axis=0) for v in per_view], [v.max(axis=0) for v in per_view]


# --- device steps ---------------------------------------------------------------------


def bin_view(dv, binning, skip_nan=True):
    """``sim.coarsen(binning, boundary="trim").mean().astype(dtype)`` of a resident view
    (registration.py:1732-1743) -> CUDA tensor of the same dtype.  ``skip_nan=False``:
    ``np.mean`` semantics (NaNs propagate; the pyramid step)."""
    import torch

    lib = _lib.load(require_device=True)
    t = dv.tensor
    nd = t.ndim
    b = [int(binning.get(d, 1)) for d in dv.dims]
    out_shape = [s // k for s, k in zip(t.shape, b)]
    out = torch.empty(out_shape, dtype=t.dtype, device=t.device)
    pad = 3 - nd
    _lib.check(
        lib.mvs_bin_mean(
            ctypes.c_void_p(t.data_ptr()), dv.mvs_dtype,
            (ctypes.c_int32 * 3)(*([1] * pad + list(t.shape))),
            (ctypes.c_int64 * 3)(*([0] * pad + list(t.stride()))),
            (ctypes.c_int32 * 3)(*([1] * pad + b)), int(bool(skip_nan)),
            ctypes.c_void_p(out.data_ptr()), _lib.current_stream_ptr(),
        ),
        "mvs_bin_mean",
    )
    return out


class PreparedPairs:
    """Result of ``prepare_pairs``: per pair the fixed / moving float32 CUDA crops on the
    common pixel grid (NaN outside) and that grid (origin / spacing / shape arrays)."""

    def __init__(self, n):
        self.fixed = [None] * n
        self.moving = [None] * n
        self.grid = [None] * n
        self.lowers = [None] * n
        self.uppers = [None] * n
        self.launches = 0


def plan_pair(axes1, axes2, affine1, affine2, tol):
    """Host geometry of one pair (registration.py:1745-1779, :280-316): crop index ranges
    per view, the common grid, and the pixel matrices / offsets ``transform_sim`` would
    hand to scipy for the fixed (identity) and the moving (``inv(A2) @ A1``) crop."""
    lowers, uppers, verts = overlap_bboxes(axes1, axes2, affine1, affine2, tol, return_vertices=True)
    eps = 1e-6
    rng, crop_axes = [], []
    for k, ax in enumerate((axes1, axes2)):
        sp = ax.spacing
        r, c = ax.select(lowers[k] - eps - sp, uppers[k] + eps + sp)  # one pixel of margin
        if any(i1 - i0 < 1 for i0, i1 in r):
            raise EngineError("empty overlap crop")
        rng.append(r)
        crop_axes.append(c)
    spacing = np.max([c.spacing for c in crop_axes], axis=0)
    shape = np.floor(np.array(uppers[0] - lowers[0]) / spacing + 1).astype(np.uint64).astype(np.int64)
    origin = np.asarray(lowers[0], dtype=np.float64)
    transf = np.matmul(np.linalg.inv(affine2), affine1)
    ndim = len(shape)
    xf = []
    for c, p in ((crop_axes[0], np.eye(ndim + 1)), (crop_axes[1], transf)):
        xf.append(geometry.pixel_affine(p, origin, spacing, c.origin, c.spacing))
    return {"ranges": rng, "origin": origin, "spacing": spacing, "shape": tuple(int(s) for s in shape),
            "xforms": xf, "lowers": lowers, "uppers": uppers, "world_vertices": verts}


class PairPlan:
    """Host geometry of a set of pairs, worked out once (the overlap polytopes cost a
    linear programme and a qhull call per pair) and reusable for every time point /
    channel of a dataset whose views keep their shapes, coordinates and transforms.

    ``views``: ``DeviceView``s, view dicts or xarray-likes (only shape / origin / spacing
    are read); ``affines``: per view the (ndim+1)^2 pre-registration transform (view
    physical -> world); ``pairs``: ``(i, j)`` view indices, i fixed / j moving;
    ``registration_binning``: dict, or None for the reference's heuristic per pair."""

    def __init__(self, views, affines, pairs, overlap_tolerance=None, registration_binning=None):
        from .fusion import DeviceView, _view_fields

        meta = []
        for v in views:
            if isinstance(v, DeviceView):
                meta.append((v.dims, v.origin, v.spacing, tuple(v.shape)))
            elif isinstance(v, dict):
                dims = geometry.spatial_dims(len(v["data"].shape))
                meta.append((dims, v["origin"], v["spacing"], tuple(int(n) for n in v["data"].shape)))
            elif hasattr(v, "coords") and hasattr(v, "dims"):  # xarray-like: no data access (may be lazy)
                dims = [d for d in v.dims if d in geometry.SPATIAL_DIMS]
                cs = {d: np.asarray(getattr(v.coords[d], "values", v.coords[d])) for d in dims}
                # the coordinate arrays themselves, not origin + spacing * arange
                meta.append((dims, None, None, tuple(len(cs[d]) for d in dims), [cs[d] for d in dims]))
            else:
                data, origin, spacing = _view_fields(v)
                dims = geometry.spatial_dims(len(data.shape))
                meta.append((dims, origin, spacing, tuple(data.shape)))
        self.ndim = len(meta[0][3])
        self.dims = geometry.spatial_dims(self.ndim)
        self.shapes = [m[3] for m in meta]
        self.affines = [np.asarray(a, dtype=np.float64) for a in affines]
        self.pairs = [tuple(int(i) for i in e) for e in pairs]
        self.tol = _tolerance(overlap_tolerance, self.dims)
        dims = self.dims

        class _Meta:
            def __init__(self, m):
                self.dims, self.origin, self.spacing, self.shape = m[:4]

        base = [_Axes(m[0], m[4]) if len(m) > 4 else _Axes.of_view(_Meta(m)) for m in meta]
        axes_at = {}

        def axes(i, b):
            if (i, b) not in axes_at:
                axes_at[(i, b)] = base[i].binned(b) if max(b) > 1 else base[i]
            return axes_at[(i, b)]

        binnings = []
        for i, j in self.pairs:
            if registration_binning is None:
                bd = optimal_registration_binning(self.shapes[i], self.shapes[j], base[i].spacing, base[j].spacing, dims)
            else:
                bd = registration_binning
            b = tuple(int(bd.get(d, 1)) for d in dims)
            binnings.append(b)
            axes(i, b), axes(j, b)

        def plan_one(k):
            (i, j), b = self.pairs[k], binnings[k]
            pl = plan_pair(axes_at[(i, b)], axes_at[(j, b)], self.affines[i], self.affines[j], self.tol)
            if any(s < 1 for s in pl["shape"]):
                raise EngineError(f"pair {(i, j)}: empty overlap grid {pl['shape']}")
            pl["binning"] = b
            # world-space box of the un-binned views (registration.py:2038-2056): the same
            # polytope when nothing was binned, else a second intersection
            if max(b) > 1:
                lo, hi = overlap_bboxes(base[i], base[j], self.affines[i], self.affines[j], self.tol, intrinsic=False)
                return pl, np.array([lo[0], hi[0]])
            v = pl["world_vertices"]
            return pl, np.array([v.min(axis=0), v.max(axis=0)])

        planned = [plan_one(k) for k in range(len(self.pairs))]
        self.items = [p for p, _ in planned]
        self.bbox = [b for _, b in planned]
        self._stacks = {}  # crop shape -> (2P, *crop) float32 stack kept for reuse_buffers
        self._xarr, self._xarr_key = {}, None  # crop shape -> resample records of the current view tensors
        self.groups = {}
        for k, pl in enumerate(self.items):
            self.groups.setdefault(pl["shape"], []).append(k)
        self._axes_at = axes_at

    @property
    def used_views(self):
        return sorted({i for e in self.pairs for i in e})

    def grid(self, k):
        pl = self.items[k]
        return {"origin": pl["origin"], "spacing": pl["spacing"], "shape": pl["shape"]}

    def prepare(self, views, reuse_buffers=False):
        """Device half: bin (one kernel per view and binning), window, and ONE resample
        launch per crop-shape group.  Returns ``PreparedPairs``.  ``reuse_buffers``: the crop
        stacks live in the plan and are overwritten by the next call (time lapses: several GB per
        crop shape for 3-D tiles, and a fresh ``cudaMalloc`` of that size costs more than the
        registration itself -- measured 200 ms vs 500-900 ms per call on C3)."""
        import torch

        from .fusion import DeviceView, to_device_view

        lib = _lib.load(require_device=True)
        if len(views) != len(self.shapes):
            raise EngineError("PairPlan.prepare: number of views differs from the planned one")
        # only the views this plan's pairs touch are uploaded (a rank of a sharded run owns
        # a subset of the pairs)
        dviews = {i: to_device_view(views[i]) for i in self.used_views}
        if any(tuple(v.shape) != self.shapes[i] for i, v in dviews.items()):
            raise EngineError("PairPlan.prepare: view shapes differ from the planned ones")
        ndim, dims = self.ndim, self.dims
        binned = {}

        def view_at(i, b):
            if (i, b) not in binned:
                if max(b) > 1:
                    ax = self._axes_at[(i, b)]
                    t = bin_view(dviews[i], dict(zip(dims, b)))
                    binned[(i, b)] = DeviceView(t, dict(zip(dims, ax.origin)), dict(zip(dims, ax.spacing)))
                else:
                    binned[(i, b)] = dviews[i]
            return binned[(i, b)]

        out = PreparedPairs(len(self.pairs))
        halo = (ctypes.c_int32 * 3)(0, 0, 0)
        # the per-crop records depend only on the plan and on where the views live: kept while the same
        # un-binned tensors come back (time lapses re-fill them; binned views are new tensors every call)
        ptr_key = tuple((i, v.tensor.data_ptr(), tuple(v.tensor.stride())) for i, v in sorted(dviews.items()))
        unbinned = all(max(pl["binning"]) <= 1 for pl in self.items)
        if not unbinned or self._xarr_key != ptr_key:
            self._xarr, self._xarr_key = {}, ptr_key if unbinned else None
        for shape, idx in self.groups.items():
            xarr = self._xarr.get(shape)
            if xarr is None:
                xarr = np.zeros(2 * len(idx), dtype=_lib.VIEW_XFORM_DTYPE)
                for r, k in enumerate(idx):
                    pl = self.items[k]
                    for side in (0, 1):
                        dv = view_at(self.pairs[k][side], pl["binning"])
                        win = dv.tensor[tuple(slice(i0, i1) for i0, i1 in pl["ranges"][side])]
                        x = xarr[2 * r + side]
                        x["data"] = win.data_ptr()
                        x["dtype"] = dv.mvs_dtype
                        x["shape"] = [1] * (3 - ndim) + list(map(int, win.shape))
                        x["stride"] = [0] * (3 - ndim) + [int(s) for s in win.stride()]
                        x["matrix"], x["offset"] = geometry.embed3(*pl["xforms"][side])
                        x["wmatrix"], x["woffset"] = geometry.embed3(np.eye(ndim), np.zeros(ndim))
                if unbinned:
                    self._xarr[shape] = xarr
            stack = self._stacks.get(shape) if reuse_buffers else None
            if stack is None or stack.shape[0] != 2 * len(idx):
                stack = torch.empty((2 * len(idx),) + tuple(shape), dtype=torch.float32, device="cuda")
                if reuse_buffers:
                    self._stacks[shape] = stack
            _lib.check(
                lib.mvs_resample_views(
                    xarr.ctypes.data_as(ctypes.c_void_p), len(xarr), None, 0,
                    (ctypes.c_int32 * 3)(*((1,) * (3 - ndim) + tuple(shape))), halo, ndim, 1,
                    ctypes.c_void_p(stack.data_ptr()), None, _lib.current_stream_ptr(),
                ),
                "mvs_resample_views",
            )
            out.launches += 1
            for r, k in enumerate(idx):
                out.fixed[k] = stack[2 * r]
                out.moving[k] = stack[2 * r + 1]
        for k, pl in enumerate(self.items):
            out.grid[k] = self.grid(k)
            out.lowers[k], out.uppers[k] = pl["lowers"], pl["uppers"]
        out._keepalive = (dviews, binned)
        return out


def prepare_pairs(views, affines, pairs, overlap_tolerance=None, registration_binning=None):
    """Crops of all ``pairs`` on their common pixel grids: ``PairPlan(...).prepare(views)``."""
    from .fusion import to_device_view

    dviews = [to_device_view(v) for v in views]
    return PairPlan(dviews, affines, pairs, overlap_tolerance, registration_binning).prepare(dviews)


def physical_transform(affine_px, grid, affine_fixed):
    """``get_affine_from_intrinsic_affine`` (registration.py:1382-1474) for the pixel-space
    branch: both crops live on ``grid`` and carry the fixed view's transform, so
    ``M_W = (A T S) M_D (A T S)^-1``."""
    n = len(grid["origin"])
    T = np.eye(n + 1)
    T[:n, n] = grid["origin"]
    S = np.diag(list(grid["spacing"]) + [1])
    d2w = np.matmul(np.array(affine_fixed), np.matmul(T, S))
    return np.matmul(d2w, np.matmul(affine_px, np.linalg.inv(d2w)))


def register_views(views, affines=None, pairs=None, overlap_tolerance=None, registration_binning=None,
                   pairwise_reg_func_kwargs=None, return_prepared=False, plan=None, pc_plans=None):
    """Batched ``register_pair_of_msims`` for an in-memory dataset: one
    ``{"transform", "quality", "bbox", "affine_matrix"}`` per pair -- ``transform`` the
    physical affine (fixed world -> moving world, :2008-2015), ``bbox`` the world-space
    overlap box ``[lower, upper]`` of the un-binned views (:2038-2056).  ``plan``: a
    ``PairPlan`` to reuse (then ``affines`` / ``pairs`` / tolerance / binning are the
    plan's); ``pc_plans``: dict that keeps the phase-correlation buffers across calls."""
    from . import registration

    reuse = plan is not None and not return_prepared  # a plan the caller keeps: its crop stacks stay too
    if plan is None:
        plan = PairPlan(views, affines, pairs, overlap_tolerance, registration_binning)
    prep = plan.prepare(views, reuse_buffers=reuse)
    kw = dict(pairwise_reg_func_kwargs or {})
    res = registration.register_pairs(
        prep.fixed, prep.moving, kw.pop("disambiguate_region_mode", None), kw.pop("upsample_factor", None),
        plans=pc_plans,
    )
    if kw:
        raise EngineError(f"unsupported pairwise_reg_func_kwargs: {sorted(kw)}")
    out = []
    for k, (i, j) in enumerate(plan.pairs):
        a_px = np.asarray(res[k]["affine_matrix"], dtype=np.float64)
        out.append({
            "transform": physical_transform(a_px, prep.grid[k], plan.affines[i]),
            "quality": res[k]["quality"],
            "bbox": plan.bbox[k],
            "affine_matrix": a_px,
        })
    return (out, prep) if return_prepared else out


# --- hook A: pairwise_executor (registration.py:2634-2655) ----------------------------


def _n_timepoints(msim):
    if isinstance(msim, dict) and "data" in msim:
        return None
    sim = msim["scale0/image"]
    return int(sim.sizes["t"]) if hasattr(sim, "dims") and "t" in sim.dims else None


def _t_coords(msim, nt):
    """Time coordinate VALUES of an msim (the reference assigns them to the results,
    registration.py:2091; its groupwise resolution selects per time point by these labels,
    param_resolution/utils.py:23-39).  An msim without "t" gets ``ensure_dim``'s single
    coordinate 0 (msi_utils.ensure_dim, registration.py:2070-2071)."""
    if nt is None:
        return np.array([0])
    sim = msim["scale0/image"]
    try:
        c = sim.coords["t"]
        vals = np.asarray(c.values if hasattr(c, "values") else c)
        if vals.shape == (nt,):
            return vals
    except Exception:
        pass
    return np.arange(nt)


def _view_and_affine(msim, transform_key, it=None):
    """(view, affine) of one element of ``msims`` at time index ``it``: a
    MultiscaleSpatialImage-like mapping (``msim["scale0/image"]``,
    ``msim["scale0"][transform_key]``; msi_utils.py:108-113, 351-361) or a plain view dict
    carrying ``"transforms": {key: affine}``."""
    if isinstance(msim, dict) and "data" in msim:
        return msim, np.asarray(msim["transforms"][transform_key], dtype=np.float64)
    sim = msim["scale0/image"]
    aff = msim["scale0"][transform_key]
    if hasattr(sim, "dims") and "t" in sim.dims:
        sim = sim.isel(t=0 if it is None else it)
    if hasattr(aff, "dims") and "t" in aff.dims:
        aff = aff.isel(t=0 if it is None or aff.sizes["t"] == 1 else it)
    if hasattr(sim, "dims") and "c" in sim.dims:
        raise EngineError("pairwise_executor: select the registration channel first (register(reg_channel=...))")
    return sim, np.asarray(getattr(aff, "data", aff), dtype=np.float64)


def pairwise_executor(msims, edges, register_kwargs):
    """Drop-in ``pairwise_executor`` for ``registration.register`` (called as
    ``pairwise_executor(msims, edges, register_kwargs)``, registration.py:2649-2655):
    all ``edges`` are prepared and registered on the GPU in a handful of launches per
    time point (the reference loops ``register_pair_of_msims`` over "t", :2061-2093; the
    host plan is reused while the transforms do not change).  Returns one
    ``{"transform" (t, n+1, n+1), "quality" (t,), "bbox" (t, 2, n)}`` per edge --
    ``xr.DataArray``s with the reference's dims when xarray is importable, else numpy
    arrays of those shapes."""
    kw = dict(register_kwargs)
    transform_key = kw.pop("transform_key")
    func = kw.pop("pairwise_reg_func", None)
    if func is not None and getattr(func, "__name__", "") != "phase_correlation_registration":
        raise EngineError("the batched executor implements phase_correlation_registration only")
    for ignored in ("points_key", "prefilter_markers", "n_parallel_pairwise_regs"):
        kw.pop(ignored, None)
    if kw.pop("reg_res_level", None) not in (None, 0):
        raise EngineError("pairwise_executor: reg_res_level other than 0 is not supported")
    tolerance = kw.pop("overlap_tolerance", None)
    binning = kw.pop("registration_binning", None)
    func_kwargs = kw.pop("pairwise_reg_func_kwargs", None)
    if kw:
        raise EngineError(f"pairwise_executor: unsupported register kwargs {sorted(kw)}")
    edges = [tuple(e) for e in edges]
    if not edges:
        return []
    nt = _n_timepoints(msims[0])
    per_t, plan, plan_affines, pc_plans = [], None, None, {}
    for it in range(nt or 1):
        va = [_view_and_affine(m, transform_key, it if nt else None) for m in msims]
        affines = [a for _, a in va]
        if plan is None or any(not np.array_equal(a, b) for a, b in zip(affines, plan_affines)):
            plan = PairPlan([v for v, _ in va], affines, edges, tolerance, binning)
            plan_affines = affines
        per_t.append(register_views([v for v, _ in va], plan=plan, pairwise_reg_func_kwargs=func_kwargs, pc_plans=pc_plans))
    for p in pc_plans.values():
        p.close()
    try:
        import xarray as xr
    except ImportError:
        xr = None
    n = plan_affines[0].shape[0] - 1
    labels = geometry.spatial_dims(n) + ["1"]
    tvals = _t_coords(msims[0], nt)
    out = []
    for k in range(len(edges)):
        tr = np.stack([r[k]["transform"] for r in per_t])
        q = np.array([r[k]["quality"] for r in per_t], dtype=float)
        bb = np.stack([r[k]["bbox"] for r in per_t])
        if xr is not None:
            tr = xr.DataArray(tr, dims=["t", "x_in", "x_out"], coords={"t": tvals, "x_in": labels, "x_out": labels})
            q = xr.DataArray(q, dims=["t"], coords={"t": tvals})
            bb = xr.DataArray(bb, dims=["t", "point_index", "dim"], coords={"t": tvals})
        out.append({"transform": tr, "quality": q, "bbox": bb})
    pairwise_executor.last_t_coords = tvals
    return out
