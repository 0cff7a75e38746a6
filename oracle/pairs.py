"""CPU oracle for pair preparation: what the reference does
between ``register()`` picking an overlap pair and the ``pairwise_reg_func`` hook
seeing two same-grid float32 crops -- and the way back from the pixel-space
translation to the physical transform.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  numpy + scipy restatement
of multiview-stitcher @ 629f72d, ``src/multiview_stitcher``:

* ``registration.py:1547-2058``  ``register_pair_of_msims`` (pixel-space branch)
* ``registration.py:194-277``    ``_get_overlap_bboxes``
* ``registration.py:280-350``    ``sims_to_intrinsic_coord_system``
* ``registration.py:1382-1474``  ``get_affine_from_intrinsic_affine``
* ``registration.py:114-191``    ``get_optimal_registration_binning``
* ``mv_graph.py:183-218, 301-338, 386-444, 475-493``  box -> half-spaces -> polytope
* ``spatial_image_utils.py:316-317, 554-589, 863-913, 1278-1300``  coordinates

A view is a dict ``{"data", "origin", "spacing"}`` (dims keyed "z","y","x"); the
pre-registration transform of each view (``attrs["transforms"][transform_key]``)
is passed beside it as an ``(ndim+1, ndim+1)`` array.  Internally every view
carries its explicit coordinate arrays, because the reference reads origin and
spacing back from the xarray coordinates after every selection / coarsening.

Pinned against the reference's own functions by ``tests/golden/make_golden_pairs.py``
(``tests/golden/pairs_golden.npz``).
"""

from __future__ import annotations

import numpy as np
from scipy.optimize import linprog
from scipy.spatial import ConvexHull, HalfspaceIntersection

from . import fusion as ofusion
from . import registration as oreg

SPATIAL_DIMS = ["z", "y", "x"]


# --------------------------------------------------------------------------
# coordinates (spatial_image_utils.py)
# --------------------------------------------------------------------------


def with_coords(view):
    """View dict with explicit coordinate arrays (spatial_image_utils.py:316-317:
    ``translation + scale * arange(size)``)."""
    if "coords" in view:
        return view
    dims = SPATIAL_DIMS[-view["data"].ndim:]
    coords = {
        d: view["origin"][d] + view["spacing"][d] * np.arange(view["data"].shape[i], dtype=float)
        for i, d in enumerate(dims)
    }
    return {"data": view["data"], "coords": coords, "dims": dims}


def origin_of(v):
    """spatial_image_utils.py:554-561."""
    return {d: float(v["coords"][d][0]) for d in v["dims"]}


def spacing_of(v):
    """spatial_image_utils.py:574-589."""
    return {d: (float(v["coords"][d][1] - v["coords"][d][0]) if len(v["coords"][d]) > 1 else 1.0) for d in v["dims"]}


def shape_of(v):
    return {d: len(v["coords"][d]) for d in v["dims"]}


def stack_props(v, affine=None, extend_by=None):
    """``get_stack_properties_from_sim`` (:863-873) + ``extend_stack_props`` (:889-913)."""
    sp = {"shape": shape_of(v), "spacing": spacing_of(v), "origin": origin_of(v)}
    if affine is not None:
        sp["transform"] = np.asarray(affine, dtype=float)
    if extend_by is not None:
        for d, val in extend_by.items():
            sp["shape"][d] += int(np.ceil(2 * val / sp["spacing"][d]))
            sp["origin"][d] -= val
    return sp


def sel_coords(v, lo, hi):
    """``sim.sel({dim: slice(lo, hi)})`` (spatial_image_utils.py:1278-1283): label
    based and inclusive on both ends -- pandas ``slice_locs`` on the increasing
    float index = ``searchsorted(lo, 'left') : searchsorted(hi, 'right')``."""
    sl = []
    coords = {}
    for i, d in enumerate(v["dims"]):
        c = v["coords"][d]
        i0 = int(np.searchsorted(c, lo[i], side="left"))
        i1 = int(np.searchsorted(c, hi[i], side="right"))
        sl.append(slice(i0, i1))
        coords[d] = c[i0:i1]
    return {"data": v["data"][tuple(sl)], "coords": coords, "dims": v["dims"]}, sl


def bin_view(v, binning):
    """``sim.coarsen(binning, boundary="trim").mean().astype(sim.dtype)``
    (registration.py:1732-1743): window means of data and of coordinates;
    integer data average in float64 and are truncated by ``astype``."""
    data = v["data"]
    dims = v["dims"]
    b = [int(binning.get(d, 1)) for d in dims]
    n = [data.shape[i] // b[i] for i in range(len(dims))]
    trimmed = data[tuple(slice(0, n[i] * b[i]) for i in range(len(dims)))]
    shp = []
    for i in range(len(dims)):
        shp += [n[i], b[i]]
    win = trimmed.reshape(shp)
    axes = tuple(range(1, 2 * len(dims), 2))
    if np.issubdtype(data.dtype, np.floating):
        out = np.nanmean(win.astype(np.float64), axis=axes)
    else:
        out = np.mean(win, axis=axes, dtype=np.float64)
    coords = {d: v["coords"][d][: n[i] * b[i]].reshape(n[i], b[i]).mean(axis=1) for i, d in enumerate(dims)}
    return {"data": out.astype(data.dtype), "coords": coords, "dims": dims}


def optimal_registration_binning(v1, v2, max_total_pixels_per_stack=400**3):
    """registration.py:114-191."""
    dims = v1["dims"]
    ndim = len(dims)
    sp = [spacing_of(v1), spacing_of(v2)]
    overlap = {d: max(v1["data"].shape[i], v2["data"].shape[i]) for i, d in enumerate(dims)}
    binning = {d: 1 for d in dims}
    spacings = sp
    while np.prod([overlap[d] / binning[d] for d in dims]) >= max_total_pixels_per_stack:
        dim_to_bin = int(np.argmin([min(spacings[k][d] for k in range(2)) for d in dims]))
        if ndim == 3 and dim_to_bin == 0:
            binning["z"] += 1
        else:
            for d in ["x", "y"]:
                binning[d] += 1
        spacings = [{d: sp[k][d] * binning[d] for d in dims} for k in range(2)]
    return binning


# --------------------------------------------------------------------------
# overlap polytope (mv_graph.py)
# --------------------------------------------------------------------------


def _apply(affine, pts):
    """``transformation.transform_pts`` (transformation.py:151-161)."""
    pts = np.asarray(pts, dtype=float)
    h = np.concatenate([pts, np.ones((pts.shape[0], 1))], axis=1)
    return np.array([np.dot(np.asarray(affine), p) for p in h])[:, :-1]


def faces_of(sp):
    """mv_graph.py:386-420: the 2*ndim faces of the (transformed) pixel-centre box."""
    dims = SPATIAL_DIMS[-len(sp["origin"]):]
    ndim = len(dims)
    gv = np.array(list(np.ndindex(*([2] * ndim))))
    faces = []
    for iax in range(ndim):
        for lface in (0, 1):
            faces.append(gv[np.where(gv[:, iax] == lface)[0]])
    faces = np.array(faces)
    faces = faces * (np.array([sp["shape"][d] for d in dims]) - 1) * np.array(
        [sp["spacing"][d] for d in dims]
    ) + np.array([sp["origin"][d] for d in dims])
    if "transform" in sp:
        shp = faces.shape
        flat = faces.reshape(-1, ndim)
        flat = np.dot(sp["transform"], np.hstack([flat, np.ones((flat.shape[0], 1))]).T).T[:, :-1]
        faces = flat.reshape(shp)
    return faces


def center_of(sp):
    """mv_graph.py:475-493."""
    dims = SPATIAL_DIMS[-len(sp["origin"]):]
    c = np.array([sp["origin"][d] + sp["spacing"][d] * (sp["shape"][d] - 1) / 2 for d in dims])
    if "transform" in sp:
        c = np.matmul(np.array(sp["transform"]), np.concatenate([c, np.ones(1)]))[: len(dims)]
    return c


def halfspace_equations(sp):
    """mv_graph.py:183-218: rows ``[n, c]`` with ``n.x + c <= 0`` inside."""
    ndim = len(sp["origin"])
    faces = faces_of(sp)
    center = center_of(sp)
    eqs = []
    for f in faces:
        if ndim == 2:
            n = np.array([-(f[1][1] - f[0][1]), f[1][0] - f[0][0]])
        else:
            n = np.cross(f[1] - f[0], f[2] - f[0])
        n = n / np.linalg.norm(n)
        c = -np.dot(n, f[0])
        if np.dot(n, center) + c > 0:
            n = -n
        c = -np.dot(n, f[0])
        eqs.append(np.concatenate([n, [c]]))
    return np.array(eqs)


def overlap_polytope(sp1, sp2):
    """mv_graph.py:301-338: Chebyshev centre by ``linprog`` as the interior point,
    then qhull's half-space intersection.  Returns ``(volume, vertices)``."""
    eqs = np.concatenate([halfspace_equations(sp1), halfspace_equations(sp2)])
    norm = np.reshape(np.linalg.norm(eqs[:, :-1], axis=1), (eqs.shape[0], 1))
    c = np.zeros((eqs.shape[1],))
    c[-1] = -1
    res = linprog(c, A_ub=np.hstack((eqs[:, :-1], norm)), b_ub=-eqs[:, -1:], bounds=(None, None))
    hs = HalfspaceIntersection(eqs, res.x[:-1])
    verts = hs.intersections
    return ConvexHull(verts).volume, verts


def overlap_bboxes(v1, v2, affine1, affine2, overlap_tolerance=None, intrinsic=True):
    """``_get_overlap_bboxes`` (registration.py:194-277): lower / upper corner of the
    overlap polytope, per view in its own intrinsic (physical, pre-transform)
    coordinates (``output_transform_key=None``) or in the world system."""
    sps = [stack_props(v, a, overlap_tolerance) for v, a in ((v1, affine1), (v2, affine2))]
    vol, corners = overlap_polytope(*sps)
    if intrinsic:
        target = [_apply(np.linalg.inv(a), corners) for a in (affine1, affine2)]
    else:
        target = [corners, corners]
    lowers = [np.min(t, axis=0) for t in target]
    uppers = [np.max(t, axis=0) for t in target]
    return lowers, uppers, vol


# --------------------------------------------------------------------------
# common pixel grid (registration.py:280-350) and the way back (:1382-1474)
# --------------------------------------------------------------------------


def to_intrinsic_grid(v1, v2, affine1, affine2, lowers, uppers):
    """``sims_to_intrinsic_coord_system``: both crops resampled (order 1, NaN outside)
    onto ``origin = lowers[0]``, ``spacing = max(spacings)``,
    ``shape = floor((uppers[0] - lowers[0]) / spacing + 1)``; the moving view through
    ``inv(affine2) @ affine1``.  Returns ``(fixed, moving, grid)`` float32."""
    dims = v1["dims"]
    spacing = np.max([[spacing_of(v)[d] for d in dims] for v in (v1, v2)], axis=0)
    transf = np.matmul(np.linalg.inv(affine2), affine1)
    shape = np.floor(np.array(uppers[0] - lowers[0]) / spacing + 1).astype(np.uint64)
    grid = {
        "origin": {d: lowers[0][i] for i, d in enumerate(dims)},
        "spacing": {d: spacing[i] for i, d in enumerate(dims)},
        "shape": {d: int(shape[i]) for i, d in enumerate(dims)},
    }
    out = []
    for v, p in ((v1, None), (v2, transf)):
        view = {"data": v["data"].astype(np.float32), "origin": origin_of(v), "spacing": spacing_of(v)}
        out.append(np.asarray(ofusion.transform_view(view, p, grid, order=1, cval=np.nan), dtype=np.float32))
    return out[0], out[1], grid


def affine_from_intrinsic_affine(data_affine, grid, affine_key):
    """``get_affine_from_intrinsic_affine`` (registration.py:1382-1474) for the
    pixel-space branch, where both images live on ``grid`` and carry the fixed
    view's transform: ``M_W = (A T S) M_D (A T S)^-1``."""
    dims = SPATIAL_DIMS[-(np.asarray(data_affine).shape[0] - 1):]
    T = oreg.affine_from_translation([grid["origin"][d] for d in dims])
    S = np.diag([grid["spacing"][d] for d in dims] + [1])
    D_to_W = np.matmul(np.array(affine_key), np.matmul(T, S))
    return np.matmul(D_to_W, np.matmul(data_affine, np.linalg.inv(D_to_W)))


def prepare_pair(view1, view2, affine1, affine2, overlap_tolerance=None, registration_binning=None):
    """registration.py:1732-1968 (pixel-space branch): bin, overlap boxes, crop with one
    pixel of margin, resample onto the fixed view's grid."""
    v1, v2 = with_coords(view1), with_coords(view2)
    dims = v1["dims"]
    if overlap_tolerance is None:
        overlap_tolerance = {d: 0.0 for d in dims}
    elif not isinstance(overlap_tolerance, dict):
        overlap_tolerance = {d: float(overlap_tolerance) for d in dims}
    else:
        overlap_tolerance = {d: float(overlap_tolerance.get(d, 0.0)) for d in dims}
    if registration_binning is None:
        registration_binning = optimal_registration_binning(v1, v2)
    if max(registration_binning.values()) > 1:
        v1, v2 = bin_view(v1, registration_binning), bin_view(v2, registration_binning)
    lowers, uppers, _ = overlap_bboxes(v1, v2, affine1, affine2, overlap_tolerance)
    tol = 1e-6
    crops = []
    for k, v in enumerate((v1, v2)):
        sp = spacing_of(v)
        lo = [lowers[k][i] - tol - sp[d] for i, d in enumerate(dims)]
        hi = [uppers[k][i] + tol + sp[d] for i, d in enumerate(dims)]
        crops.append(sel_coords(v, lo, hi)[0])
    fixed, moving, grid = to_intrinsic_grid(crops[0], crops[1], affine1, affine2, lowers, uppers)
    return {"fixed": fixed, "moving": moving, "grid": grid, "lowers": lowers, "uppers": uppers,
            "binning": registration_binning}


def register_pair(view1, view2, affine1, affine2, overlap_tolerance=None, registration_binning=None,
                  pairwise_reg_func=None, pairwise_reg_func_kwargs=None):
    """``register_pair_of_msims`` (registration.py:1547-2058) for an image-data hook in
    pixel space: returns ``{"transform", "quality", "bbox"}`` with the physical
    transform (fixed world -> moving world) and the world-space overlap box of the
    un-binned views (:2038-2056)."""
    affine1, affine2 = np.asarray(affine1, dtype=float), np.asarray(affine2, dtype=float)
    prep = prepare_pair(view1, view2, affine1, affine2, overlap_tolerance, registration_binning)
    func = pairwise_reg_func or oreg.phase_correlation_registration
    res = oreg.dispatch_pairwise_reg_func(func, prep["fixed"], prep["moving"], **(pairwise_reg_func_kwargs or {}))
    transform = affine_from_intrinsic_affine(np.array(res["affine_matrix"]), prep["grid"], affine1)
    v1, v2 = with_coords(view1), with_coords(view2)
    dims = v1["dims"]
    if overlap_tolerance is None or not isinstance(overlap_tolerance, dict):
        tolv = 0.0 if overlap_tolerance is None else float(overlap_tolerance)
        overlap_tolerance = {d: tolv for d in dims}
    else:
        overlap_tolerance = {d: float(overlap_tolerance.get(d, 0.0)) for d in dims}
    lo, hi, _ = overlap_bboxes(v1, v2, affine1, affine2, overlap_tolerance, intrinsic=False)
    return {"transform": transform, "quality": res["quality"], "bbox": np.array([lo[0], hi[0]]),
            "prepared": prep, "affine_matrix": np.array(res["affine_matrix"])}
