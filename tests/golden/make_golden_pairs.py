"""Golden fixtures for pair preparation, produced by the REFERENCE'S OWN code.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_pairs.py

Runs, on the seeded inputs of ``cases.pair_cases()``, the reference's real
``_get_overlap_bboxes`` (registration.py:194-277, with the real ``mv_graph``
half-space geometry), ``sims_to_intrinsic_coord_system`` (:280-350, real
``transform_sim``), ``phase_correlation_registration`` (:353-565) and
``get_affine_from_intrinsic_affine`` (:1382-1474) in the order
``register_pair_of_msims`` (:1732-2056) chains them, and stores lowers / uppers /
the two crops / the pixel affine / the physical transform / the world bbox in
``tests/golden/pairs_golden.npz``.  The xarray layer (coordinate selection,
coarsening) is restated in ``_ref_loader`` -- see there.
"""

from __future__ import annotations

import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import cases  # noqa: E402
from _ref_loader import load_reference_pairs  # noqa: E402

DIMS = ["z", "y", "x"]
KEY = "stage"


def coarsen(sim, binning, FakeSim):
    """xarray ``coarsen(boundary='trim').mean().astype(dtype)`` (registration.py:1732-1743)."""
    dims = list(sim.dims)
    b = [int(binning[d]) for d in dims]
    n = [sim.data.shape[i] // b[i] for i in range(len(dims))]
    t = sim.data[tuple(slice(0, n[i] * b[i]) for i in range(len(dims)))]
    shp = []
    for i in range(len(dims)):
        shp += [n[i], b[i]]
    axes = tuple(range(1, 2 * len(dims), 2))
    m = np.mean(t.reshape(shp), axis=axes, dtype=np.float64).astype(sim.data.dtype)
    origin, spacing, cs = {}, {}, {}
    for i, d in enumerate(dims):
        c = (sim.origin[d] + sim.spacing[d] * np.arange(sim.data.shape[i], dtype=float))[: n[i] * b[i]]
        c = c.reshape(n[i], b[i]).mean(axis=1)
        origin[d], spacing[d], cs[d] = c[0], c[1] - c[0], c
    return FakeSim(m, dims, origin, spacing, attrs=dict(sim.attrs), coords=cs)


def main():
    ref = load_reference_pairs()
    reg, FakeSim, fa = ref.registration, ref.FakeSim, ref.fake_affine
    out = {}
    for name, case in cases.pair_cases(extra=True).items():
        ndim = case["views"][0]["data"].ndim
        dims = DIMS[-ndim:]
        sims = [
            FakeSim(v["data"], dims, v["origin"], v["spacing"], attrs={"transforms": {KEY: fa(a)}})
            for v, a in zip(case["views"], case["affines"])
        ]
        kw = case["kwargs"]
        tolv = kw.get("overlap_tolerance")
        if isinstance(tolv, dict):  # registration.py:1624-1637
            tol_d = {d: float(tolv.get(d, 0.0)) for d in dims}
        else:
            tol_d = {d: 0.0 if tolv is None else float(tolv) for d in dims}
        binning = kw["registration_binning"]
        b_sims = [coarsen(s, binning, FakeSim) for s in sims] if max(binning.values()) > 1 else sims

        ov = reg._get_overlap_bboxes(b_sims[0], b_sims[1], input_transform_key=KEY,
                                     output_transform_key=None, overlap_tolerance=tol_d)
        lowers, uppers = ov["lowers"], ov["uppers"]
        sp = [ref.registration.spatial_image_utils.get_spacing_from_sim(s) for s in b_sims]
        tol = 1e-6
        crops = [
            ref.registration.spatial_image_utils.sim_sel_coords(
                s,
                {d: slice(lowers[k][i] - tol - sp[k][d], uppers[k][i] + tol + sp[k][d]) for i, d in enumerate(dims)},
            )
            for k, s in enumerate(b_sims)
        ]
        ps = reg.sims_to_intrinsic_coord_system(crops[0], crops[1], transform_key=KEY, overlap_bboxes=(lowers, uppers))
        fixed, moving = np.asarray(ps[0].data), np.asarray(ps[1].data)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            res = reg.phase_correlation_registration(ps[0], ps[1])
        affine = np.array(res["affine_matrix"])
        phys = reg.get_affine_from_intrinsic_affine(
            data_affine=affine, sim_fixed=ps[0], sim_moving=ps[1], transform_key_fixed=KEY, transform_key_moving=KEY
        )
        ovw = reg._get_overlap_bboxes(sims[0], sims[1], input_transform_key=KEY, output_transform_key=KEY,
                                      overlap_tolerance=tol_d)
        out[name + "/lowers"] = np.array(lowers)
        out[name + "/uppers"] = np.array(uppers)
        out[name + "/fixed"] = fixed.astype(np.float32)
        out[name + "/moving"] = moving.astype(np.float32)
        out[name + "/grid_origin"] = np.array([ps[0].origin[d] for d in dims])
        out[name + "/grid_spacing"] = np.array([ps[0].spacing[d] for d in dims])
        out[name + "/affine_matrix"] = affine
        out[name + "/quality"] = np.array(res["quality"], dtype=float)
        out[name + "/transform"] = np.asarray(phys, dtype=float)
        out[name + "/bbox"] = np.array([ovw["lowers"][0], ovw["uppers"][0]])
        print(name, fixed.shape, "t_px", affine[:ndim, ndim], "q", float(res["quality"]),
              "nan frac", float(np.isnan(moving).mean()))
    path = os.path.join(HERE, "pairs_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
