"""Generate golden fixtures by running the REFERENCE'S OWN code.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Writes ``tests/golden/fusion_golden.npz`` and
``tests/golden/registration_golden.npz``.  See ``_ref_loader.py`` for how the
reference's hot-path modules are imported without their data-model
dependencies.  Inputs come from ``cases.py`` (seeded) and are NOT stored.
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
from _ref_loader import load_reference  # noqa: E402

DIMS = ["z", "y", "x"]


def main():
    ref = load_reference()
    fc, w = ref.fusion_core, ref.weights
    FakeSim = ref.FakeSim

    out = {}
    for name, case in cases.fusion_cases().items():
        views, params, kwargs = case["views"], case["params"], dict(case["kwargs"])
        ndim = views[0]["data"].ndim
        dims = DIMS[-ndim:]
        sims = [FakeSim(v["data"], dims, v["origin"], v["spacing"]) for v in views]
        bbs = [
            {
                "origin": dict(v["origin"]),
                "spacing": dict(v["spacing"]),
                "shape": dict(zip(dims, v["data"].shape)),
            }
            for v in views
        ]
        # reference's own union stack properties (fusion/_core.py:1821-1992)
        sp = fc.calc_stack_properties_from_view_properties_and_params(
            bbs, params, spacing=views[0]["spacing"], mode="union"
        )
        osp = {
            k: {d: (int(v[i]) if k == "shape" else float(v[i])) for i, d in enumerate(dims)}
            for k, v in sp.items()
        }
        fusion_func = getattr(fc, kwargs.pop("fusion_func", "weighted_average_fusion"))
        weights_func = kwargs.pop("weights_func", None)
        if weights_func is not None:
            weights_func = getattr(w, weights_func)

        captured = {}

        if fusion_func is fc.weighted_average_fusion:

            def ff(transformed_views, blending_weights, fusion_weights=None):
                captured["views"] = transformed_views.copy()
                captured["bw"] = blending_weights.copy()
                if fusion_weights is not None:
                    captured["fw"] = np.asarray(fusion_weights).copy()
                return fc.weighted_average_fusion(
                    transformed_views, blending_weights, fusion_weights
                )

        else:
            ff = fusion_func

        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            fused = fc.fuse_np(
                sims,
                params,
                osp,
                fusion_func=ff,
                weights_func=weights_func,
                full_view_bbs=bbs,
                **kwargs,
            )
        out[name + "/fused"] = fused
        out[name + "/origin"] = np.array([osp["origin"][d] for d in dims])
        out[name + "/shape"] = np.array([osp["shape"][d] for d in dims])
        out[name + "/spacing"] = np.array([osp["spacing"][d] for d in dims])
        for k, v in captured.items():
            out[name + "/" + k] = v.astype(np.float32)
        print(name, fused.shape, fused.dtype, float(fused.mean()))

        # a sub-chunk with halo + trim through the reference's fuse_np, for the
        # content-based cases (halo semantics, fusion/_core.py:1687-1711)
        if weights_func is not None:
            ov = 4
            start = {d: 5 for d in dims}
            shape = {d: min(20, osp["shape"][d] - 8) for d in dims}
            hbb = {
                "origin": {
                    d: osp["origin"][d] + (start[d] - ov) * osp["spacing"][d] for d in dims
                },
                "spacing": osp["spacing"],
                "shape": {d: shape[d] + 2 * ov for d in dims},
            }
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                sub = fc.fuse_np(
                    sims,
                    params,
                    hbb,
                    fusion_func=fusion_func,
                    weights_func=weights_func,
                    full_view_bbs=bbs,
                    trim_overlap_in_pixels=ov,
                    **kwargs,
                )
            out[name + "/sub_fused"] = sub
            out[name + "/sub_start"] = np.array([start[d] for d in dims])
            out[name + "/sub_shape"] = np.array([shape[d] for d in dims])
            out[name + "/sub_overlap"] = np.array(ov)

    np.savez_compressed(os.path.join(HERE, "fusion_golden.npz"), **out)

    # ---------------- registration -------------------------------------
    reg = ref.registration

    class Wrap:
        def __init__(self, a):
            self.data = a

    rout = {}
    for name, (f, m, tr) in cases.registration_cases().items():
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            res = reg.phase_correlation_registration(Wrap(f), Wrap(m))
        rout[name + "/affine"] = np.asarray(res["affine_matrix"], dtype=np.float64)
        rout[name + "/quality"] = np.array(res["quality"], dtype=np.float64)
        print(name, np.asarray(res["affine_matrix"])[:-1, -1], res["quality"], "gt", tr)
    np.savez_compressed(os.path.join(HERE, "registration_golden.npz"), **rout)


if __name__ == "__main__":
    main()
