"""Generates tests/golden/deconv_golden.npz by running the REFERENCE's own
``fusion.mv_deconv.multi_view_deconvolution`` (fusion/mv_deconv.py:251-500, loaded through
_ref_loader) on seeded stacks.  Run in this container only (needs /root/reference):

    python tests/golden/make_golden_deconv.py
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import _ref_loader  # noqa: E402


def deconv_cases():
    """name -> (views (V, *spatial) float32 NaN outside, normalised blending weights, kwargs)."""
    from scipy.ndimage import gaussian_filter

    rng = np.random.default_rng(21)
    out = {}

    def stack(shape, n_views, blur):
        truth = np.zeros(shape, np.float32)
        idx = tuple(rng.integers(3, s - 3, 25) for s in shape)
        truth[idx] = rng.random(25).astype(np.float32) * 500 + 200
        truth = gaussian_filter(truth, 1.0) * 20 + 5
        views, weights = [], []
        for v in range(n_views):
            sig = [blur * (2.0 if d == v % len(shape) else 0.8) for d in range(len(shape))]
            img = gaussian_filter(truth, sig).astype(np.float32) + rng.random(shape).astype(np.float32)
            w = np.ones(shape, np.float32)
            sl = [slice(None)] * len(shape)
            ax = (v + 1) % len(shape)
            cut = shape[ax] // 4
            sl[ax] = slice(0, cut) if v % 2 == 0 else slice(shape[ax] - cut, None)
            img[tuple(sl)] = np.nan
            w[tuple(sl)] = 0
            ramp = np.linspace(0, 1, 6, dtype=np.float32)
            sl2 = [slice(None)] * len(shape)
            sl2[ax] = slice(cut, cut + 6) if v % 2 == 0 else slice(shape[ax] - cut - 6, shape[ax] - cut)
            shp = [1] * len(shape)
            shp[ax] = 6
            w[tuple(sl2)] *= (ramp if v % 2 == 0 else ramp[::-1]).reshape(shp)
            views.append(img)
            weights.append(w)
        views, weights = np.stack(views), np.stack(weights)
        s = weights.sum(0)
        s[s == 0] = 1
        return views, (weights / s).astype(np.float32)

    v2, w2 = stack((48, 60), 2, 1.2)
    out["2d_default"] = (v2, w2, dict(n_iterations=4))
    out["2d_opt1_reg"] = (v2, w2, dict(n_iterations=3, psf_type="OPTIMIZATION_I", lambda_reg=0.006,
                                       psfs=[_psf((7, 5), (1.4, 0.9)), _psf((5, 7), (0.9, 1.4))]))
    out["2d_independent_erode"] = (v2, w2, dict(n_iterations=2, psf_type="INDEPENDENT", sample_boundary_erosion_px=2))
    v3, w3 = stack((18, 26, 30), 3, 1.0)
    out["3d_spacing"] = (v3, w3, dict(n_iterations=2, output_spacing={"z": 0.8, "y": 0.3, "x": 0.3}))
    out["3d_opt2"] = (v3, w3, dict(n_iterations=2, psf_type="OPTIMIZATION_II", psfs=[_psf((5, 5, 5), (1.0, 0.8, 0.8))] * 3))
    return out


def _psf(shape, sigma):
    from scipy.ndimage import gaussian_filter

    p = np.zeros(shape, np.float32)
    p[tuple(s // 2 for s in shape)] = 1
    return gaussian_filter(p, sigma)


if __name__ == "__main__":
    _ref_loader.load_reference()
    mvd = importlib.import_module("multiview_stitcher.fusion.mv_deconv")
    arrays = {}
    for name, (views, weights, kw) in deconv_cases().items():
        arrays[name] = mvd.multi_view_deconvolution(views.copy(), weights.copy(), **kw)
        print(name, arrays[name].shape, float(arrays[name].min()), float(arrays[name].max()))
    arrays["overlap_default"] = np.array(mvd.multi_view_deconvolution.required_overlap({}))
    arrays["overlap_spacing"] = np.array(mvd.multi_view_deconvolution.required_overlap({"output_spacing": {"z": 0.8, "y": 0.3, "x": 0.3}}))
    np.savez_compressed(os.path.join(HERE, "deconv_golden.npz"), **arrays)
