"""Generates tests/golden/dct_golden.npz by running the REFERENCE's own
``weights.content_based_dct`` (weights.py:77-290, loaded through _ref_loader) on seeded
stacks.  Run in this container only (needs /root/reference):

    python tests/golden/make_golden_dct.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import _ref_loader  # noqa: E402


def dct_cases():
    """name -> (stack (V, *spatial) float32 with NaN outside, kwargs)."""
    rng = np.random.default_rng(11)
    out = {}

    def smooth(shape, k):
        from scipy.ndimage import gaussian_filter

        a = gaussian_filter(rng.random(shape).astype(np.float32), k)
        return ((a - a.min()) / (a.max() - a.min()) * 1000).astype(np.float32)

    v = np.stack([smooth((70, 100), s) for s in (1.0, 2.5, 0.5)])
    v[0, :20, :33] = np.nan
    v[2, 50:, 60:] = np.nan
    out["2d_default16"] = (v, dict(dct_size=16))
    out["2d_dict_exp2"] = (v, dict(dct_size={"y": 32, "x": 16}, exponent=2.0, output_chunksize={"y": 64, "x": 64}))
    out["2d_no_otf"] = (v[:2], dict(dct_size=8, otf_support_fraction=None))
    w = np.stack([smooth((20, 45, 40), s) for s in (0.7, 1.5)])
    w[1, :, :12, :] = np.nan
    w[0, 15:, 30:, 25:] = np.nan
    out["3d_default"] = (w, dict(dct_size=16, otf_support_fraction=0.5))
    out["3d_aniso"] = (w, dict(dct_size={"z": 8, "y": 32, "x": 16}, otf_support_fraction=0.75))
    return out


if __name__ == "__main__":
    ref = _ref_loader.load_reference()
    arrays = {}
    for name, (stack, kw) in dct_cases().items():
        arrays[name] = ref.weights.content_based_dct(stack.copy(), **kw)
        print(name, arrays[name].shape, float(np.nanmin(arrays[name])), float(np.nanmax(arrays[name])))
    np.savez_compressed(os.path.join(HERE, "dct_golden.npz"), **arrays)
