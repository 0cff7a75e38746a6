#!/usr/bin/env python
"""A/B of the Bluestein kernel length on C2's crop shapes: the phase-correlation stage
(`mvs_pc_correlate`, 20 pairs per shape) with the default 1024-point chirp-z kernel and with
MVS_BLUESTEIN_SMOOTH=1 (640 = 20*4*4*2 points for 257 <= n <= 320).  The switch is read once
per process, so each arm runs in its own interpreter.  Prints one JSON line: per arm the
stage time and whether integer peaks and upsampled-DFT samples agree between the arms.

    python scripts/bench_bluestein.py
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

ARM = r"""
import json, sys, numpy as np, torch
sys.path.insert(0, %r)
from multiview_stitcher_b200 import registration, synthetic
out = {}
peaks_all, up_all = [], []
for shape in ((2048, 307), (307, 2048)):
    n = 20
    fixed = [synthetic.make_tile(shape, (1000 * i, 77 * i), np.float32, seed=3) for i in range(n)]
    moving = [synthetic.make_tile(shape, (1000 * i + 2, 77 * i - 1), np.float32, seed=3) for i in range(n)]
    plan = registration.PhaseCorrPlan(shape, n, 10)
    plan.load_pairs(fixed, moving)
    for _ in range(3):
        peaks, up = plan.correlate()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        plan.correlate()
    e1.record()
    torch.cuda.synchronize()
    out["x".join(map(str, shape))] = e0.elapsed_time(e1) / 5
    peaks_all.append(peaks.tolist())
    up_all.append([float(np.abs(up).max()), float(np.abs(up).sum())])
    plan.close()
print(json.dumps({"ms": out, "peaks": peaks_all, "updft": up_all}))
""" % ROOT


def run(smooth):
    env = dict(os.environ)
    env.pop("MVS_BLUESTEIN_SMOOTH", None)
    if smooth:
        env["MVS_BLUESTEIN_SMOOTH"] = "1"
    r = subprocess.run([sys.executable, "-c", ARM], env=env, capture_output=True, text=True, timeout=300)
    if r.returncode != 0:
        return {"error": r.stderr[-800:]}
    return json.loads(r.stdout.strip().splitlines()[-1])


def main():
    a, b = run(False), run(True)
    res = {"m1024": a.get("ms", a), "m640": b.get("ms", b)}
    if "peaks" in a and "peaks" in b:
        res["peaks_equal"] = a["peaks"] == b["peaks"]
        res["updft_rel_diff"] = max(abs(x[1] - y[1]) / max(abs(x[1]), 1e-30) for x, y in zip(a["updft"], b["updft"]))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
