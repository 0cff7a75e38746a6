"""Call-to-call variation of the 3-D registration from tiles (64 C3 face pairs), per stage."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from multiview_stitcher_b200 import pairs as pairs_mod, registration, synthetic
grid = bench.C3["grid"]
views, stage, true = synthetic.make_grid(grid, bench.C3["tile"], bench.C3["overlap"], np.uint16, jitter=2, seed=bench.SEED, subpixel=True)
pairs = bench._c3_pairs(grid)
plans = {}
pplan = pairs_mod.PairPlan(views, stage, pairs, registration_binning={"z": 1, "y": 1, "x": 1})
# instrument the stages of PhaseCorrPlan
stamps = []
def wrap(cls, name):
    f = getattr(cls, name)
    def g(self, *a, **k):
        t0 = time.perf_counter(); r = f(self, *a, **k); stamps.append((name, self.shape, (time.perf_counter() - t0) * 1e3)); return r
    setattr(cls, name, g)
for n in ("load_pairs", "correlate", "candidate_ssim", "spearman_batch", "candidate_stats"):
    wrap(registration.PhaseCorrPlan, n)
import gc
orig_prepare = pairs_mod.PairPlan.prepare
def timed_prepare(self, *a, **k):
    t0 = time.perf_counter(); r = orig_prepare(self, *a, **k); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    stamps.append(("prepare_host", (0,), (t1 - t0) * 1e3)); stamps.append(("prepare_gpu", (0,), (t2 - t1) * 1e3)); return r
pairs_mod.PairPlan.prepare = timed_prepare
if os.environ.get("NOGC"): gc.disable()
for i in range(12):
    stamps.clear()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    pairs_mod.register_views(views, plan=pplan, pc_plans=plans)
    torch.cuda.synchronize()
    tot = (time.perf_counter() - t0) * 1e3
    print(f"call {i}: {tot:.0f} ms  " + "  ".join(f"{n}{tuple(s)[-1]}:{t:.0f}" for n, s, t in stamps))
