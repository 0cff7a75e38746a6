#!/usr/bin/env python
"""Benchmark of the registration-and-fusion hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--only c2,c3,c4,c5] [--no-cpu]

Prints ONE JSON line (DESIGN.md "Measurement").  All inputs are tiles of the
band-limited analytic field at FRACTIONAL positions (jitter ~ U(-2, 2) px,
SURVEY.md 8d), so every interpolation fraction is non-zero.

Headline (`value`, `roofline`, `e2e`, `cpu_baseline`): BASELINE config 1 ("C2"),
5x5 grid of 2048x2048 float32 tiles, 15 % overlap -- one step = the whole stack
fused with cosine-edge blending (and, reported beside it, all 40 overlap pairs
registered by phase correlation).

* ``value``     fused Mvoxels/s, tiles resident in HBM, CUDA-event timed.
* ``e2e``       same metric through the reference's hook C (`batch_func`:
                `BatchFuser.__call__` on the partial `fuse()` builds, fusion/_core.py:1133-1141)
                with PAGEABLE numpy tiles in and a numpy destination array out; staging
                through pinned memory, H2D, fusion and D2H are all inside the timing.
* ``roofline``  algorithmic bytes of the fused resample-blend kernel / its measured
                duration vs the measured HBM peak.
* ``cpu_baseline`` the oracle (numpy/scipy restatement of the reference's path) on a
                bounded sample of the same workload on the host cores.
* ``configs``   the other BASELINE configs: C3 (4x4x2 grid of 256x512x512 uint16, blend /
                content-weighted fusion, 64 pairs), C4 (4 affine views of 512x1024x1024),
                C5 (8x8 grid of 512x2048x2048 uint16, one tile row per GPU).

N > 1 (torchrun, one rank per GPU): `value` stays C2 with one replica per rank (chunks
and pairs are independent units -> "weak").  The sharded jobs are in ``configs``:
C3 as ONE job strong-scaled over the ranks (pairs round-robin, chunk bands with tile
replication, and the tile-partitioned run whose border boxes cross NVLink as NCCL
send/recv of partial (sum w*v, sum w)); C5 weak-scaled: every rank holds one tile row
of the 8x8 grid, N ranks fuse N rows, only the row-overlap boxes are exchanged.
``--impl reference`` times the CPU path alone (the reference is pure Python on scipy
and cannot be imported in this image; the oracle port makes the same library calls).
"""

from __future__ import annotations

import argparse
import functools
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOAD = {
    "workload": "C2: 5x5 grid of 2-D tiles 2048x2048 float32, 15% overlap (307 px), sub-pixel tile positions "
    "(jitter U(-2,2) px), phase-correlation registration + cosine-edge weighted-average fusion",
    "grid": [5, 5],
    "tile": [2048, 2048],
    "overlap_px": 307,
    "tile_dtype": "float32",
    "output_chunks": "2048x2048 (reference default)",
    "interpolation_order": 1,
    "jitter": "U(-2, 2) px per axis on a 1/64 px grid (fractional: every lerp fraction non-zero)",
    "l2": "inputs+outputs (744 MB/step) exceed the 126 MB L2; no explicit flush",
}
GRID, TILE, OVERLAP = (5, 5), (2048, 2048), (307, 307)
C3 = {"grid": (2, 4, 4), "tile": (256, 512, 512), "overlap": (26, 51, 51)}
C5 = {"grid": (1, 8, 8), "tile": (512, 2048, 2048), "overlap": (0, 205, 205)}
METRIC = "fused_Mvoxels_per_sec"
UNIT = "Mvoxel/s"
SEED = 0


def _ncu_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture."""
    for name in ("r02_traffic.json", "r01c_traffic.json"):
        p = os.path.join(ROOT, "profiles", name)
        try:
            with open(p) as f:
                return int(json.load(f)["traffic_bytes_per_launch"]), f"profiles/{name}"
        except Exception:
            continue
    return None, None


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = (
        "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
        "clocks_event_reasons.sw_power_cap"
    )

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(
                    ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                    capture_output=True, text=True, timeout=5,
                ).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows if len(r) > 2 + i)]
        return {
            "sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": max(mx) if mx else None,
            "reasons": reasons,
            "samples": len(self.rows),
        }


# ----------------------------------------------------------------------------
# CPU arm: the oracle on a bounded sample.  Nothing here touches the CUDA library:
# tiles come from the numpy mirror of the generator (synthetic.field_tile_host).
# ----------------------------------------------------------------------------


def _host_c2_views(seed=SEED):
    """C2's 25 tiles on the host (numpy mirror of the analytic field; equal to the GPU
    generator's tiles up to float32 rounding)."""
    from multiview_stitcher_b200 import synthetic

    true, stage, idx = synthetic.grid_layout(GRID, TILE, OVERLAP, jitter=2, seed=seed, subpixel=True)
    views, params = [], []
    for k, (s_org, t_org) in enumerate(zip(stage, true)):
        data = synthetic.field_tile_host(TILE, t_org, np.float32, seed, tile_id=k)
        views.append({"data": data, "origin": {"y": float(s_org[0]), "x": float(s_org[1])}, "spacing": {"y": 1.0, "x": 1.0}})
        p = np.eye(3)
        p[:2, 2] = t_org - s_org
        params.append(p)
    return views, params


def _cpu_fuse_chunk(args):
    from oracle import fusion as of

    views, params, bbs, hbb = args
    return of.fuse_np(views, params, hbb, full_view_bbs=bbs).shape


def cpu_fusion_sample(views, params, n_jobs=None):
    """Oracle fuse_np over ALL of C2's 25 output chunks (2048^2, the reference's default
    chunking), one chunk per worker process (joblib), each handed only the views that
    touch it.  Returns (Mvoxel/s, cores, sample description)."""
    from joblib import Parallel, delayed

    from oracle import fusion as of

    bbs = [of.view_bb(v) for v in views]
    osp = of.calc_stack_properties(bbs, params, views[0]["spacing"])
    cores = n_jobs or os.cpu_count() or 1
    chunks = of.chunk_bbs(osp, {"y": 2048, "x": 2048})
    jobs, vox = [], 0
    for cbb, _ in chunks:
        sel = [k for k in range(len(views)) if of._view_touches(bbs[k], params[k], cbb, 2)]
        jobs.append(([views[k] for k in sel], [params[k] for k in sel], [bbs[k] for k in sel], cbb))
        vox += int(np.prod([cbb["shape"][d] for d in "yx"]))
    t0 = time.perf_counter()
    Parallel(n_jobs=cores)(delayed(_cpu_fuse_chunk)(j) for j in jobs)
    dt = time.perf_counter() - t0
    return vox / dt / 1e6, cores, f"all {len(chunks)} output chunks (2048x2048) of C2, oracle fuse_np, joblib x{cores}"


def c2_pairs():
    """C2's 40 face-adjacent tile pairs as (index a, index b, axis)."""
    pairs = []
    for iy in range(GRID[0]):
        for ix in range(GRID[1]):
            a = iy * GRID[1] + ix
            if ix + 1 < GRID[1]:
                pairs.append((a, a + 1, 1))
            if iy + 1 < GRID[0]:
                pairs.append((a, a + GRID[1], 0))
    return pairs


def pair_crops(tiles, pairs):
    """Overlap crops (stage geometry): fixed = trailing strip of tile a, moving =
    leading strip of tile b.  Works for numpy arrays and torch tensors."""
    fixed, moving = [], []
    for a, b, axis in pairs:
        ov = OVERLAP[axis]
        if axis == 1:
            fixed.append(tiles[a][:, TILE[1] - ov :])
            moving.append(tiles[b][:, :ov])
        else:
            fixed.append(tiles[a][TILE[0] - ov :, :])
            moving.append(tiles[b][:ov, :])
    return fixed, moving


def _cpu_register_pair(args):
    from oracle import registration as oreg

    f, m = args
    return oreg.phase_correlation_registration(f, m)["quality"]


def cpu_registration_sample(views, n_pairs=None, n_jobs=None):
    """Oracle phase_correlation_registration on a sample of C2's pairs, one pair
    per worker process.  Returns (pairs/s, cores, sample description)."""
    from joblib import Parallel, delayed

    tiles = [v["data"] for v in views]
    pairs = c2_pairs()
    cores = n_jobs or os.cpu_count() or 1
    n_pairs = n_pairs or min(len(pairs), max(cores, 4))
    fixed, moving = pair_crops(tiles, pairs[:n_pairs])
    jobs = [(np.ascontiguousarray(f), np.ascontiguousarray(m)) for f, m in zip(fixed, moving)]
    t0 = time.perf_counter()
    Parallel(n_jobs=cores)(delayed(_cpu_register_pair)(j) for j in jobs)
    dt = time.perf_counter() - t0
    return n_pairs / dt, cores, f"{n_pairs} of {len(pairs)} overlap pairs (2048x307) of C2, oracle phase_correlation_registration, joblib x{cores}"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    views, params = _host_c2_views()
    vals, pvals = [], []
    # a step is seconds of CPU work on every host core: one warm-up pass, at most three timed
    cpu_fusion_sample(views, params)
    for _ in range(max(1, min(args.steps, 3))):
        v, cores, sample = cpu_fusion_sample(views, params)
        pv, _, psample = cpu_registration_sample(views)
        vals.append(v)
        pvals.append(pv)
    value = float(np.mean(vals))
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": len(vals),
        "warmup": min(args.warmup, 1),
        "ms_per_step": 81261210 / value / 1e3 if value else None,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": WORKLOAD,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "pairs_per_sec": float(np.mean(pvals)), "pairs_sample": psample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "registration": {"pairs_per_sec": float(np.mean(pvals)), "unit": "pairs/s"},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------


class Ctx:
    """Rank / world plumbing and max-over-ranks CUDA-event timing."""

    def __init__(self):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([float(x)], device="cuda", dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x):
        t = self.torch.tensor([float(x)], device="cuda", dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def timed(self, fn, steps, warmup=1):
        """ms per step of fn(): barrier + synchronize on both sides, CUDA events on the
        current stream, MAX over ranks."""
        torch = self.torch
        for _ in range(warmup):
            fn()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1) / steps)


def _median_ms(cx, fn, steps, warmup=1):
    """Median wall time per call of a host-driven stage (each call synchronised), MAX over ranks.
    The registration stages return to the host several times per call; on the shared boxes one call
    in five can take twice as long (interpreter / scheduler hiccups), which a mean of 2-5 calls
    turns into a number that changes by 50 % from run to run.  Returns (median, mean)."""
    torch = cx.torch
    for _ in range(warmup):
        fn()
    cx.barrier()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    cx.barrier()
    return cx.max_over_ranks(float(np.median(ts))), cx.max_over_ranks(float(np.mean(ts)))


def _roof(bytes_per_launch, ms, kernel=None):
    peak, src = _peaks()
    ach = bytes_per_launch / (ms * 1e-3) / 1e9
    r = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
         "algorithmic_bytes_per_launch": int(bytes_per_launch), "kernel_ms": ms}
    if kernel:
        r["kernel"] = kernel
    return r


def _fake_partial(msims, osp, chunksize, out, **fuse_kwargs):
    """The functools.partial `fuse(output_zarr_url=...)` hands a batch_func
    (fusion/_core.py:1133-1141, 2285-2293): keywords of _fuse_chunk_to_zarr."""

    def _never(block_id, **kw):
        raise AssertionError("the per-block CPU path must not run")

    fk = {"images": msims, "transform_key": "reg", "fusion_func": None, "weights_func": None,
          "interpolation_order": 1, "blending_widths": None, "backend": None, "output_chunksize": chunksize}
    fk.update(fuse_kwargs)
    return functools.partial(_never, output_stack_properties=osp, ns_shape={}, nsdims=[], fuse_kwargs=fk,
                             output_chunksize=chunksize, output_zarr_array=out)


def bench_c2(cx, args):
    torch = cx.torch
    from multiview_stitcher_b200 import fusion, geometry, pairs as pairs_mod, registration, synthetic
    from multiview_stitcher_b200.batch import BatchFuser, block_geometry

    world = cx.world
    seed = SEED + cx.rank  # every rank fuses its own replica (weak scaling)
    views, stage, true = synthetic.make_grid(GRID, TILE, OVERLAP, np.float32, jitter=2, seed=seed, subpixel=True)
    bbs = [v.bb() for v in views]
    osp = geometry.union_stack_props(bbs, true, bbs[0]["spacing"])
    plan = fusion.FusionPlan(views, true, osp)
    launches_per_step = plan.launches_per_run

    pairs = c2_pairs()
    tiles = [v.tensor for v in views]
    fixed, moving = pair_crops(tiles, pairs)
    fixed = [f.contiguous() for f in fixed]
    moving = [m.contiguous() for m in moving]
    pc_plans = {}

    def reg_step():
        return registration.register_pairs(fixed, moving, plans=pc_plans)

    for _ in range(max(args.warmup, 3)):
        plan.run()
    reg_res = reg_step()
    reg_step()
    # true pairwise shift = jitter difference (fractional): accuracy of the algorithm on this data
    true_t = np.array([t[:2, 2] for t in true])
    reg_err = max(float(np.abs(r["affine_matrix"][:2, 2] + (true_t[b] - true_t[a])).max())
                  for r, (a, b, _) in zip(reg_res, pairs))
    launches0 = sum(p.launch_count for p in pc_plans.values())
    n_reg = 5
    reg_ms, reg_ms_mean = _median_ms(cx, reg_step, n_reg, 0)
    reg_launches = (sum(p.launch_count for p in pc_plans.values()) - launches0) // n_reg

    # the same 40 pairs from the resident TILES (hook A's work): overlap boxes, crop windows,
    # resampling onto the fixed tile's grid, registration, physical transform
    t_plan = time.perf_counter()
    pair_plan = pairs_mod.PairPlan(views, stage, [(a, b) for a, b, _ in pairs], registration_binning={"y": 1, "x": 1})
    plan_ms = (time.perf_counter() - t_plan) * 1e3
    rv = pairs_mod.register_views(views, plan=pair_plan, pc_plans=pc_plans)
    rv_err = max(float(np.abs(r["transform"][:2, 2] + (true_t[b] - true_t[a])).max()) for r, (a, b, _) in zip(rv, pairs))
    rv_ms, rv_ms_mean = _median_ms(cx, lambda: pairs_mod.register_views(views, plan=pair_plan, pc_plans=pc_plans), n_reg, 1)

    # phase-correlation stage alone (FFT -> cross power -> IFFT -> peak -> upsampled DFT):
    # algorithmic bytes (32*ndim + 8) * N per pair over its CUDA-event time
    def pc_time(plans):
        b, ms = 0, 0.0
        for pcp in plans:
            for _ in range(2):
                pcp.correlate()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            for _ in range(3):
                pcp.correlate()
            c1.record()
            torch.cuda.synchronize()
            ms += c0.elapsed_time(c1) / 3
            b += (32 * pcp.ndim + 8) * pcp.voxels * pcp.n
        return b, ms

    pc_bytes, pc_ms = pc_time(pc_plans.values())
    p2_pairs = [(a, b) for a, b, ax in pairs if ax == 1]
    p2_fixed = [tiles[a][:, TILE[1] - 256:].contiguous() for a, b in p2_pairs]
    p2_moving = [tiles[b][:, :256].contiguous() for a, b in p2_pairs]
    p2_plan = registration.PhaseCorrPlan((TILE[0], 256), len(p2_pairs), 10)
    p2_plan.load_pairs(p2_fixed, p2_moving)
    p2_bytes, p2_ms = pc_time([p2_plan])
    p2_plan.close()
    del p2_fixed, p2_moving

    # ---- the headline: K fused steps, each step also timed on its own ----
    cx.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ev0.record()
    for i in range(args.steps):
        kev[i][0].record()
        plan.run()
        kev[i][1].record()
    ev1.record()
    cx.barrier()
    total_ms = cx.max_over_ranks(ev0.elapsed_time(ev1))
    fuse_kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    vox_per_step = plan.out_voxels
    value = vox_per_step * world * args.steps / (total_ms * 1e-3) / 1e6

    # ---- end to end through hook C with pageable numpy buffers ----
    host_tiles = [v.tensor.cpu().numpy() for v in views]  # plain (pageable) numpy, as a reader hands them over
    msims = [{"data": d, "origin": v.origin, "spacing": v.spacing, "transforms": {"reg": p}}
             for d, v, p in zip(host_tiles, views, true)]
    out_shape = tuple(int(osp["shape"][d]) for d in "yx")
    out_host = np.zeros(out_shape, dtype=np.float32)  # the destination "zarr" array: plain numpy
    chunksize = {"y": 2048, "x": 2048}
    fuse_chunk = _fake_partial(msims, osp, chunksize, out_host)
    block_ids = sorted(block_geometry(osp, chunksize))
    bf = BatchFuser()

    def e2e_step():
        bf.reset()  # tiles cross PCIe again every step
        bf(fuse_chunk, block_ids)

    e2e_step()
    dev_ref = plan.out.cpu().numpy()
    e2e_ok = bool(np.array_equal(out_host, dev_ref))
    h0, d0 = bf.h2d_bytes, bf.d2h_bytes
    e2e_step()
    h2d, d2h = bf.h2d_bytes - h0, bf.d2h_bytes - d0
    cx.barrier()
    n_e2e = max(2, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = cx.max_over_ranks((time.perf_counter() - t0) / n_e2e)
    cx.barrier()
    e2e_value = vox_per_step * world / e2e_s / 1e6
    e2e_launches = bf.launches

    # secondary: the engine's own host API with PRE-PINNED buffers (round-1 figure)
    pinned = [{"data": torch.from_numpy(d).pin_memory(), "origin": v.origin, "spacing": v.spacing}
              for d, v in zip(host_tiles, views)]
    out_pin = torch.empty(out_shape, dtype=torch.float32).pin_memory()
    fuser = fusion.HostFuser(pinned, true, output_stack_properties=osp)

    def pinned_step():
        fuser(pinned, out_pin)
        torch.cuda.synchronize()

    pinned_step()
    cx.barrier()
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        pinned_step()
    pin_s = cx.max_over_ranks((time.perf_counter() - t0) / n_e2e)
    fuser.close()
    bf.close()
    del pinned, out_pin

    # registration end to end: host crops in (pageable numpy) -> transforms out
    hf, hm = pair_crops(host_tiles, pairs)
    hf = [np.ascontiguousarray(a) for a in hf]
    hm = [np.ascontiguousarray(a) for a in hm]
    reg_e2e_ms, reg_e2e_mean = _median_ms(cx, lambda: registration.register_pairs(hf, hm, plans=pc_plans), 5, 1)
    reg_e2e_s = reg_e2e_ms * 1e-3

    # ---- output side: fused stack -> OME-Zarr 0.4 (pyramid levels binned on the device,
    # chunks encoded on the device, raw chunk files written through pinned staging) and read back ----
    zrec = None
    if cx.rank == 0:
        import shutil
        import tempfile

        from multiview_stitcher_b200 import ngff_io

        zdir = tempfile.mkdtemp(prefix="mvs_b200_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        try:
            dv = fusion.DeviceView(plan.out, osp["origin"], osp["spacing"])
            url = os.path.join(zdir, "fused.zarr")
            ngff_io.write_sim_to_ome_zarr(dv, url, overwrite=True, chunks=chunksize)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            wres = ngff_io.write_sim_to_ome_zarr(dv, url, overwrite=True, chunks=chunksize)
            torch.cuda.synchronize()
            w_s = time.perf_counter() - t0
            t0 = time.perf_counter()
            back = ngff_io.read_sim_from_ome_zarr(url, 0)
            torch.cuda.synchronize()
            r_s = time.perf_counter() - t0
            zrec = {"what": "ngff_io.write_sim_to_ome_zarr(fused C2 stack): mvs_bin_mean per level + mvs_chunks_pack + "
                            "mvs_chunks_store (raw Zarr v2 chunks of 2048x2048, store in " + os.path.dirname(zdir) + "), "
                            "then read_sim_from_ome_zarr level 0 (mvs_chunks_load + mvs_chunks_unpack)",
                    "levels": len(wres["shapes"]), "bytes_written": int(wres["bytes_written"]),
                    "write_ms": w_s * 1e3, "write_Mvoxel_per_s": vox_per_step / w_s / 1e6,
                    "read_level0_ms": r_s * 1e3, "read_back_equal": bool(torch.equal(back.tensor, plan.out))}
            del back
        except Exception as e:  # a full or missing scratch directory must not take the headline down
            zrec = {"error": f"{type(e).__name__}: {e}"}
        finally:
            shutil.rmtree(zdir, ignore_errors=True)

    traffic, traffic_src = _ncu_traffic()
    roof = _roof(plan.algorithmic_bytes(), fuse_kernel_ms, "fuse_stencil_kernel<2,float,WAVG> (TMA-staged translation path)")
    roof.update({"peak_source": _peaks()[1], "frac_of_nominal_8TBps": roof["achieved"] / 8000.0,
                 "traffic": traffic, "traffic_source": traffic_src})
    rec = {
        "value": value,
        "ms_per_step": total_ms / args.steps,
        "vox_per_step": vox_per_step,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "path": "hook C: BatchFuser.__call__(fuse_chunk partial, all 25 block ids) -- pageable numpy tiles in, "
                        "numpy destination array out; pinned staging + H2D + fusion + D2H inside the timing",
                "ms_per_step": e2e_s * 1e3, "equals_device_result": e2e_ok, "gpu_launches_per_step": e2e_launches // (n_e2e + 2),
                "pre_pinned_HostFuser": {"value": vox_per_step * world / pin_s / 1e6, "ms_per_step": pin_s * 1e3,
                                         "what": "fusion.HostFuser with buffers pinned outside the timing (round-1 e2e)"}},
        "registration": {
            "pairs_per_sec": len(pairs) * world / (reg_ms * 1e-3),
            "unit": "pairs/s",
            "pairs_per_step": len(pairs),
            "ms_per_step": reg_ms,
            "ms_per_step_mean": reg_ms_mean,
            "timing": "median of >= 5 synchronised calls (mean beside it); the fusion figures are means over K steps",
            "crop": "2048x307 / 307x2048 float32",
            "max_abs_shift_error_px": reg_err,
            "shift_error_note": "vs the generator's fractional jitter difference; the algorithm's sub-pixel grid is 0.1 px (upsample_factor 10)",
            "gpu_launches_per_step": reg_launches,
            "e2e": {"pairs_per_sec": len(pairs) * world / reg_e2e_s, "ms_per_step": reg_e2e_s * 1e3, "ms_per_step_mean": reg_e2e_mean,
                    "path": "registration.register_pairs: pageable numpy crops in (2 x 40 x 2.5 MB H2D), affine + quality out"},
            "from_tiles": {
                "what": "pairs.register_views (hook A's work): crop to the overlap box + resample onto the fixed tile's grid "
                        "(1 launch per crop shape) + registration + physical transform, tiles resident",
                "pairs_per_sec": len(pairs) * world / (rv_ms * 1e-3),
                "ms_per_step": rv_ms,
                "ms_per_step_mean": rv_ms_mean,
                "max_abs_shift_error_px": rv_err,
                "host_geometry_plan_ms_once": plan_ms,
            },
            "phasecorr_roofline": dict(_roof(pc_bytes, pc_ms), stage="mvs_pc_correlate on C2's 40 crops (307 is prime: Bluestein axes)"),
            "phasecorr_roofline_pow2": dict(_roof(p2_bytes, p2_ms), stage="mvs_pc_correlate on 20 crops of 2048x256 (power-of-two axes)"),
        },
        "gpu_launches": launches_per_step * args.steps + reg_launches * n_reg,
        "roofline": roof,
        "output_zarr": zrec,
    }
    plan.close()
    for p in pc_plans.values():
        p.close()
    return rec


def _content_roofline(views, params, osp, bbs, b_fuse, ms, chunk_subset, halo):
    """algorithmic bytes of content-weighted fusion = B_fuse + 108 B (3-D) per contributing
    view-voxel INCLUDING the halo (resampled view stored once, two NaN-normalised separable Gaussians
    of 3 passes over a (value, mask) pair each, final read); views counted per chunk like
    fuse_with_weights selects them (transformed bounding box touches the chunk + halo)."""
    from multiview_stitcher_b200 import geometry

    dims = list("zyx")
    o_org, o_sp, _ = geometry.bb_arrays(osp, dims)
    aabbs = [geometry.transformed_aabb(bb, p, dims) for bb, p in zip(bbs, params)]
    grid = geometry.chunk_grid(osp, geometry.DEFAULT_CHUNKSIZE_3D)
    view_vox = halo_vox = 0
    share = 0.0
    for ci in chunk_subset:
        start, shape = grid[ci]
        lo = (o_org + o_sp * np.array(start)) - halo * o_sp
        hi = lo + (np.array(shape) + 2 * halo - 1) * o_sp
        nv = sum(1 for alo, ahi in aabbs if not (np.any(ahi < lo - 1e-6) or np.any(alo > hi + 1e-6)))
        hv = int(np.prod(np.array(shape) + 2 * halo))
        view_vox += nv * hv
        halo_vox += hv
        share += float(np.prod(shape))
    out_vox = float(np.prod([osp["shape"][d] for d in dims]))
    nbytes = b_fuse * share / out_vox + 108.0 * view_vox  # this rank's share of B_fuse + its chunks' passes
    roof = _roof(nbytes, ms, "content-weighted chunk pipeline (resample + 2 nan-Gaussians + blend)")
    roof["note"] = ("bound by the FP64 pipe, not HBM: the Gaussian taps are evaluated in scipy's exact float64 "
                    "operation order (DADD, DMUL, DADD per tap pair)")
    return {"roofline": roof, "view_voxels_incl_halo": int(view_vox), "halo_voxels_this_rank": int(halo_vox),
            "algorithmic_bytes_note": "B_fuse share + 108 B per contributing view-voxel incl. the 22 px halo (SURVEY 8d)"}


def _c3_pairs(grid):
    idx = list(np.ndindex(*grid))
    pos = {c: i for i, c in enumerate(idx)}
    pairs = []
    for c in idx:
        for ax in range(3):
            n = list(c)
            n[ax] += 1
            if tuple(n) in pos:
                pairs.append((pos[c], pos[tuple(n)]))
    return pairs


def bench_c3(cx, args):
    """C3 as ONE job: at N ranks the 64 pairs are dealt round-robin, the 128 output chunks
    go to the ranks in bands (tiles replicated), and the tile-partitioned run keeps 32/N
    tiles per rank and exchanges border boxes over NVLink.  Strong scaling."""
    torch = cx.torch
    from multiview_stitcher_b200 import content, distributed, fusion, geometry, synthetic

    grid, tile, ov = C3["grid"], C3["tile"], C3["overlap"]
    views, stage, true = synthetic.make_grid(grid, tile, ov, np.uint16, jitter=2, seed=SEED, subpixel=True)
    bbs = [v.bb() for v in views]
    osp = geometry.union_stack_props(bbs, true, bbs[0]["spacing"])
    vox = int(np.prod([osp["shape"][d] for d in "zyx"]))
    b_fuse = sum(v.tensor.numel() * 2 for v in views) + vox * 2
    n_steps = max(3, min(args.steps, 10))
    rec = {"workload": "C3: 4x4x2 grid of 3-D tiles (z256,y512,x512) uint16, 10% overlap, sub-pixel positions; "
                       f"output {[osp['shape'][d] for d in 'zyx']}", "scaling": "strong" if cx.world > 1 else None}

    # (1) blend fusion, whole stack on this GPU (N = 1 figure; at N > 1 every rank runs it too)
    plan = fusion.FusionPlan(views, true, osp)
    ms = cx.timed(plan.run, n_steps, 2)
    rec["fuse_blend_one_gpu"] = {"ms": ms, "Mvoxel_per_s": vox / ms / 1e3, "roofline": _roof(b_fuse, ms, "fuse_stencil 3-D uint16"),
                                 "launches": plan.launches_per_run}
    ref_out = plan.out
    launches = plan.launches_per_run * n_steps

    if cx.world > 1:
        # (2) chunk bands per rank, tiles replicated, no communication
        cs = geometry.DEFAULT_CHUNKSIZE_3D
        sf = distributed.ShardedFuser(views, true, osp, cs)
        ms_sh = cx.timed(sf.run, n_steps, 1)
        rec["fuse_sharded"] = {"ms": ms_sh, "Mvoxel_per_s": vox / ms_sh / 1e3, "chunks_this_rank": len(sf.owned),
                               "what": "distributed.ShardedFuser: chunk bands per rank (tiles replicated), planned once, no collective"}
        sf.close()
        del sf
        # (3) tiles partitioned: each tile lives on one rank; border boxes cross NVLink
        idx = list(np.ndindex(*grid))
        cols = grid[1] * grid[2]
        owners = [((c[1] * grid[2] + c[2]) * cx.world) // cols for c in idx]
        local = {i: views[i] for i in range(len(views)) if owners[i] == cx.rank}
        for mode in ("halo", "partial"):
            tp = distributed.TilePartitionedFuser(local, bbs, true, owners, osp, cs, mode=mode)
            ms_tp = cx.timed(tp.run, 3, 1)
            # parity vs the one-GPU result on the chunks this rank owns (uint16: <= 1 LSB; halo: equal)
            part = tp.partition
            sl = tuple(slice(a, a + n) for a, n in zip(tp.out_start, tp.out.shape))
            diff = (tp.out.to(torch.int32) - ref_out[sl].to(torch.int32)).abs()
            mask = torch.zeros_like(diff, dtype=torch.bool)
            for ci, (cs_, cn_) in enumerate(part.grid):
                if part.owner_of[ci] == cx.rank:
                    mask[tuple(slice(a - o, a - o + n) for a, o, n in zip(cs_, tp.out_start, cn_))] = True
            max_lsb = cx.max_over_ranks(float((diff * mask).max().item()))
            xbytes = part.halo_bytes(2) if mode == "halo" else part.exchanged_bytes()
            rec["fuse_tile_partitioned_" + mode] = {
                "ms": ms_tp, "Mvoxel_per_s": vox / ms_tp / 1e3, "tiles_per_rank": len(local),
                "nvlink_bytes_per_job": int(xbytes), "sent_bytes_all_ranks": int(cx.sum_over_ranks(tp.sent_bytes)),
                "border_boxes": len(part.entries),
                "full_stack_reduce_bytes": int(8 * vox), "exchange_fraction_of_full_reduce": xbytes / (8.0 * vox),
                "max_abs_diff_vs_one_gpu_lsb": max_lsb, "launches_per_step": tp.launches,
                "what": ("raw uint16 windows of the foreign tiles -> owner (NCCL send/recv, hidden behind the direct launch), "
                         "border boxes fused by the ordinary kernel" if mode == "halo" else
                         "partial (sum w*v, sum w) float32 of the border boxes -> owner (NCCL send/recv), add + divide + cast"),
            }
            tp.close()
            del tp, diff, mask
            torch.cuda.empty_cache()
    plan.close()
    del ref_out

    # (4) registration of all 64 face pairs from the resident tiles (pairs round-robin over ranks)
    from multiview_stitcher_b200 import pairs as pairs_mod

    pairs = _c3_pairs(grid)
    binning = {"z": 1, "y": 1, "x": 1}
    owned = distributed.shard_round_robin(len(pairs), cx.rank, cx.world)
    my_pairs = [pairs[i] for i in owned]
    pc_plans = {}
    pplan = pairs_mod.PairPlan(views, stage, my_pairs, registration_binning=binning)
    res = pairs_mod.register_views(views, plan=pplan, pc_plans=pc_plans)
    tt = np.array([t[:3, 3] for t in true])
    err = max(float(np.abs(r["transform"][:3, 3] + (tt[b] - tt[a])).max()) for r, (a, b) in zip(res, my_pairs))
    # the first calls grow the plans' scratch buffers (several GB per crop shape: cudaFree / cudaMalloc
    # synchronise the device); measured call times 519 911 720 254 275 211 213 246 ms -> 3 warm-up calls
    ms_reg, ms_reg_mean = _median_ms(cx, lambda: pairs_mod.register_views(views, plan=pplan, pc_plans=pc_plans), 5, 3)
    rec["registration_from_tiles"] = {"pairs": len(pairs), "pairs_this_rank": len(my_pairs), "ms": ms_reg, "ms_mean": ms_reg_mean,
                                      "pairs_per_sec": len(pairs) / (ms_reg * 1e-3),
                                      "max_abs_shift_error_px": cx.max_over_ranks(err),
                                      "shift_error_note": "3-D default upsample_factor 2: the algorithm's sub-pixel grid is 0.5 px"}
    for p in pc_plans.values():
        p.close()

    # (5) content-weighted fusion (the mode C3 names): chunks of 256^3 + 22 px halo, sigma 5 / 11
    n_chunks = len(geometry.chunk_grid(osp, geometry.DEFAULT_CHUNKSIZE_3D))
    mine = distributed.shard_slabs(n_chunks, cx.rank, cx.world)

    def content_run():
        return content.fuse_with_weights(views, true, osp, None, fusion.weighted_average_fusion, fusion.content_based,
                                         None, 1, None, chunk_subset=mine)

    ms_c = cx.timed(content_run, 1, 1)
    rec["fuse_content_weighted"] = {"ms": ms_c, "Mvoxel_per_s": vox / ms_c / 1e3, "chunks_this_rank": len(mine)}
    rec["fuse_content_weighted"].update(_content_roofline(views, true, osp, bbs, b_fuse, ms_c, mine, 22))
    rec["gpu_launches"] = launches
    del views
    torch.cuda.empty_cache()
    return rec


def bench_c4(cx, args):
    """C4: 4 views of (512,1024,1024) uint16, rotations 0/90/180/270 deg about y with a +-2 deg tilt
    and 0.5 % anisotropic scale, spacing z=2: general-affine path.  One GPU (rank 0's figure)."""
    torch = cx.torch
    from multiview_stitcher_b200 import content, fusion, geometry, synthetic
    from multiview_stitcher_b200.fusion import DeviceView

    shape = (512, 1024, 1024)
    spacing = {"z": 2.0, "y": 1.0, "x": 1.0}
    ext = np.array([shape[0] * 2.0, shape[1] * 1.0, shape[2] * 1.0])
    centre = ext / 2
    views, params = [], []
    for k in range(4):
        t = synthetic.make_tile_field(shape, (0.0, 0.0, 0.0), np.uint16, seed=SEED + 10 + k, tile_id=k)
        views.append(DeviceView(t, {"z": 0.0, "y": 0.0, "x": 0.0}, spacing))
        a = np.deg2rad(90.0 * k)
        tilt = np.deg2rad(2.0 if k % 2 else -2.0)
        ry = np.array([[np.cos(a), 0, -np.sin(a)], [0, 1, 0], [np.sin(a), 0, np.cos(a)]])
        rx = np.array([[np.cos(tilt), np.sin(tilt), 0], [-np.sin(tilt), np.cos(tilt), 0], [0, 0, 1]])
        m = ry @ rx @ np.diag([1.0, 1.005, 0.995])
        p = np.eye(4)
        p[:3, :3] = m
        p[:3, 3] = centre - m @ centre
        params.append(p)
    bbs = [v.bb() for v in views]
    osp = geometry.union_stack_props(bbs, params, spacing)
    vox = int(np.prod([osp["shape"][d] for d in "zyx"]))
    plan = fusion.FusionPlan(views, params, osp)
    ms = cx.timed(plan.run, 3, 1)
    b = plan.algorithmic_bytes()
    rec = {"workload": "C4: 4 views (z512,y1024,x1024) uint16, spacing z=2, rotations about y + tilt + 0.5% scale; "
                       f"output {[osp['shape'][d] for d in 'zyx']}",
           "fuse_blend": {"ms": ms, "Mvoxel_per_s": vox / ms / 1e3, "roofline": _roof(b, ms, "fuse_affine_kernel<3,1,WAVG> (general affine, TMA-staged bricks)")},
           "gpu_launches": plan.launches_per_run * 3}
    plan.close()

    def content_run():
        return content.fuse_with_weights(views, params, osp, None, fusion.weighted_average_fusion, fusion.content_based, None, 1, None)

    ms_c = cx.timed(content_run, 1, 1)
    rec["fuse_content_weighted"] = {"ms": ms_c, "Mvoxel_per_s": vox / ms_c / 1e3}
    n_chunks = len(geometry.chunk_grid(osp, geometry.DEFAULT_CHUNKSIZE_3D))
    rec["fuse_content_weighted"].update(_content_roofline(views, params, osp, bbs, b, ms_c, range(n_chunks), 22))
    del views
    torch.cuda.empty_cache()
    return rec


def bench_c5(cx, args):
    """C5 weak-scaled: rank r holds tile row r of the 8x8 grid (8 tiles of 512x2048x2048 uint16
    = 34 GB); N ranks fuse the first N rows as one tile-partitioned job (N = 8: the whole 137 Gvoxel
    stack).  Only the row-overlap boxes cross NVLink."""
    torch = cx.torch
    from multiview_stitcher_b200 import distributed, geometry, synthetic

    grid = (1, cx.world, C5["grid"][2])
    tile, ov = C5["tile"], C5["overlap"]
    owners = [c[1] for c in np.ndindex(*grid)]
    mine = {i for i, o in enumerate(owners) if o == cx.rank}
    t0 = time.perf_counter()
    views, stage, true = synthetic.make_grid(grid, tile, ov, np.uint16, jitter=2, seed=SEED, subpixel=True, only=mine)
    torch.cuda.synchronize()
    gen_s = time.perf_counter() - t0
    dims = ["z", "y", "x"]
    sp = {d: 1.0 for d in dims}
    bbs = []
    for k, c in enumerate(np.ndindex(*grid)):
        s_org = np.array(c) * (np.array(tile) - np.array(ov))
        bbs.append({"origin": dict(zip(dims, map(float, s_org))), "spacing": dict(sp), "shape": dict(zip(dims, tile))})
    osp = geometry.union_stack_props(bbs, true, sp)
    vox = int(np.prod([osp["shape"][d] for d in dims]))
    local = {i: views[i] for i in mine}
    n_steps = max(2, min(args.steps, 5))
    rec = {
        "workload": f"C5: {cx.world} row(s) of the 8x8 grid of 3-D tiles (z512,y2048,x2048) uint16, 10% overlap, one row of 8 tiles "
                    f"(34 GB) per GPU; output {[osp['shape'][d] for d in dims]}",
        "scaling": "weak", "out_voxels": vox, "tiles_per_rank": len(mine),
        "input_bytes_all_ranks": int(len(owners) * np.prod(tile) * 2), "generate_tiles_s": gen_s, "gpu_launches": 0,
    }
    for mode in ("halo", "partial") if cx.world > 1 else ("halo",):
        tp = distributed.TilePartitionedFuser(local, bbs, true, owners, osp, geometry.DEFAULT_CHUNKSIZE_3D, mode=mode)
        ms = cx.timed(tp.run, n_steps, 1)
        part = tp.partition
        xbytes = part.halo_bytes(2) if mode == "halo" else part.exchanged_bytes()
        sub = {"ms": ms, "Mvoxel_per_s": vox / ms / 1e3, "nvlink_bytes_per_job": int(xbytes),
               "sent_bytes_all_ranks": int(cx.sum_over_ranks(tp.sent_bytes)), "border_boxes": len(part.entries),
               "launches_per_step": tp.launches}
        rec["gpu_launches"] += tp.launches * (n_steps + 1)
        if mode == "halo":
            # kernel-only figure of this rank's direct boxes (the fused stencil launch)
            ms_direct = cx.timed(tp.direct.run, n_steps, 0)
            own_vox = sum(int(np.prod(n)) for ci, (s_, n) in enumerate(part.grid) if part.owner_of[ci] == cx.rank)
            b_local = len(mine) * int(np.prod(tile)) * 2 + own_vox * 2
            rec.update({"ms": ms, "Mvoxel_per_s": vox / ms / 1e3})
            rec["direct_launch"] = {
                "ms": ms_direct, "roofline": _roof(b_local, ms_direct, "fuse_stencil 3-D uint16, this rank's direct boxes"),
                "note": "algorithmic bytes = this rank's tiles + the chunks it owns (the border boxes are in the bytes but "
                        "not in this launch, so the fraction is a slight over-estimate at N > 1)"}
        rec["exchange_" + mode] = sub
        tp.close()
        del tp
        torch.cuda.empty_cache()
    del views, local
    torch.cuda.empty_cache()
    return rec


def run_ours(args):
    cx = Ctx()
    torch = cx.torch
    only = set((args.only or "c2,c3,c4,c5").split(","))
    clocks = ClockSampler(cx.local_rank)
    clocks.__enter__()
    c2 = bench_c2(cx, args)
    torch.cuda.empty_cache()
    configs = {}
    errors = {}
    for name, fn in (("C3", bench_c3), ("C4", bench_c4), ("C5", bench_c5)):
        if name.lower() not in only:
            continue
        if name == "C4" and cx.world > 1:
            continue  # single-GPU config; measured at N = 1
        try:
            configs[name] = fn(cx, args)
        except Exception as e:  # a sub-record must not take the headline down with it
            import traceback

            errors[name] = f"{type(e).__name__}: {e}"
            if cx.rank == 0:
                traceback.print_exc()
            if cx.world > 1:
                raise
        torch.cuda.empty_cache()
    clocks.__exit__(None, None, None)

    if cx.rank == 0:
        line = {
            "metric": METRIC,
            "value": c2["value"],
            "unit": UNIT,
            "n_gpus": cx.world,
            "steps": args.steps,
            "warmup": max(args.warmup, 3),
            "ms_per_step": c2["ms_per_step"],
            "higher_is_better": True,
            "scaling": "weak",
            "scaling_note": "value = C2, one replica per rank (independent units, no data-path collective); the sharded jobs "
                            "(C3 strong, C5 weak with border exchange over NVLink) are under configs",
            "vs_baseline": None,
            "dtype": "f32",
            "data": "synthetic",
            "config": WORKLOAD,
            "e2e": c2["e2e"],
            "registration": c2["registration"],
            "gpu_launches": c2["gpu_launches"] + sum(c.get("gpu_launches", 0) for c in configs.values()),
            "clocks": clocks.summary(),
            "roofline": c2["roofline"],
            "output_zarr": c2["output_zarr"],
            "configs": configs,
        }
        if errors:
            line["config_errors"] = errors
        if cx.world == 1 and not args.no_cpu:
            views, params = _host_c2_views()
            v, cores, sample = cpu_fusion_sample(views, params)
            pv, _, psample = cpu_registration_sample(views)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                                    "pairs_per_sec": pv, "pairs_sample": psample}
        print(json.dumps(line))
    if cx.world > 1:
        cx.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--only", default=None, help="comma list of configs to run beside C2 (c3,c4,c5)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
