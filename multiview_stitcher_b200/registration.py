"""Host side of the batched phase-correlation registration path.

``phase_correlation_registration`` has the signature of the reference's
default ``pairwise_reg_func`` (registration.py:353-358) and returns the same
``{"affine_matrix", "quality"}`` dict, so ``registration.register(...,
pairwise_reg_func=phase_correlation_registration)`` calls it unchanged.
``register_pairs`` is the batched form (all overlap pairs of a tile grid in a
handful of launches) that a ``pairwise_executor`` (registration.py:2634-2655)
drives.

What runs where: FFTs, cross-power spectra, peak search, upsampled DFT,
candidate resampling, masks, SSIM and Spearman run on the GPU
(csrc/phasecorr.cu, csrc/disambig.cu).  The data-dependent glue between the
stages -- candidate expansion (:461-477), the 10 % overlap rule (:503), bbox
slices (:507-528), window size (:535-536), the append-or-skip list semantics
(:530-533) and the final argmax (:558-563) -- is O(candidates) host logic and
mirrors the reference line by line.  No CPU fallback exists for the array work.
"""

from __future__ import annotations

import ctypes
import itertools
import os
import warnings

import numpy as np

from . import _lib
from ._lib import EngineError

__all__ = [
    "phase_correlation_registration",
    "register_pairs",
    "dispatch_pairwise_reg_func",
    "PhaseCorrPlan",
]


def affine_from_translation(translation):
    """Homogeneous matrix of a translation (param_utils.py:7-14)."""
    t = np.asarray(translation, dtype=np.float64)
    m = np.eye(len(t) + 1)
    m[:-1, -1] = t
    return m


class PhaseCorrPlan:
    """Device buffers + FFT tables for pairs of one crop shape."""

    def __init__(self, shape, max_pairs, upsample_factor):
        lib = _lib.load(require_device=True)
        self._lib = lib
        self.ndim = len(shape)
        if self.ndim not in (2, 3):
            raise EngineError(f"phase correlation needs 2-D or 3-D crops, got {self.ndim}-D")
        self.shape = tuple(int(s) for s in shape)
        self.shape3 = (1,) * (3 - self.ndim) + self.shape
        self.max_pairs = int(max_pairs)
        self.upsample = int(upsample_factor)
        if self.upsample != upsample_factor:
            raise EngineError("upsample_factor must be an integer")
        handle = ctypes.c_void_p()
        shp = (ctypes.c_int32 * 3)(*self.shape3)
        _lib.check(
            lib.mvs_pc_plan_create(ctypes.byref(handle), self.ndim, shp, self.max_pairs, self.upsample),
            "mvs_pc_plan_create",
        )
        self._h = handle
        region, vox, launches = ctypes.c_int(), ctypes.c_int64(), ctypes.c_int()
        lib.mvs_pc_plan_info(handle, ctypes.byref(region), ctypes.byref(vox), ctypes.byref(launches))
        self.region = region.value
        self.voxels = vox.value
        self.launches_per_correlate = launches.value
        self.launch_count = 0
        self._keep = None

    # ---- stages --------------------------------------------------------------

    def load_pairs(self, fixed, moving):
        """fixed / moving: lists of CUDA float32 contiguous tensors."""
        n = len(fixed)
        if n < 1 or n > self.max_pairs or len(moving) != n:
            raise EngineError(f"{n} pairs for a plan of {self.max_pairs}")
        for t in list(fixed) + list(moving):
            if tuple(t.shape) != self.shape:
                raise EngineError(f"crop shape {tuple(t.shape)} != plan shape {self.shape}")
        self._keep = (fixed, moving)  # keep inputs alive while the plan refers to them
        fa = (ctypes.c_void_p * n)(*[t.data_ptr() for t in fixed])
        ma = (ctypes.c_void_p * n)(*[t.data_ptr() for t in moving])
        stats = np.zeros((2 * n, 9), dtype=np.float64)
        _lib.check(
            self._lib.mvs_pc_load_pairs(self._h, n, fa, ma, stats.ctypes.data_as(ctypes.c_void_p), _lib.current_stream_ptr()),
            "mvs_pc_load_pairs",
        )
        self.n = n
        self.launch_count += 2
        return stats.reshape(n, 2, 9)

    def correlate(self):
        n = self.n
        peaks = np.zeros((n, 2, 3), dtype=np.int32)
        rn = self.region**self.ndim if self.upsample > 1 else 1
        updft = np.zeros((n, 2, rn, 2), dtype=np.float64)
        _lib.check(
            self._lib.mvs_pc_correlate(
                self._h, n, peaks.ctypes.data_as(ctypes.c_void_p), updft.ctypes.data_as(ctypes.c_void_p), _lib.current_stream_ptr()
            ),
            "mvs_pc_correlate",
        )
        self.launch_count += self.launches_per_correlate
        return peaks, updft[..., 0] + 1j * updft[..., 1]

    def _cand_arrays(self, cand_pair, cand_t):
        cp = np.ascontiguousarray(cand_pair, dtype=np.int32)
        ct = np.zeros((len(cp), 3), dtype=np.float64)
        ct[:, 3 - self.ndim :] = np.asarray(cand_t, dtype=np.float64).reshape(len(cp), self.ndim)
        return cp, ct

    def candidate_stats(self, cand_pair, cand_t):
        cp, ct = self._cand_arrays(cand_pair, cand_t)
        out = np.zeros((len(cp), 8), dtype=np.int64)
        _lib.check(
            self._lib.mvs_pc_candidate_stats(
                self._h, len(cp), cp.ctypes.data_as(ctypes.c_void_p), ct.ctypes.data_as(ctypes.c_void_p),
                out.ctypes.data_as(ctypes.c_void_p), _lib.current_stream_ptr(),
            ),
            "mvs_pc_candidate_stats",
        )
        self.launch_count += 1
        return out

    def candidate_ssim(self, cand_pair, cand_t, slices, win):
        cp, ct = self._cand_arrays(cand_pair, cand_t)
        sl = np.zeros((len(cp), 6), dtype=np.int32)
        sl[:, :3] = 0
        sl[:, 3:6] = 1
        s = np.asarray(slices, dtype=np.int32).reshape(len(cp), 2, self.ndim)
        sl[:, 3 - self.ndim : 3] = s[:, 0]
        sl[:, 6 - self.ndim : 6] = s[:, 1]
        w = np.ascontiguousarray(win, dtype=np.int32)
        out = np.zeros((len(cp), 2), dtype=np.float64)
        _lib.check(
            self._lib.mvs_pc_candidate_ssim(
                self._h, len(cp), cp.ctypes.data_as(ctypes.c_void_p), ct.ctypes.data_as(ctypes.c_void_p),
                sl.ctypes.data_as(ctypes.c_void_p), w.ctypes.data_as(ctypes.c_void_p),
                out.ctypes.data_as(ctypes.c_void_p), _lib.current_stream_ptr(),
            ),
            "mvs_pc_candidate_ssim",
        )
        self.launch_count += 1
        return out

    def spearman_batch(self, pairs, ts, n_mask):
        """Spearman of im0[mask] vs im1t[mask] for each (pair, t); one sync."""
        n = len(pairs)
        cp, ct = self._cand_arrays(pairs, ts)
        nm = np.ascontiguousarray(n_mask, dtype=np.int64)
        rho = np.zeros(n, dtype=np.float64)
        _lib.check(
            self._lib.mvs_pc_spearman_batch(
                self._h, n, cp.ctypes.data_as(ctypes.c_void_p), ct.ctypes.data_as(ctypes.c_void_p),
                nm.ctypes.data_as(ctypes.c_void_p), rho.ctypes.data_as(ctypes.c_void_p), _lib.current_stream_ptr(),
            ),
            "mvs_pc_spearman_batch",
        )
        # per sub-batch of <= 4 pairs: keys, 2 x (histogram + exclusive sum + 4 onesweep passes),
        # 2 x (run flags + 2 x (scan init + scan)), rank, pearson  (ncu launch list: profiles/r02_reg_launches_summary.txt)
        self.launch_count += 29 * ((n + 3) // 4)
        return rho

    def spearman(self, pair, t, n_mask):
        return float(self.spearman_batch([pair], [t], [n_mask])[0])

    def close(self):
        if getattr(self, "_h", None):
            self._lib.mvs_pc_plan_destroy(self._h)
            self._h = None
            self._keep = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# --- host glue (mirrors registration.py:410-563) ------------------------------


def _subpixel_shifts(plan, peaks, updft):
    """skimage.phase_cross_correlation's float32 shift arithmetic on the
    engine's integer peaks and upsampled-DFT samples (vectorised over pairs).
    Returns shifts[pair][norm] (norm 0 = None, 1 = "phase")."""
    n, ndim, u = peaks.shape[0], plan.ndim, plan.upsample
    uf = np.float32(u)
    region = plan.region
    shift = peaks[:, :, 3 - ndim :].astype(np.float32)
    if u > 1:
        dftshift = np.fix(np.float32(region) / np.float32(2.0))
        shift = np.round(shift * uf) / uf
        cc = np.abs(updft.reshape(n, 2, -1).astype(np.complex64))
        flat = np.argmax(cc, axis=-1)
        maxima = np.stack(np.unravel_index(flat, (region,) * ndim), axis=-1).astype(np.float32) - dftshift
        shift = shift + maxima / uf
    for d in range(ndim):
        if plan.shape[d] == 1:
            shift[:, :, d] = 0
    return shift.astype(np.float32)


def _expand_candidates(shift_cands, shape, max_shift_per_dim):
    """registration.py:461-477: per shift and dim {s, -s, -(s - N), -s - N} (one
    option when s == 0), in np.ndindex order; float32 arithmetic like the
    reference's (skimage returns float32 shifts)."""
    ndim = len(shape)
    t_candidates = []
    for sc in shift_cands:
        opts = []
        for d in range(ndim):
            s_ = np.float32(sc[d])
            if s_ == 0:
                opts.append((float(s_),))
            else:
                opts.append((float(s_), float(-s_), float(-(s_ - shape[d])), float(-s_ - shape[d])))
        for t in itertools.product(*opts):
            if max(abs(x) for x in t) < max_shift_per_dim:
                t_candidates.append(list(t))
    return t_candidates


def _valid_range(n, t):
    """Integers x in [0, n) with 0 <= x + t <= n - 1 in float64 -- the outside
    predicate of scipy's affine_transform (and of the engine's shifted_value).
    ``t``: float64 array; returns (lo, hi) int64 arrays (hi < lo: empty)."""
    t = np.asarray(t, dtype=np.float64)
    lo = np.maximum(0, np.ceil(-t)).astype(np.int64)
    lo = np.where((lo > 0) & ((lo - 1).astype(np.float64) + t >= 0.0), lo - 1, lo)
    lo = np.where((lo < n) & (lo.astype(np.float64) + t < 0.0), lo + 1, lo)
    hi = np.minimum(n - 1, np.floor(float(n - 1) - t)).astype(np.int64)
    hi = np.where((hi < n - 1) & ((hi + 1).astype(np.float64) + t <= float(n - 1)), hi + 1, hi)
    hi = np.where((hi >= 0) & (hi.astype(np.float64) + t > float(n - 1)), hi - 1, hi)
    return lo, hi


def _box_candidate_stats(shape, ts):
    """mvs_pc_candidate_stats for pairs without NaNs, ``ts``: (n, ndim) float64.
    Rows of [n_mask, n_valid, lo zyx, hi zyx]."""
    ts = np.asarray(ts, dtype=np.float64).reshape(-1, len(shape))
    ndim = len(shape)
    out = np.zeros((len(ts), 8), dtype=np.int64)
    count = np.ones(len(ts), dtype=np.int64)
    for d in range(ndim):
        lo, hi = _valid_range(int(shape[d]), ts[:, d])
        count *= np.maximum(0, hi - lo + 1)
        out[:, 2 + 3 - ndim + d] = lo
        out[:, 5 + 3 - ndim + d] = hi
    empty = count == 0
    out[empty, 2:5] = np.iinfo(np.int32).max
    out[empty, 5:8] = -1
    out[:, 0] = out[:, 1] = count
    return out


def _register_loaded(plan, stats, disambiguate_region_mode=None, return_details=False):
    """Stages B-E for the pairs loaded into ``plan``; one result dict per pair."""
    n, ndim, shape = plan.n, plan.ndim, plan.shape
    peaks, updft = plan.correlate()
    shifts = _subpixel_shifts(plan, peaks, updft)

    # per pair: shift candidates in the reference's order ("phase", None[, masked])
    per_pair = []
    for i in range(n):
        has_nan = bool(stats[i, 0, 2] > 0 or stats[i, 1, 2] > 0)
        sc = [shifts[i, 1], shifts[i, 0]]
        if has_nan:
            # the masked call (:433-443) is handed isnan masks (True = invalid) and
            # degenerates to a zero shift ("masked candidate")
            sc.append(np.zeros(ndim))
        mode = disambiguate_region_mode or ("intersection" if has_nan else "union")
        t_cands = _expand_candidates(sc, shape, max(shape))
        per_pair.append({"has_nan": has_nan, "mode": mode, "shift_candidates": sc, "t": t_cands})

    cand_pair = [i for i, pp in enumerate(per_pair) for _ in pp["t"]]
    cand_t = [t for pp in per_pair for t in pp["t"]]
    results = [None] * n
    if not cand_t:
        return [[np.zeros(ndim)] for _ in range(n)]  # registration.py:479-480
    # identical (pair, t) candidates (both normalisations often agree) are evaluated
    # once; the reference evaluates them twice with identical results
    uniq, umap = {}, []
    for cp_, ct_ in zip(cand_pair, cand_t):
        key = (cp_,) + tuple(ct_)
        umap.append(uniq.setdefault(key, len(uniq)))
    ukeys = list(uniq.keys())
    # NaN-free pairs: the valid region of the shifted image is the box the
    # coordinate predicate cuts out, so its statistics have a closed form; only
    # pairs with NaNs need the counting kernel
    cstats_u = np.zeros((len(ukeys), 8), dtype=np.int64)
    need_kernel = [j for j, k in enumerate(ukeys) if per_pair[k[0]]["has_nan"]]
    closed = [j for j, k in enumerate(ukeys) if not per_pair[k[0]]["has_nan"]]
    if closed:
        cstats_u[closed] = _box_candidate_stats(shape, [ukeys[j][1:] for j in closed])
    if need_kernel:
        cstats_u[need_kernel] = plan.candidate_stats(
            [ukeys[j][0] for j in need_kernel], np.array([ukeys[j][1:] for j in need_kernel], dtype=np.float64)
        )
    cstats = cstats_u[np.array(umap)]

    # decide which candidates need SSIM (:501-536), vectorised over all candidates
    n_c = len(cand_pair)
    cp_arr = np.asarray(cand_pair)
    nmask_all = cstats[:, 0]
    valid1 = np.array([int(plan.voxels - stats[i, 1, 2]) for i in range(n)], dtype=np.float64)[cp_arr]
    with np.errstate(divide="ignore", invalid="ignore"):
        low = (nmask_all == 0) | (nmask_all.astype(np.float64) / valid1 < 0.1)
    im0_lo = stats[:, 0, 3 + 3 - ndim : 6].astype(np.int64)[cp_arr]
    im0_hi = stats[:, 0, 6 + 3 - ndim : 9].astype(np.int64)[cp_arr]
    lo1 = cstats[:, 2 + 3 - ndim : 5]
    hi1 = cstats[:, 5 + 3 - ndim : 8]
    union = np.array([pp["mode"] == "union" for pp in per_pair])[cp_arr][:, None]
    lo_all = np.where(union, np.minimum(im0_lo, lo1), np.maximum(im0_lo, lo1))
    hi_all = np.where(union, np.maximum(im0_hi, hi1), np.minimum(im0_hi, hi1)) + 1
    min_shape = (hi_all - lo_all).min(axis=1)
    win_all = np.minimum(7, min_shape - ((min_shape - 1) % 2))
    ssim_req = []  # (flat cand index, slices lo, hi, win)
    pos = 0
    for i, pp in enumerate(per_pair):
        kinds = []
        for _ in pp["t"]:
            if low[pos]:
                kinds.append(("low", None))
            else:
                info = (pos, lo_all[pos], hi_all[pos], int(win_all[pos]))
                kinds.append(("eval", info))
                if info[3] >= 3:
                    ssim_req.append(info)
            pos += 1
        pp["kind"] = kinds
    assert pos == n_c

    ssim_out = {}
    if ssim_req:
        first_req = {}
        for r in ssim_req:
            first_req.setdefault(umap[r[0]], r)
        by_u = {}
        for w in sorted({r[3] for r in first_req.values()}):
            # one batch per window size (the kernels are specialised on it)
            reqs = [r for r in first_req.values() if r[3] == w]
            idx = [r[0] for r in reqs]
            res = plan.candidate_ssim(
                [cand_pair[j] for j in idx],
                np.array([cand_t[j] for j in idx], dtype=np.float64),
                np.array([[r[1], r[2]] for r in reqs]),
                [r[3] for r in reqs],
            )
            by_u.update({umap[j]: r for j, r in zip(idx, res)})
        for r in ssim_req:
            ssim_out[r[0]] = by_u[umap[r[0]]]

    winners = []  # (pair, candidate index whose ranks are needed)
    for i, pp in enumerate(per_pair):
        if not pp["t"]:
            results[i] = [np.zeros(ndim)]  # registration.py:479-480
            continue
        disamb, quality_src = [], []  # quality_src: None (-1) or the candidate index to rank
        for ci, (kind, info) in enumerate(pp["kind"]):
            if kind == "low":
                disamb.append(-1)
                quality_src.append(None)
                continue
            pos_j, lo, hi, win = info
            if win >= 3:
                ssim_val, vmax = ssim_out[pos_j]
                flat = vmax <= 0.0  # np.nanmax(im1t[slices]) <= im1_min (== 0 after rescale)
                if flat:
                    continue  # :530-533 -- nothing appended
                disamb.append(ssim_val)
            else:
                # SSIM window < 3 (:536-538).  The flat-region skip (:530) is not
                # evaluated for such 1-2 px slivers: rescaled, non-constant images
                # are not identically 0 there.
                disamb.append(-1)
            quality_src.append(ci)
        argmax_index = int(np.nanargmax(disamb))
        t = pp["t"][argmax_index]  # same (mis)alignment as the reference when entries were skipped
        src = quality_src[argmax_index]
        res = {"affine_matrix": affine_from_translation(t), "quality": -1}
        if src is not None:
            winners.append((i, src))
        if return_details:
            res["shift_candidates"] = pp["shift_candidates"]
            res["t_candidates"] = pp["t"]
            res["ssim"] = disamb
        results[i] = res
    if winners:
        # quality is read only at the argmax (:558-563): rank just those candidates
        first_pos = np.cumsum([0] + [len(pp["t"]) for pp in per_pair])
        rho = plan.spearman_batch(
            [w[0] for w in winners],
            np.array([per_pair[w[0]]["t"][w[1]] for w in winners], dtype=np.float64),
            [int(cstats[first_pos[w[0]] + w[1]][0]) for w in winners],
        )
        for (i, _), r in zip(winners, rho):
            results[i]["quality"] = float(r)
    return results


def _to_device_f32(a):
    import torch

    if isinstance(a, torch.Tensor):
        return a.to("cuda", dtype=torch.float32).contiguous()
    if hasattr(a, "data") and not isinstance(a, np.ndarray):
        a = a.data  # xr.DataArray-like (registration.py:377-378)
    if hasattr(a, "__cuda_array_interface__"):
        return torch.as_tensor(a, device="cuda").to(torch.float32).contiguous()
    a = np.ascontiguousarray(a, dtype=np.float32)
    t = torch.empty(a.shape, dtype=torch.float32, device="cuda")
    _lib.copy_h2d(t, a)  # pageable host crop -> device through the pinned staging ring
    return t


def _to_device_many(arrays, n):
    """Device float32 tensors of all crops: host arrays travel as ONE staged transfer."""
    import torch

    out, host_idx, host_arr = [None] * len(arrays), [], []
    for i, a in enumerate(arrays):
        if not isinstance(a, torch.Tensor) and hasattr(a, "data") and not isinstance(a, np.ndarray):
            a = a.data  # xr.DataArray-like (registration.py:377-378)
        if isinstance(a, torch.Tensor) or hasattr(a, "__cuda_array_interface__"):
            out[i] = _to_device_f32(a)
        else:
            host_idx.append(i)
            host_arr.append(np.ascontiguousarray(a, dtype=np.float32))
    if host_arr:
        tens = [torch.empty(a.shape, dtype=torch.float32, device="cuda") for a in host_arr]
        _lib.copy_h2d_many(tens, host_arr)
        for i, t in zip(host_idx, tens):
            out[i] = t
    return out[:n], out[n:]


_SPLIT = int(os.environ.get("MVS_REG_SPLIT", "1"))


def register_pairs(fixed_list, moving_list, disambiguate_region_mode=None, upsample_factor=None, return_details=False, plans=None):
    """Batched phase-correlation registration.  ``fixed_list[i]`` /
    ``moving_list[i]`` are same-shape float arrays (host or CUDA, NaN = outside);
    pairs are grouped by shape and each group runs as one batch.  Returns one
    ``{"affine_matrix", "quality"}`` per pair (see
    ``phase_correlation_registration``).  Constant images get the caller-side
    guard's answer (identity, quality NaN, UserWarning; registration.py:1504-1530).
    ``plans``: optional dict reused across calls to keep device buffers/tables."""
    n = len(fixed_list)
    if len(moving_list) != n:
        raise EngineError("fixed_list and moving_list differ in length")
    def _shape(a):
        if hasattr(a, "data") and not hasattr(a, "shape"):
            a = a.data
        return tuple(int(s) for s in a.shape)

    groups = {}
    for i, (f, m) in enumerate(zip(fixed_list, moving_list)):
        if _shape(f) != _shape(m):
            raise EngineError(f"pair {i}: shapes differ {_shape(f)} vs {_shape(m)}")
        groups.setdefault(_shape(f), []).append(i)
    results = [None] * n
    # host crops go up first, through the staging ring (measured: overlapping one shape's upload with
    # another's kernels is slower -- 24-57 ms when the group threads upload, 23-24 ms when this thread
    # uploads group by group -- against 21.6 ms for C2's 40 pairs: the registration threads, the
    # interpreter and the copy pool fight over the host's cores)
    fixed, moving = _to_device_many(list(fixed_list) + list(moving_list), n)

    def run_group(shape, idx, part=0):
        ndim = len(shape)
        u = upsample_factor if upsample_factor is not None else (10 if ndim == 2 else 2)
        key = (shape, u) if part == 0 else (shape, u, part)
        plan = plans.get(key) if plans is not None else None
        if plan is None or plan.max_pairs < len(idx):
            plan = PhaseCorrPlan(shape, len(idx), u)
            if plans is not None:
                plans[key] = plan
        stats = plan.load_pairs([fixed[i] for i in idx], [moving[i] for i in idx])
        constant = [bool(stats[k, 0, 0] == stats[k, 0, 1] or stats[k, 1, 0] == stats[k, 1, 1]) for k in range(len(idx))]
        if any(constant):
            keep = [k for k in range(len(idx)) if not constant[k]]
            for k in range(len(idx)):
                if constant[k]:
                    warnings.warn(
                        "An overlap region between tiles/views is all zero or constant. Assuming identity transform.",
                        UserWarning,
                        stacklevel=2,
                    )
                    results[idx[k]] = {"affine_matrix": np.eye(ndim + 1), "quality": np.nan}
            if keep:
                stats = plan.load_pairs([fixed[idx[k]] for k in keep], [moving[idx[k]] for k in keep])
                out = _register_loaded(plan, stats, disambiguate_region_mode, return_details)
                for k, r in zip(keep, out):
                    results[idx[k]] = r
        else:
            out = _register_loaded(plan, stats, disambiguate_region_mode, return_details)
            for k, r in zip(range(len(idx)), out):
                results[idx[k]] = r
        if plans is None:
            plan.close()

    # large groups are cut into sub-batches that run side by side (own plan, stream and thread):
    # the stages of one batch return to the host between launches, and another batch's kernels
    # fill those gaps
    items = []
    for shape, idx in groups.items():
        parts = max(1, min(_SPLIT, len(idx) // 8))
        for k in range(parts):
            items.append((shape, idx[k::parts], k))
    if len(items) == 1:
        run_group(*items[0])
        return results
    # Several crop shapes (e.g. the x- and the y-neighbours of a tile grid): each group
    # runs on its own stream from its own thread, so the host glue of one group (the
    # stages return to the host between launches) overlaps the kernels of another.
    # ctypes and the synchronising CUDA calls release the GIL.
    import torch
    from concurrent.futures import ThreadPoolExecutor

    cur = torch.cuda.current_stream()
    ready = torch.cuda.Event()
    ready.record(cur)
    device = torch.cuda.current_device()

    def worker(item):
        torch.cuda.set_device(device)
        st = torch.cuda.Stream()
        st.wait_event(ready)  # the crops were produced on the caller's stream
        with torch.cuda.stream(st):
            run_group(*item)
        st.synchronize()

    with ThreadPoolExecutor(max_workers=min(len(items), 8)) as pool:
        for f in [pool.submit(worker, it) for it in items]:
            f.result()
    return results


def phase_correlation_registration(fixed_data, moving_data, disambiguate_region_mode=None, **skimage_phase_corr_kwargs):
    """Drop-in ``pairwise_reg_func`` (registration.py:353-565): translation
    between two same-grid images by phase correlation, sub-pixel refined by an
    upsampled DFT, sign/wrap ambiguity resolved by SSIM, quality = Spearman.

    ``fixed_data`` / ``moving_data``: arrays or ``xr.DataArray``-likes (``.data``
    is used), float, NaN = outside.  Returns ``{"affine_matrix": (ndim+1)^2
    translation mapping fixed px -> moving px, "quality": float}``.
    """
    kwargs = dict(skimage_phase_corr_kwargs)
    u = kwargs.pop("upsample_factor", None)
    if kwargs:
        raise EngineError(f"unsupported phase_cross_correlation arguments: {sorted(kwargs)}")
    return register_pairs([fixed_data], [moving_data], disambiguate_region_mode, u)[0]


def dispatch_pairwise_reg_func(pairwise_reg_func, fixed_data=None, moving_data=None, skip_constant_check=False, **kwargs):
    """registration.py:1477-1544 for image data: the constant-image guard, then
    the hook.  (``register_pairs`` applies the same guard itself.)"""
    if fixed_data is not None and moving_data is not None:
        kwargs["fixed_data"] = fixed_data
        kwargs["moving_data"] = moving_data
    return pairwise_reg_func(**kwargs)


def pairwise_executor(msims, edges, register_kwargs):
    """Hook A of ``registration.register`` (registration.py:2649-2655): see
    ``pairs.pairwise_executor`` (batched pair preparation + registration on the GPU)."""
    from . import pairs

    return pairs.pairwise_executor(msims, edges, register_kwargs)
