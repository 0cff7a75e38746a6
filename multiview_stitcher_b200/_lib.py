"""ctypes binding of the engine's C ABI (include/mvs_b200.h).

The product path has NO CPU fallback: if ``libmvs_b200.so`` is missing or no
CUDA device is usable, calls raise ``EngineUnavailable``.
"""

from __future__ import annotations

import ctypes
import os
import threading

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libmvs_b200.so")

MVS_U8, MVS_U16, MVS_F32 = 0, 1, 2
MVS_FUSE_WAVG, MVS_FUSE_MAX, MVS_FUSE_MEAN = 0, 1, 2

_NP_TO_MVS = {np.dtype(np.uint8): MVS_U8, np.dtype(np.uint16): MVS_U16, np.dtype(np.float32): MVS_F32}
_MVS_TO_NP = {v: k for k, v in _NP_TO_MVS.items()}


class EngineUnavailable(RuntimeError):
    """The CUDA extension is missing / not loadable / has no device."""


class EngineError(RuntimeError):
    """A C-ABI call returned a non-zero status."""


# numpy mirrors of the C structs (align=True reproduces the C layout; checked
# against mvs_struct_sizes() at load time)
VIEW_XFORM_DTYPE = np.dtype(
    [
        ("data", np.uint64),
        ("dtype", np.int32),
        ("shape", np.int32, (3,)),
        ("stride", np.int64, (3,)),
        ("matrix", np.float64, (9,)),
        ("offset", np.float64, (3,)),
        ("wmatrix", np.float64, (9,)),
        ("woffset", np.float64, (3,)),
        ("table", np.int32),
        ("reserved", np.int32),
    ],
    align=True,
)

CHUNK_DTYPE = np.dtype(
    [
        ("out", np.uint64),
        ("out_dtype", np.int32),
        ("shape", np.int32, (3,)),
        ("stride", np.int64, (3,)),
        ("halo", np.int32, (3,)),
        ("first_xform", np.int32),
        ("n_xforms", np.int32),
        ("acc_num", np.uint64),
        ("acc_den", np.uint64),
    ],
    align=True,
)


def mvs_dtype(np_dtype) -> int:
    try:
        return _NP_TO_MVS[np.dtype(np_dtype)]
    except KeyError:
        raise EngineError(f"unsupported voxel dtype {np_dtype} (uint8, uint16, float32)") from None


_lock = threading.Lock()
_lib = None
_device_ok = False

_P = ctypes.c_void_p
_SIGNATURES = {
    "mvs_last_error": (ctypes.c_char_p, []),
    "mvs_abi_version": (ctypes.c_int, []),
    "mvs_device_info": (
        ctypes.c_int,
        [ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)],
    ),
    "mvs_struct_sizes": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]),
    "mvs_fuse_plan_create": (
        ctypes.c_int,
        [ctypes.POINTER(_P), _P, ctypes.c_int, _P, ctypes.c_int, _P, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _P],
    ),
    "mvs_fuse_plan_run": (ctypes.c_int, [_P, _P]),
    "mvs_fuse_plan_run_chunks": (ctypes.c_int, [_P, ctypes.c_int, ctypes.c_int, _P]),
    "mvs_fuse_plan_info": (
        ctypes.c_int,
        [_P, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64)],
    ),
    "mvs_fuse_plan_destroy": (ctypes.c_int, [_P]),
    "mvs_fuse_finalize": (ctypes.c_int, [_P, _P, _P, ctypes.c_int, ctypes.c_int64, _P]),
    "mvs_fuse_finalize_boxes": (ctypes.c_int, [_P, ctypes.c_int, _P]),
    "mvs_resample_views": (ctypes.c_int, [_P, ctypes.c_int, _P, ctypes.c_int, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32), ctypes.c_int, ctypes.c_int, _P, _P, _P]),
    "mvs_normalize_weights": (ctypes.c_int, [_P, ctypes.c_int, ctypes.c_int64, _P]),
    "mvs_gaussian_filter": (ctypes.c_int, [_P, _P, ctypes.c_int, ctypes.POINTER(ctypes.c_int32), ctypes.c_int, _P, ctypes.c_int, _P]),
    "mvs_content_based": (ctypes.c_int, [_P, _P, ctypes.c_int, ctypes.POINTER(ctypes.c_int32), ctypes.c_int, _P, ctypes.c_int, _P, ctypes.c_int, _P, _P]),
    "mvs_content_based_dct": (ctypes.c_int, [_P, ctypes.c_int, ctypes.POINTER(ctypes.c_int32), ctypes.c_int, ctypes.POINTER(ctypes.c_int32), ctypes.c_float, ctypes.c_float, _P, _P]),
    "mvs_convolve": (ctypes.c_int, [_P, _P, ctypes.POINTER(ctypes.c_int32), _P, ctypes.POINTER(ctypes.c_int32), ctypes.c_int, ctypes.c_float, _P]),
    "mvs_mv_deconvolution": (ctypes.c_int, [_P, _P, ctypes.c_int, ctypes.POINTER(ctypes.c_int32), ctypes.c_int, _P, _P, ctypes.POINTER(ctypes.c_int32),
                                            ctypes.c_int, ctypes.c_float, ctypes.c_float, ctypes.c_int, _P, _P]),
    "mvs_fuse_stack": (ctypes.c_int, [_P, _P, _P, ctypes.c_int, ctypes.c_int64, ctypes.c_int, _P, _P]),
    "mvs_trim_cast": (ctypes.c_int, [_P, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32), _P, ctypes.c_int, ctypes.POINTER(ctypes.c_int64), _P]),
    "mvs_pc_plan_create": (ctypes.c_int, [ctypes.POINTER(_P), ctypes.c_int, ctypes.POINTER(ctypes.c_int32), ctypes.c_int, ctypes.c_int]),
    "mvs_pc_plan_destroy": (ctypes.c_int, [_P]),
    "mvs_pc_plan_info": (ctypes.c_int, [_P, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int)]),
    "mvs_pc_debug_copy": (ctypes.c_int, [_P, ctypes.c_int, ctypes.c_int, _P]),
    "mvs_pc_load_pairs": (ctypes.c_int, [_P, ctypes.c_int, _P, _P, _P, _P]),
    "mvs_pc_correlate": (ctypes.c_int, [_P, ctypes.c_int, _P, _P, _P]),
    "mvs_pc_candidate_stats": (ctypes.c_int, [_P, ctypes.c_int, _P, _P, _P, _P]),
    "mvs_pc_candidate_ssim": (ctypes.c_int, [_P, ctypes.c_int, _P, _P, _P, _P, _P, _P]),
    "mvs_pc_spearman_batch": (ctypes.c_int, [_P, ctypes.c_int, _P, _P, _P, _P, _P]),
    "mvs_bin_mean": (
        ctypes.c_int,
        [_P, ctypes.c_int, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int32), ctypes.c_int, _P, _P],
    ),
    "mvs_synth_tile": (
        ctypes.c_int,
        [_P, ctypes.c_int, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64), ctypes.c_uint32, _P],
    ),
    "mvs_copy_h2d_2d": (ctypes.c_int, [_P, ctypes.c_size_t, _P, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, _P]),
    "mvs_copy_d2h_2d": (ctypes.c_int, [_P, ctypes.c_size_t, _P, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, _P]),
    "mvs_copy_h2d_many": (ctypes.c_int, [ctypes.c_int, _P, _P, _P, _P]),
    "mvs_copy_h2d_3d": (ctypes.c_int, [_P, ctypes.c_size_t, ctypes.c_size_t, _P, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, _P]),
    "mvs_copy_d2h_3d": (ctypes.c_int, [_P, ctypes.c_size_t, ctypes.c_size_t, _P, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, _P]),
    "mvs_io_files": (ctypes.c_int, [_P, _P, _P, ctypes.c_int, ctypes.c_int]),
    "mvs_chunks_pack": (ctypes.c_int, [_P, ctypes.c_int, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int32), _P, _P]),
    "mvs_chunks_unpack": (ctypes.c_int, [_P, ctypes.c_int, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int32), _P, _P]),
    "mvs_chunks_store": (ctypes.c_int, [_P, ctypes.c_size_t, ctypes.c_int, _P, _P]),
    "mvs_chunks_load": (ctypes.c_int, [_P, ctypes.c_size_t, ctypes.c_int, _P, _P]),
    "mvs_synth_field": (
        ctypes.c_int,
        [_P, ctypes.c_int, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int64), _P, _P, _P, _P, ctypes.c_int,
         ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_uint32, _P],
    ),
}


def exported_symbols():
    """Names the header declares (tests check the .so exports each)."""
    return sorted(_SIGNATURES)


def load(require_device=False):
    """Load libmvs_b200.so (once).  Raises EngineUnavailable if missing."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise EngineUnavailable(
                    f"{LIB_PATH} not found - build it with `python -m multiview_stitcher_b200.build` "
                    "(there is no CPU fallback)"
                )
            try:
                lib = ctypes.CDLL(LIB_PATH)
            except OSError as e:  # pragma: no cover
                raise EngineUnavailable(f"cannot load {LIB_PATH}: {e}") from e
            for name, (res, args) in _SIGNATURES.items():
                try:
                    fn = getattr(lib, name)
                except AttributeError as e:
                    raise EngineUnavailable(f"{LIB_PATH} does not export {name}") from e
                fn.restype = res
                fn.argtypes = args
            a, b = ctypes.c_int(), ctypes.c_int()
            lib.mvs_struct_sizes(ctypes.byref(a), ctypes.byref(b))
            if a.value != VIEW_XFORM_DTYPE.itemsize or b.value != CHUNK_DTYPE.itemsize:
                raise EngineUnavailable(
                    f"struct layout mismatch: C ({a.value}, {b.value}) vs numpy "
                    f"({VIEW_XFORM_DTYPE.itemsize}, {CHUNK_DTYPE.itemsize})"
                )
            _lib = lib
    if require_device and not _device_ok:
        device_info()
    return _lib


def last_error() -> str:
    return load().mvs_last_error().decode("utf-8", "replace")


def check(status: int, what: str):
    if status != 0:
        raise EngineError(f"{what} failed (status {status}): {last_error()}")


def device_info():
    lib = load()
    name = ctypes.create_string_buffer(256)
    sm, maj, mnr = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    st = lib.mvs_device_info(name, 256, ctypes.byref(sm), ctypes.byref(maj), ctypes.byref(mnr))
    if st != 0:
        raise EngineUnavailable(f"no usable CUDA device: {last_error()}")
    global _device_ok
    _device_ok = True
    return {"name": name.value.decode(), "sm_count": sm.value, "cc": (maj.value, mnr.value)}


def current_stream_ptr():
    import torch

    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _pitched(shape, strides_bytes, itemsize):
    """(outer index tuples, planes, rows, width bytes, row pitch, plane pitch) of an array window
    whose last axis is contiguous: a stack of pitched 3-D copies (one for up to 3 axes)."""
    if len(shape) == 0:
        return [()], 1, 1, itemsize, itemsize, itemsize
    if strides_bytes[-1] != itemsize and shape[-1] > 1:
        raise EngineError("staged copies need a contiguous last axis")
    width = int(shape[-1]) * itemsize
    if len(shape) == 1:
        return [()], 1, 1, width, width, width
    rows, pitch = int(shape[-2]), int(strides_bytes[-2])
    if len(shape) == 2:
        return [()], 1, rows, width, pitch, pitch * rows
    outer = list(np.ndindex(*shape[:-3])) if len(shape) > 3 else [()]
    return outer, int(shape[-3]), rows, width, pitch, int(strides_bytes[-3])


def _staged(fn_name, dev_tensor, host_array, stream_ptr):
    lib = load(require_device=True)
    if tuple(dev_tensor.shape) != tuple(host_array.shape) or dev_tensor.element_size() != host_array.itemsize:
        raise EngineError(f"{fn_name}: shape / item size mismatch")
    es = host_array.itemsize
    outer, planes, rows, width, hp, hpl = _pitched(host_array.shape, host_array.strides, es)
    dstr = [s * es for s in dev_tensor.stride()]
    _, _, _, _, dp, dpl = _pitched(tuple(dev_tensor.shape), dstr, es)
    if min(hp, dp) < width or (planes > 1 and (hpl < hp * rows or dpl < dp * rows)) or (rows > 1 and (hp <= 0 or dp <= 0)):
        raise EngineError(f"{fn_name}: overlapping / negative strides are not supported")
    st = stream_ptr if stream_ptr is not None else current_stream_ptr()
    fn = getattr(lib, fn_name)
    no = len(outer[0])
    for idx in outer:
        ho = sum(i * s for i, s in zip(idx, host_array.strides[:no]))
        do = sum(i * s for i, s in zip(idx, dstr[:no]))
        d = ctypes.c_void_p(dev_tensor.data_ptr() + do)
        h = ctypes.c_void_p(host_array.ctypes.data + ho)
        if fn_name == "mvs_copy_h2d_3d":
            check(fn(d, dp, dpl, h, hp, hpl, width, rows, planes, st), fn_name)
        else:
            check(fn(h, hp, hpl, d, dp, dpl, width, rows, planes, st), fn_name)
    return int(host_array.size) * es


def copy_h2d(dst_tensor, src_array, stream_ptr=None):
    """Pageable numpy window -> CUDA tensor window (same shape, contiguous last axis)
    through the engine's pinned staging lanes.  Returns the bytes moved."""
    return _staged("mvs_copy_h2d_3d", dst_tensor, src_array, stream_ptr)


def copy_d2h(dst_array, src_tensor, stream_ptr=None):
    """CUDA tensor window -> pageable numpy window (blocking)."""
    if not dst_array.flags.writeable:
        raise EngineError("copy_d2h: destination is read-only")
    return _staged("mvs_copy_d2h_3d", src_tensor, dst_array, stream_ptr)


def copy_h2d_many(dst_tensors, src_arrays, stream_ptr=None):
    """Contiguous numpy arrays -> CUDA tensors of the same size, as ONE pipelined transfer through the
    staging ring (per-array calls leave the copy pool idle between arrays).  Returns the bytes moved."""
    lib = load(require_device=True)
    n = len(dst_tensors)
    if n != len(src_arrays):
        raise EngineError("copy_h2d_many: list lengths differ")
    if n == 0:
        return 0
    d = (ctypes.c_void_p * n)()
    h = (ctypes.c_void_p * n)()
    b = (ctypes.c_size_t * n)()
    for i, (t, a) in enumerate(zip(dst_tensors, src_arrays)):
        if not a.flags.c_contiguous or not t.is_contiguous() or t.numel() * t.element_size() != a.nbytes:
            raise EngineError("copy_h2d_many: arrays must be contiguous and of equal size")
        d[i], h[i], b[i] = t.data_ptr(), a.ctypes.data, a.nbytes
    st = stream_ptr if stream_ptr is not None else current_stream_ptr()
    check(lib.mvs_copy_h2d_many(n, d, h, b, st), "mvs_copy_h2d_many")
    return int(sum(b))
