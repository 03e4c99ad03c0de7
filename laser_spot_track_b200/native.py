"""ctypes binding of include/lst_b200.h -- the same stub a JNI/FFM shim would generate.

Loading never falls back to anything: if ``liblst_b200.so`` is missing or a computing entry
point is called without a CUDA device, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liblst_b200.so")

LST_OK, LST_ERR_INVALID, LST_ERR_UNSUPPORTED, LST_ERR_CUDA, LST_ERR_NOMEM, LST_ERR_STATE = 0, -1, -2, -3, -4, -5
ABI_VERSION = 3          # LST_ABI_VERSION of include/lst_b200.h
PIX_U8, PIX_U16, PIX_F32, PIX_RGB = 8, 16, 32, 24
RANGE_DISPLAY, RANGE_CROP, RANGE_FIXED, RANGE_FRAME = 0, 1, 2, 3
LST_WOULD_BLOCK = 1
QUANT_ROUND255, QUANT_TRUNC256 = 0, 1
STATUS_OK, STATUS_SKIP, STATUS_STOP = 0, 1, 2

# every symbol include/lst_b200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "lst_abi_version", "lst_device_count", "lst_create", "lst_destroy", "lst_last_error",
    "lst_set_reference", "lst_track", "lst_track_ptrs", "lst_track_device", "lst_track_device_ptrs", "lst_get_state", "lst_set_state",
    "lst_match_single", "lst_match_test", "lst_preprocess", "lst_calc_displacement",
    "lst_alloc_pinned", "lst_free_pinned", "lst_host_register", "lst_host_unregister",
    "lst_device_alloc", "lst_device_free", "lst_device_upload",
    "lst_get_stats", "lst_reset_stats", "lst_set_kernel_timing", "lst_measure_alu_peaks",
    "lst_host_window_for", "lst_host_template_rect", "lst_host_resolve",
    "lst_multi_create", "lst_multi_destroy", "lst_multi_last_error", "lst_multi_set_reference", "lst_multi_track",
    "lst_multi_get_state", "lst_multi_set_state", "lst_multi_retracked", "lst_multi_track_device",
    "lst_multi_get_stats", "lst_multi_reset_stats",
    "lst_tiff_probe", "lst_tiff_read_rect", "lst_set_reference_file", "lst_track_files", "lst_results_rows",
    "lst_set_frame_ranges", "lst_submit", "lst_fetch", "lst_override", "lst_set_queue_depth",
]


class Params(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("width", C.c_int32), ("height", C.c_int32), ("pixel_type", C.c_int32),
        ("method", C.c_int32), ("templ_size", C.c_int32), ("s_area", C.c_int32), ("sub_pixel", C.c_int32),
        ("match_intensity", C.c_int32), ("range_mode", C.c_int32), ("quant_mode", C.c_int32),
        ("doubling", C.c_int32), ("auto_skip", C.c_int32), ("max_s_area", C.c_int32), ("max_batch", C.c_int32),
        ("ingest_mode", C.c_int32),
        ("mark_dist", C.c_double), ("range_lo", C.c_double), ("range_hi", C.c_double),
        ("match_threshold", C.c_double), ("ncc_engine", C.c_int32), ("reserved1", C.c_int32),
        ("rgb_weights", C.c_double * 3),
    ]


class TemplateMatch(C.Structure):
    _fields_ = [
        ("x", C.c_double), ("y", C.c_double), ("score", C.c_double),
        ("ix", C.c_int32), ("iy", C.c_int32), ("wx", C.c_int32), ("wy", C.c_int32), ("ww", C.c_int32),
        ("wh", C.c_int32), ("s_used", C.c_int32), ("accepted", C.c_int32),
    ]


class Record(C.Structure):
    _fields_ = [
        ("frame", C.c_int64), ("status", C.c_int32), ("reject", C.c_int32),
        ("tm", TemplateMatch * 5),
        ("dis", C.c_double * 10),
        ("dX_pix", C.c_double), ("dY_pix", C.c_double),
        ("x_abs", C.c_double), ("y_abs", C.c_double), ("dL", C.c_double),
        ("passes", C.c_int32), ("reserved", C.c_int32),
    ]


class ReferenceInfo(C.Structure):
    _fields_ = [("rect", (C.c_int32 * 4) * 5), ("mideal", C.c_double * 5), ("ref_record", Record)]


class Stats(C.Structure):
    _fields_ = [
        ("frames", C.c_int64), ("tasks", C.c_int64), ("respeculated", C.c_int64), ("passes", C.c_int64),
        ("kernel_launches", C.c_int64), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
        ("ncc_kernel_ms", C.c_double), ("ncc_kernel_launches", C.c_int64), ("ncc_algorithmic_macs", C.c_double),
        ("roi_fallback_tasks", C.c_int64), ("device_wait_ms", C.c_double), ("host_ms", C.c_double),
        ("preprocess_ms", C.c_double), ("box_sums_ms", C.c_double), ("epilogue_ms", C.c_double),
        ("preprocess_px", C.c_double), ("preprocess_launches", C.c_int64),
    ]


class TiffInfo(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("bits", C.c_int32), ("sample_format", C.c_int32),
                ("big_endian", C.c_int32), ("rows_per_strip", C.c_int32), ("n_strips", C.c_int32),
                ("imagej_images", C.c_int32), ("compression", C.c_int32), ("predictor", C.c_int32),
                ("first_strip_offset", C.c_int64)]


class Task(C.Structure):
    _fields_ = [("frame", C.c_int32), ("tpl", C.c_int32), ("wx", C.c_int32), ("wy", C.c_int32),
                ("ww", C.c_int32), ("wh", C.c_int32), ("s_used", C.c_int32), ("reserved", C.c_int32)]


class TaskResult(C.Structure):
    _fields_ = [("x", C.c_double), ("y", C.c_double), ("score", C.c_double),
                ("ix", C.c_int32), ("iy", C.c_int32), ("accepted", C.c_int32), ("reserved", C.c_int32)]


COMPUTE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(Task), C.c_int32, C.POINTER(TaskResult))

_lib = None


def load() -> C.CDLL:
    """Load liblst_b200.so (built in-tree by laser_spot_track_b200.build).  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m laser_spot_track_b200.build` "
                           "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_double, C.c_size_t
    P = C.POINTER
    sig = {
        "lst_abi_version": ([], C.c_int),
        "lst_device_count": ([], C.c_int),
        "lst_create": ([P(Params), C.c_int, P(vp)], C.c_int),
        "lst_destroy": ([vp], None),
        "lst_last_error": ([vp], C.c_char_p),
        "lst_set_reference": ([vp, vp, P(C.c_float), P(ReferenceInfo)], C.c_int),
        "lst_track": ([vp, vp, sz, i64, i32, P(Record)], C.c_int),
        "lst_track_ptrs": ([vp, P(vp), i64, i32, P(Record)], C.c_int),
        "lst_track_device": ([vp, vp, sz, i64, i32, P(Record)], C.c_int),
        "lst_track_device_ptrs": ([vp, P(vp), i64, i32, P(Record)], C.c_int),
        "lst_get_state": ([vp, P(dbl)], C.c_int),
        "lst_set_state": ([vp, P(dbl)], C.c_int),
        "lst_match_single": ([vp, vp, i32, i32, vp, i32, i32, i32, i32, P(dbl), vp], C.c_int),
        "lst_match_test": ([vp, vp, i32, i32, i32, P(dbl)], C.c_int),
        "lst_preprocess": ([vp, vp, i32, i32, i32, i32, vp], C.c_int),
        "lst_calc_displacement": ([P(C.c_float), P(dbl), dbl, P(dbl), P(dbl)], C.c_int),
        "lst_alloc_pinned": ([sz, P(vp)], C.c_int),
        "lst_free_pinned": ([vp], C.c_int),
        "lst_host_register": ([vp, sz], C.c_int),
        "lst_host_unregister": ([vp], C.c_int),
        "lst_device_alloc": ([vp, sz, P(vp)], C.c_int),
        "lst_device_free": ([vp, vp], C.c_int),
        "lst_device_upload": ([vp, vp, vp, sz], C.c_int),
        "lst_get_stats": ([vp, P(Stats)], C.c_int),
        "lst_reset_stats": ([vp], C.c_int),
        "lst_set_kernel_timing": ([vp, i32], C.c_int),
        "lst_measure_alu_peaks": ([C.c_int, dbl, P(dbl)], C.c_int),
        "lst_host_window_for": ([P(Params), P(i32), dbl, dbl, P(i32)], C.c_int),
        "lst_host_template_rect": ([P(Params), C.c_float, C.c_float, P(i32)], C.c_int),
        "lst_results_rows": ([P(Record), i64, dbl, dbl, i64, P(dbl), P(i64)], C.c_int64),
        "lst_tiff_probe": ([C.c_char_p, i32, P(TiffInfo), C.c_char_p, i32], C.c_int),
        "lst_tiff_read_rect": ([C.c_char_p, i32, i32, i32, i32, i32, vp, C.c_char_p, i32], C.c_int),
        "lst_set_reference_file": ([vp, C.c_char_p, i32, P(C.c_float), P(ReferenceInfo)], C.c_int),
        "lst_track_files": ([vp, P(C.c_char_p), P(i32), i64, i32, P(Record)], C.c_int),
        "lst_multi_create": ([P(Params), P(i32), i32, P(vp)], C.c_int),
        "lst_multi_destroy": ([vp], None),
        "lst_multi_last_error": ([vp], C.c_char_p),
        "lst_multi_set_reference": ([vp, vp, P(C.c_float), P(ReferenceInfo)], C.c_int),
        "lst_multi_track": ([vp, P(vp), i64, i32, P(Record)], C.c_int),
        "lst_multi_track_device": ([vp, P(vp), i64, i32, P(Record)], C.c_int),
        "lst_multi_get_stats": ([vp, P(Stats)], C.c_int),
        "lst_multi_reset_stats": ([vp], C.c_int),
        "lst_multi_get_state": ([vp, P(dbl)], C.c_int),
        "lst_multi_set_state": ([vp, P(dbl)], C.c_int),
        "lst_multi_retracked": ([vp], C.c_int64),
        "lst_set_frame_ranges": ([vp, P(dbl), i32], C.c_int),
        "lst_submit": ([vp, P(vp), i64, i32], C.c_int),
        "lst_fetch": ([vp, P(Record), i32, P(i32), i32], C.c_int),
        "lst_override": ([vp, i64, P(dbl), P(i64)], C.c_int),
        "lst_set_queue_depth": ([vp, i32], C.c_int),
        "lst_host_resolve": ([P(Params), P((i32 * 4) * 5), P(C.c_float), P(dbl), P(dbl), i64, i32, COMPUTE_FN, vp,
                              P(Record), P(Stats)], C.c_int),
    }
    for name, (args, res) in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = res
    _lib = lib
    return lib


def default_params(width: int, height: int, pixel_type: int, **kw) -> Params:
    """Defaults of the dialog, LST4:82-89 / 147."""
    p = Params()
    p.struct_size = C.sizeof(Params)
    p.width, p.height, p.pixel_type = width, height, pixel_type
    p.method, p.templ_size, p.s_area = 5, 120, 50
    p.sub_pixel, p.match_intensity = 1, 1
    p.range_mode, p.quant_mode = RANGE_DISPLAY, QUANT_ROUND255
    p.doubling, p.auto_skip, p.max_s_area, p.max_batch = 0, 1, 500, 0
    p.mark_dist, p.range_lo, p.range_hi = 100.0, 0.0, 255.0
    p.match_threshold = 0.0   # 0 = matchThreshold[method] of LST4:147
    p.ncc_engine = 0
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        if k == "rgb_weights":
            v = (C.c_double * 3)(*[float(x) for x in (v or (0.0, 0.0, 0.0))])
        setattr(p, k, v)
    return p
