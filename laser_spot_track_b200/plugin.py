"""Host-side mirror of the reference plugin's interface for the tracking path.

The reference is an ImageJ ``PlugInFilter`` (one Java class); its toolchain (JVM, Maven) is not
in this image, so the caller's side of the C ABI is written in Python, keeping the reference's
method names, argument meaning and error behaviour:

=====================================  =========================================================
reference (Laser_Spot_Track4.java)     here
=====================================  =========================================================
getUserParameters()        LST4:2370   ``LaserSpotTrack4(width, height, bit_depth, **dialog)``
template selection         LST4:486    ``set_reference(ref_slice, points)``
for-loop + analyseSlice    LST4:793    ``track(stack)`` / ``analyseSlice(slice_pixels)``
doMatch_coord_res (static) LST4:2502   ``doMatch_coord_res(src8, tpl8, method, subPix)``
doMatch_test (static)      LST4:2724   ``doMatch_test(src8, method)``
calcDisplacement()         LST4:2318   ``calcDisplacement(ref_xy, dis, markDist, spot0)``
ResultsTable rows          LST4:864    ``results_table(records)``
=====================================  =========================================================

Everything computes on the GPU through ``liblst_b200.so``; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import native as N

_PIX = {8: (N.PIX_U8, np.uint8), 16: (N.PIX_U16, np.uint16), 32: (N.PIX_F32, np.float32),
        24: (N.PIX_RGB, np.uint32)}   # 24: ColorProcessor pixels, 0x00RRGGBB per pixel; (h, w, 3) uint8 RGB arrays are packed
TEMPLATE_NAMES = ("spot", "mark1", "mark2", "mark3", "mark4")
RESULT_COLUMNS = ("Time", "Frame", "File", "dX_pix", "dY_pix", "X_abs", "Y_abs", "dL")   # LST4:867-875


class LstError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"lst_b200 error {code}: {msg}")
        self.code = code


class LaserSpotTrack4:
    """One tracking session on one GPU (one ``lst_ctx``)."""

    def __init__(self, width: int, height: int, bit_depth: int = 8, *, device: int = 0,
                 method: int = 5, templSize: int = 120, sArea: int = 50, markDist: float = 100.0,
                 subPixel: bool = True, matchIntensity: bool = True, alwaysAutoSkip: bool = True,
                 doubling: bool = False, maxSArea: int = 500, range_mode: int = N.RANGE_DISPLAY,
                 range_lo: float = 0.0, range_hi: float = 255.0, quant_mode: int = N.QUANT_ROUND255,
                 max_batch: int = 0, ingest_mode: int = 0, ncc_engine: int = 0, matchThreshold: float = 0.0,
                 rgb_weights=None):
        if bit_depth not in _PIX:
            raise LstError(N.LST_ERR_UNSUPPORTED, "Unsupported image type")       # IJ.error, LST4:2537
        self._lib = N.load()
        self.width, self.height, self.bit_depth = width, height, bit_depth
        self._np_dtype = _PIX[bit_depth][1]
        self.params = N.default_params(
            width, height, _PIX[bit_depth][0], method=method, templ_size=templSize, s_area=sArea,
            mark_dist=markDist, sub_pixel=int(subPixel), match_intensity=int(matchIntensity),
            auto_skip=int(alwaysAutoSkip), doubling=int(doubling), max_s_area=maxSArea,
            range_mode=range_mode, range_lo=range_lo, range_hi=range_hi, quant_mode=quant_mode,
            max_batch=max_batch, ingest_mode=ingest_mode, ncc_engine=ncc_engine, match_threshold=matchThreshold,
            rgb_weights=rgb_weights)
        self._ctx = C.c_void_p()
        rc = self._lib.lst_create(C.byref(self.params), device, C.byref(self._ctx))
        if rc != N.LST_OK:
            msg = self._lib.lst_last_error(None).decode()
            self._ctx = None
            raise LstError(rc, msg)
        self.reference: Optional[N.ReferenceInfo] = None
        self.ref_xy: Optional[np.ndarray] = None

    # -- lifecycle --------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.lst_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc: int):
        if rc != N.LST_OK:
            raise LstError(rc, self._lib.lst_last_error(self._ctx).decode())

    def _frame(self, a: np.ndarray) -> np.ndarray:
        if self._np_dtype == np.uint32 and a.dtype == np.uint8 and a.shape[-1] == 3:
            a = (a[..., 0].astype(np.uint32) << 16) | (a[..., 1].astype(np.uint32) << 8) | a[..., 2].astype(np.uint32)
        a = np.ascontiguousarray(a, dtype=self._np_dtype)
        if a.shape[-2:] != (self.height, self.width):
            raise LstError(N.LST_ERR_INVALID, "slice size differs from the stack's (LST4:813-817)")
        return a

    # -- template selection: LST4:486-702 -----------------------------------------------------
    def set_frame_ranges(self, lo_hi) -> None:
        """LST_RANGE_FRAME: [min, max] of the slices of the NEXT call that takes slices (slice_proc.getMin() / getMax(),
        LST4:1308-1312); lo_hi = n pairs.  Without it the library reduces every slice on the device."""
        a = np.ascontiguousarray(lo_hi, dtype=np.float64).reshape(-1, 2)
        self._check(self._lib.lst_set_frame_ranges(self._ctx, a.ctypes.data_as(C.POINTER(C.c_double)), len(a)))

    def set_reference(self, ref_slice: np.ndarray, points: Sequence[Tuple[float, float]]) -> N.ReferenceInfo:
        """points: click points of spot, mark1, mark2, mark3, mark4 (x, y)."""
        fr = self._frame(ref_slice)
        xy = np.asarray(points, dtype=np.float32).reshape(10)
        info = N.ReferenceInfo()
        self._check(self._lib.lst_set_reference(self._ctx, fr.ctypes.data, xy.ctypes.data_as(C.POINTER(C.c_float)),
                                                C.byref(info)))
        self.reference, self.ref_xy = info, xy
        return info

    # -- the loop body ---------------------------------------------------------------------------
    def track(self, stack: np.ndarray, first_frame_index: int = 1) -> "C.Array[N.Record]":
        """analyseSlice + calcDisplacement for every slice of ``stack`` ([n, H, W]) in order."""
        st = self._frame(stack)
        if st.ndim == 2:
            st = st[None]
        n = st.shape[0]
        out = (N.Record * n)()
        self._check(self._lib.lst_track(self._ctx, st.ctypes.data, 0, first_frame_index, n, out))
        return out

    def track_slices(self, slices: Iterable[np.ndarray], first_frame_index: int = 1):
        """Same, from separately allocated slices (stack.getProcessor(i).getPixels())."""
        fr = [self._frame(s) for s in slices]
        n = len(fr)
        ptrs = (C.c_void_p * n)(*[f.ctypes.data for f in fr])
        out = (N.Record * n)()
        self._check(self._lib.lst_track_ptrs(self._ctx, ptrs, first_frame_index, n, out))
        return out

    def track_device(self, dev_ptr: int, n: int, frame_stride_bytes: int = 0, first_frame_index: int = 1):
        out = (N.Record * n)()
        self._check(self._lib.lst_track_device(self._ctx, C.c_void_p(dev_ptr), frame_stride_bytes,
                                               first_frame_index, n, out))
        return out

    def track_device_ptrs(self, dev_ptrs: Sequence[int], first_frame_index: int = 1, out=None):
        n = len(dev_ptrs)
        ptrs = dev_ptrs if isinstance(dev_ptrs, C.Array) else (C.c_void_p * n)(*dev_ptrs)
        out = (N.Record * n)() if out is None else out
        self._check(self._lib.lst_track_device_ptrs(self._ctx, ptrs, first_frame_index, n, out))
        return out

    # ---- without blocking the caller (submit / fetch / override; the loop of LST4:793-902 keeps decoding) ----
    def submit(self, host_ptrs, first_frame_index: int) -> bool:
        """Queue a batch of slices (raw host addresses, valid until their records are fetched).  False = the queue is
        full (LST_WOULD_BLOCK): fetch first."""
        n = len(host_ptrs)
        ptrs = host_ptrs if isinstance(host_ptrs, C.Array) else (C.c_void_p * n)(*host_ptrs)
        rc = self._lib.lst_submit(self._ctx, ptrs, first_frame_index, n)
        if rc == N.LST_WOULD_BLOCK:
            return False
        self._check(rc)
        return True

    def fetch(self, max_records: int, wait: bool = False):
        """Finished records in slice order (possibly none)."""
        out = (N.Record * max_records)()
        n = C.c_int32(0)
        rc = self._lib.lst_fetch(self._ctx, out, max_records, C.byref(n), int(wait))
        if rc != N.LST_WOULD_BLOCK:
            self._check(rc)
        return [out[i] for i in range(n.value)]

    def override(self, frame_index: int, dis) -> int:
        """The caller's verdict on slice ``frame_index`` (failureQuestionDlg "keep", LST4:1281-1300): later slices are
        dropped, the state becomes ``dis``; returns the index to submit again from."""
        d = (C.c_double * 10)(*[float(v) for v in dis])
        resume = C.c_int64(0)
        self._check(self._lib.lst_override(self._ctx, frame_index, d, C.byref(resume)))
        return int(resume.value)

    def track_host_ptrs(self, host_ptrs, first_frame_index: int = 1, out=None):
        """Slices given as raw host addresses (e.g. a pinned ring): no numpy conversion on the way."""
        n = len(host_ptrs)
        ptrs = host_ptrs if isinstance(host_ptrs, C.Array) else (C.c_void_p * n)(*host_ptrs)
        out = (N.Record * n)() if out is None else out
        self._check(self._lib.lst_track_ptrs(self._ctx, ptrs, first_frame_index, n, out))
        return out

    # -- slices as files ---------------------------------------------------------------
    def set_reference_file(self, path: str, ref_xy, page: int = 0) -> N.ReferenceInfo:
        """``set_reference`` with the reference slice read from an uncompressed grey-scale TIFF file."""
        xy = (C.c_float * 10)(*[float(v) for p in ref_xy for v in p])
        info = N.ReferenceInfo()
        self._check(self._lib.lst_set_reference_file(self._ctx, os.fsencode(path), int(page), xy, C.byref(info)))
        return info

    def track_files(self, paths, pages=None, first_frame_index: int = 1):
        """Track the slices stored in ``paths`` (one uncompressed grey-scale TIFF per slice, or pages of
        a multi-page / ImageJ stack file with ``pages``): what the reference's virtual stack feeds
        analyseSlice (LST4:808-831).  Only the rows of the five search regions are read from each file."""
        n = len(paths)
        arr = (C.c_char_p * n)(*[os.fsencode(p) for p in paths])
        pg = None if pages is None else (C.c_int32 * n)(*[int(p) for p in pages])
        out = (N.Record * n)()
        self._check(self._lib.lst_track_files(self._ctx, arr, pg, first_frame_index, n, out))
        return out

    def analyseSlice(self, slice_pixels: np.ndarray, frame_index: int = 0) -> N.Record:
        """One slice; returns the record whose ``status`` is analyseSlice's return value (LST4:1302)."""
        return self.track(slice_pixels, frame_index)[0]

    # -- state --------------------------------------------------------------------------------------
    def get_state(self) -> np.ndarray:
        d = (C.c_double * 10)()
        self._check(self._lib.lst_get_state(self._ctx, d))
        return np.array(d[:])

    def set_state(self, dis: Sequence[float]):
        d = (C.c_double * 10)(*[float(v) for v in dis])
        self._check(self._lib.lst_set_state(self._ctx, d))

    # -- the reference's public statics ----------------------------------------------------------------
    def doMatch_coord_res(self, src8: np.ndarray, tpl8: np.ndarray, method: int = 5, subPix: bool = True,
                          want_map: bool = False):
        """doMatch_coord_res(src, tpl, method, subPix, null) on 8-bit images -> [x, y, score] (LST4:2502)."""
        s = np.ascontiguousarray(src8, dtype=np.uint8)
        t = np.ascontiguousarray(tpl8, dtype=np.uint8)
        out = (C.c_double * 3)()
        m = None
        mp = None
        if want_map:
            m = np.zeros((s.shape[0] - t.shape[0] + 1, s.shape[1] - t.shape[1] + 1), dtype=np.float32)
            mp = m.ctypes.data
        self._check(self._lib.lst_match_single(self._ctx, s.ctypes.data, s.shape[1], s.shape[0], t.ctypes.data,
                                               t.shape[1], t.shape[0], int(method), int(subPix), out, mp))
        res = [out[0], out[1], out[2]]
        return (res, m) if want_map else res

    def doMatch_test(self, src8: np.ndarray, method: int = 5) -> float:
        s = np.ascontiguousarray(src8, dtype=np.uint8)
        out = C.c_double()
        self._check(self._lib.lst_match_test(self._ctx, s.ctypes.data, s.shape[1], s.shape[0], int(method), C.byref(out)))
        return out.value

    def preprocess(self, frame: np.ndarray, x0: int, y0: int, w: int, h: int) -> np.ndarray:
        """The 8-bit image handed to matchTemplate for one rectangle of ``frame``."""
        fr = self._frame(frame)
        rgb3 = self.params.pixel_type == N.PIX_RGB and not self.params.match_intensity
        out = np.zeros((3, h, w) if rgb3 else (h, w), dtype=np.uint8)
        self._check(self._lib.lst_preprocess(self._ctx, fr.ctypes.data, x0, y0, w, h, out.ctypes.data))
        return np.ascontiguousarray(out.transpose(1, 2, 0)) if rgb3 else out   # 3-channel match: (h, w, 3), channels R, G, B

    @staticmethod
    def calcDisplacement(ref_xy, dis, markDist: float, spot0=None):
        lib = N.load()
        r = np.asarray(ref_xy, dtype=np.float32).reshape(10)
        d = (C.c_double * 10)(*[float(v) for v in np.asarray(dis).reshape(10)])
        s0 = (C.c_double * 2)(*spot0) if spot0 is not None else None
        out = (C.c_double * 5)()
        rc = lib.lst_calc_displacement(r.ctypes.data_as(C.POINTER(C.c_float)), d, float(markDist), s0, out)
        if rc != N.LST_OK:
            raise LstError(rc, "lst_calc_displacement")
        return list(out)

    # -- diagnostics ------------------------------------------------------------------------------------
    def stats(self) -> N.Stats:
        s = N.Stats()
        self._check(self._lib.lst_get_stats(self._ctx, C.byref(s)))
        return s

    def reset_stats(self):
        self._check(self._lib.lst_reset_stats(self._ctx))

    def set_kernel_timing(self, on: bool):
        self._check(self._lib.lst_set_kernel_timing(self._ctx, int(on)))

    def device_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        self._check(self._lib.lst_device_alloc(self._ctx, nbytes, C.byref(p)))
        return p.value

    def device_free(self, ptr: int):
        self._check(self._lib.lst_device_free(self._ctx, C.c_void_p(ptr)))

    def device_upload(self, dst: int, src: np.ndarray):
        src = np.ascontiguousarray(src)
        self._check(self._lib.lst_device_upload(self._ctx, C.c_void_p(dst), src.ctypes.data, src.nbytes))


class LaserSpotTrack4Multi:
    """The same session spread over several GPUs of one process (lst_multi_*): slices are cut into
    contiguous chunks, one per device; records are identical to the single-GPU ones."""

    def __init__(self, width: int, height: int, bit_depth: int, devices: Sequence[int], **dialog):
        proto = LaserSpotTrack4.__new__(LaserSpotTrack4)
        self._lib = N.load()
        self.width, self.height, self._np_dtype = width, height, _PIX[bit_depth][1]
        kw = dict(method=5, templSize=120, sArea=50, markDist=100.0, subPixel=True, matchIntensity=True,
                  alwaysAutoSkip=True, doubling=False, maxSArea=500, range_mode=N.RANGE_DISPLAY, range_lo=0.0,
                  range_hi=255.0, quant_mode=N.QUANT_ROUND255, max_batch=0, ingest_mode=0, ncc_engine=0)
        kw.update(dialog)
        self.params = N.default_params(
            width, height, _PIX[bit_depth][0], method=kw["method"], templ_size=kw["templSize"], s_area=kw["sArea"],
            mark_dist=kw["markDist"], sub_pixel=int(kw["subPixel"]), match_intensity=int(kw["matchIntensity"]),
            auto_skip=int(kw["alwaysAutoSkip"]), doubling=int(kw["doubling"]), max_s_area=kw["maxSArea"],
            range_mode=kw["range_mode"], range_lo=kw["range_lo"], range_hi=kw["range_hi"],
            quant_mode=kw["quant_mode"], max_batch=kw["max_batch"], ingest_mode=kw["ingest_mode"],
            ncc_engine=kw["ncc_engine"])
        dev = (C.c_int32 * len(devices))(*devices)
        self._m = C.c_void_p()
        rc = self._lib.lst_multi_create(C.byref(self.params), dev, len(devices), C.byref(self._m))
        if rc != N.LST_OK:
            self._m = None
            raise LstError(rc, self._lib.lst_last_error(None).decode())

    def close(self):
        if getattr(self, "_m", None):
            self._lib.lst_multi_destroy(self._m)
            self._m = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != N.LST_OK:
            raise LstError(rc, self._lib.lst_multi_last_error(self._m).decode())

    def set_reference(self, ref_slice, points):
        fr = np.ascontiguousarray(ref_slice, dtype=self._np_dtype)
        xy = np.asarray(points, dtype=np.float32).reshape(10)
        info = N.ReferenceInfo()
        self._check(self._lib.lst_multi_set_reference(self._m, fr.ctypes.data, xy.ctypes.data_as(C.POINTER(C.c_float)),
                                                      C.byref(info)))
        return info

    def track(self, stack, first_frame_index: int = 1):
        st = np.ascontiguousarray(stack, dtype=self._np_dtype)
        n = st.shape[0]
        ptrs = (C.c_void_p * n)(*[st[i].ctypes.data for i in range(n)])
        out = (N.Record * n)()
        self._check(self._lib.lst_multi_track(self._m, ptrs, first_frame_index, n, out))
        return out

    def track_host_ptrs(self, host_ptrs, first_frame_index: int = 1, out=None):
        n = len(host_ptrs)
        ptrs = host_ptrs if isinstance(host_ptrs, C.Array) else (C.c_void_p * n)(*host_ptrs)
        out = (N.Record * n)() if out is None else out
        self._check(self._lib.lst_multi_track(self._m, ptrs, first_frame_index, n, out))
        return out

    def track_device_ptrs(self, dev_ptrs, first_frame_index: int = 1, out=None):
        """Slices already in device memory; with a device listed twice in ``devices`` the two lanes overlap
        each other's host-side chain walks on that GPU."""
        n = len(dev_ptrs)
        ptrs = dev_ptrs if isinstance(dev_ptrs, C.Array) else (C.c_void_p * n)(*dev_ptrs)
        out = (N.Record * n)() if out is None else out
        self._check(self._lib.lst_multi_track_device(self._m, ptrs, first_frame_index, n, out))
        return out

    def get_state(self):
        d = (C.c_double * 10)()
        self._check(self._lib.lst_multi_get_state(self._m, d))
        return np.array(d[:])

    def set_state(self, dis):
        d = (C.c_double * 10)(*[float(v) for v in dis])
        self._check(self._lib.lst_multi_set_state(self._m, d))

    def stats(self) -> N.Stats:
        s = N.Stats()
        self._check(self._lib.lst_multi_get_stats(self._m, C.byref(s)))
        return s

    def reset_stats(self):
        self._check(self._lib.lst_multi_reset_stats(self._m))

    @property
    def retracked(self) -> int:
        return int(self._lib.lst_multi_retracked(self._m))


def tiff_probe(path: str, page: int = 0) -> N.TiffInfo:
    """Layout of a slice file as ``lst_track_files`` sees it (host only); raises LstError when unsupported."""
    info = N.TiffInfo()
    err = C.create_string_buffer(256)
    rc = N.load().lst_tiff_probe(os.fsencode(path), int(page), C.byref(info), err, 256)
    if rc != N.LST_OK:
        raise LstError(rc, err.value.decode(errors="replace"))
    return info


def tiff_read(path: str, page: int = 0, rect=None) -> np.ndarray:
    """The pixels of a slice file (or of the rectangle ``(x, y, w, h)`` of it) as the reader threads of
    ``lst_track_files`` decode them: host only."""
    info = tiff_probe(path, page)
    x, y, w, h = rect if rect is not None else (0, 0, info.width, info.height)
    dt = {8: np.uint8, 16: np.uint16, 32: np.float32, 24: np.uint32}[info.bits]   # 24: ColorProcessor pixels 0x00RRGGBB (3-component JPEG)
    out = np.zeros((h, w), dtype=dt)
    err = C.create_string_buffer(256)
    rc = N.load().lst_tiff_read_rect(os.fsencode(path), int(page), int(x), int(y), int(w), int(h), out.ctypes.data, err, 256)
    if rc != N.LST_OK:
        raise LstError(rc, err.value.decode(errors="replace"))
    return out


def results_rows(records, time0: float = 0.0, time_step: float = 1.0, first_row_number: int = 2):
    """``lst_results_rows``: the numeric ResultsTable columns {Time, Frame, dX_pix, dY_pix, X_abs, Y_abs, dL}
    of a whole ctypes array of records at once (LST4:864-875); returns (rows[n, 7], source index[n])."""
    n = len(records)
    rows = np.zeros((max(n, 1), 7), dtype=np.float64)
    src = np.zeros(max(n, 1), dtype=np.int64)
    k = N.load().lst_results_rows(records, n, float(time0), float(time_step), int(first_row_number),
                                  rows.ctypes.data_as(C.POINTER(C.c_double)), src.ctypes.data_as(C.POINTER(C.c_int64)))
    if k < 0:
        raise LstError(int(k), "lst_results_rows failed")
    return rows[:k], src[:k]


def results_table(records, ref_info: Optional[N.ReferenceInfo] = None, time_step: float = 1.0,
                  file_label: str = "") -> List[dict]:
    """The ResultsTable rows the plugin's loop appends (LST4:707-718 row 0, LST4:864-875):
    skipped frames add no row; Time advances by the constant step (no EXIF) per analysed frame."""
    rows = []
    if ref_info is not None:
        r = ref_info.ref_record
        rows.append(dict(Time=0.0, Frame="1", File=file_label, dX_pix=r.dX_pix, dY_pix=r.dY_pix,
                         X_abs=r.x_abs, Y_abs=r.y_abs, dL=r.dL))
    seconds = 0.0
    for r in records:
        if r.status != N.STATUS_OK:
            continue
        seconds += time_step
        rows.append(dict(Time=seconds, Frame=str(len(rows) + 1), File=file_label, dX_pix=r.dX_pix,
                         dY_pix=r.dY_pix, X_abs=r.x_abs, Y_abs=r.y_abs, dL=r.dL))
    return rows
