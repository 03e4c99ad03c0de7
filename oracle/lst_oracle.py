"""CPU oracle for the per-frame tracking hot path of anotherche/laser-spot-track.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  The product path
(``laser_spot_track_b200``) never imports this module and fails loudly when its CUDA
library is missing.

PARITY UNPINNED.  The reference ships no tests, fixtures or golden vectors and cannot be
built or run here (no JVM / Maven; dialog driven).  The arithmetic of the path lives in
un-vendored third-party code:

* ``net.imagej:ij`` (managed by pom-scijava 36.0.0, pom.xml:8-13,88-91) -- crop,
  ``ImageConverter.convertToGray32``, ``GaussianBlur.blurGaussian``,
  ``FloatProcessor.getBufferedImage`` (float -> 8-bit).  Restated below from the
  published ImageJ sources *from recollection*; every such function says so.
* ``org.bytedeco:javacv-platform:1.5.8`` -> OpenCV 4.6.0 (pom.xml:94-96,277) --
  ``matchTemplate(TM_CCOEFF_NORMED)``, ``minMaxLoc``.  Restated below
  (``ncc_map_exact``) and pinned against the OpenCV that *is* in this image
  (``cv2`` 4.13, ``ncc_map_cv2``): same ``templmatch.cpp`` algorithm family.

Everything that is the reference's own arithmetic is restated from
``/root/reference/src/main/java/laser_spot_track4/Laser_Spot_Track4.java`` (``LST4:N``).

Conventions: images are ``numpy`` arrays indexed ``[y, x]``; float32 arithmetic follows
Java ``float`` semantics (every operation rounded to float32, no FMA); ``double`` is
Python ``float`` / ``numpy.float64``.
"""
from __future__ import annotations

import ctypes
import math
import os
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

f32 = np.float32

# --------------------------------------------------------------------------------------
# parameters of the path (LST4:82-89 defaults, LST4:2370-2425 dialog)
# --------------------------------------------------------------------------------------
RANGE_DISPLAY = 0   # float->8-bit range = display range carried through convertToGray32:
                    #   u8 crop -> [0,255]; u16 / f32 crop -> the crop's own pre-blur min/max
RANGE_CROP = 1      # always the crop's own pre-blur min/max (FloatProcessor.findMinAndMax)
RANGE_FIXED = 2     # caller-supplied [lo, hi]
RANGE_FRAME = 3     # the whole slice's pre-blur min/max: what a crop inherits if createProcessor hands the parent's
                    # min/max down (SURVEY 8c, H1 alternative; slice_proc.getMin()/getMax(), LST4:1308-1312)
QUANT_ROUND255 = 0  # scale = 255/(max-min), ivalue = (int)(v*scale + 0.5f)  (recent ImageJ)
QUANT_TRUNC256 = 1  # scale = 256/(max-min), ivalue = (int)(v*scale)          (old ImageJ)

MATCH_THRESHOLD = (0.1, 0.1, 0.05, 0.05, 0.2, 0.2)  # LST4:147
TM_CCOEFF_NORMED = 5

DBL_EPSILON = 2.220446049250313e-16
FLT_EPSILON = 1.1920928955078125e-07


@dataclass
class Params:
    """The seven dialog parameters (LST4:2401-2407) plus the oracle's unknown-convention knobs."""
    width: int
    height: int
    templ_size: int = 120          # templSize  LST4:84
    s_area: int = 50               # sArea      LST4:83
    mark_dist: float = 100.0       # markDist   LST4:85
    method: int = TM_CCOEFF_NORMED # 0..5 (LST4:2381); 5 is the plugin's default and the north_star path
    sub_pixel: bool = True         # LST4:89
    match_intensity: bool = True   # LST4:88
    range_mode: int = RANGE_DISPLAY
    range_lo: float = 0.0
    range_hi: float = 255.0
    quant_mode: int = QUANT_ROUND255
    doubling: bool = False         # search-area doubling recovery (LST4:1666-1735, 2058-2128); SURVEY 8f-N1
    auto_skip: bool = True         # headless policy: reject => skip the frame (LST4:2146-2148)
    max_s_area: int = 500          # maxSArea LST4:131
    rgb_weights: Tuple[float, float, float] = (1.0 / 3.0, 1.0 / 3.0, 1.0 / 3.0)   # ColorProcessor weighting factors


# --------------------------------------------------------------------------------------
# ImageJ restatements (3P-recall)
# --------------------------------------------------------------------------------------
def make_gaussian_kernel(sigma: float = 2.0, accuracy: float = 0.02, max_radius: int = 50):
    """ij/plugin/filter/GaussianBlur.java makeGaussianKernel [3P-recall].

    Returns (kern, kern_sum) float32 arrays of length kRadius; kern[0] is the centre tap,
    kern_sum[i] = sum of taps i+1..inf (used for edge-pixel replication).
    For sigma=2, accuracy=0.02 (LST4:526,1469): kRadius = ceil(2*sqrt(-2 ln 0.02))+1 = 7 -> 13 taps.
    """
    k_radius = int(math.ceil(sigma * math.sqrt(-2.0 * math.log(accuracy)))) + 1
    if max_radius < 50:
        max_radius = 50
    if k_radius > max_radius:
        k_radius = max_radius
    kern = np.zeros(k_radius, dtype=f32)
    ksum = np.zeros(k_radius, dtype=f32)
    for i in range(k_radius):
        kern[i] = f32(math.exp(-0.5 * i * i / sigma / sigma))
    if k_radius < max_radius and k_radius > 3:  # polynomial edge correction
        sqrt_slope = float("inf")
        r = k_radius
        while r > k_radius // 2:
            r -= 1
            a = math.sqrt(float(kern[r])) / (k_radius - r)
            if a < sqrt_slope:
                sqrt_slope = a
            else:
                break
        for r1 in range(r + 2, k_radius):
            kern[r1] = f32((k_radius - r1) * (k_radius - r1) * sqrt_slope * sqrt_slope)
    if k_radius < max_radius:
        s = float(kern[0])
        for i in range(1, k_radius):
            s += 2.0 * float(kern[i])
    else:
        s = sigma * math.sqrt(2.0 * math.pi)
    rsum = 0.5 + 0.5 * float(kern[0]) / s
    for i in range(k_radius):
        v = float(kern[i]) / s
        kern[i] = f32(v)
        rsum -= v
        ksum[i] = f32(rsum)
    return kern, ksum


def convolve_line_scalar(line: np.ndarray, kern: np.ndarray, ksum: np.ndarray) -> np.ndarray:
    """GaussianBlur.convolveLine [3P-recall], scalar, whole line read and written
    (readFrom=0, readTo=length, writeFrom=0, writeTo=length -- the processor has no ROI).
    Pure-Python; used only to pin ``blur_lines`` on small cases."""
    n = len(line)
    kr = len(kern)
    out = np.zeros(n, dtype=f32)
    first, last = f32(line[0]), f32(line[n - 1])
    first_part = kr if kr < n else n
    i = 0
    while i < first_part:
        res = f32(line[i] * kern[0])
        res = f32(res + f32(ksum[i] * first))
        if i + kr > n:
            res = f32(res + f32(ksum[n - i - 1] * last))
        for k in range(1, kr):
            v = f32(0)
            if i - k >= 0:
                v = f32(v + line[i - k])
            if i + k < n:
                v = f32(v + line[i + k])
            res = f32(res + f32(kern[k] * v))
        out[i] = res
        i += 1
    i_end_inside = n - kr if n - kr < n else n
    while i < i_end_inside:
        res = f32(line[i] * kern[0])
        for k in range(1, kr):
            res = f32(res + f32(kern[k] * f32(line[i - k] + line[i + k])))
        out[i] = res
        i += 1
    while i < n:
        res = f32(line[i] * kern[0])
        if i < kr:
            res = f32(res + f32(ksum[i] * first))
        if i + kr >= n:
            res = f32(res + f32(ksum[n - i - 1] * last))
        for k in range(1, kr):
            v = f32(0)
            if i - k >= 0:
                v = f32(v + line[i - k])
            if i + k < n:
                v = f32(v + line[i + k])
            res = f32(res + f32(kern[k] * v))
        out[i] = res
        i += 1
    return out


def blur_lines(a: np.ndarray, kern: np.ndarray, ksum: np.ndarray) -> np.ndarray:
    """convolveLine applied to every row of ``a`` (float32 [n_lines, length]); vectorised
    but with the scalar routine's exact float32 operation order."""
    a = np.ascontiguousarray(a, dtype=f32)
    n = a.shape[1]
    kr = len(kern)
    if n < 2 * kr:  # short lines: the three loops of convolveLine overlap; use the scalar form
        return np.stack([convolve_line_scalar(r, kern, ksum) for r in a])
    first = a[:, :1]
    last = a[:, n - 1:]
    res = a * kern[0]                                   # input[i]*kern0
    res[:, :kr] = res[:, :kr] + ksum[None, :kr] * first  # += kernSum[i]*first         (i < kRadius)
    # += kernSum[length-i-1]*last  for i + kRadius >= length (third loop), i.e. i >= n-kr
    idx = np.arange(n - kr, n)
    res[:, n - kr:] = res[:, n - kr:] + ksum[None, n - idx - 1] * last
    for k in range(1, kr):
        v = np.zeros_like(a)
        v[:, k:] = v[:, k:] + a[:, : n - k]             # v += input[i-k]  (i-k >= 0)
        v[:, : n - k] = v[:, : n - k] + a[:, k:]        # v += input[i+k]  (i+k < length)
        res = res + kern[k] * v
    return res.astype(f32, copy=False)


_KERN, _KSUM = make_gaussian_kernel(2.0, 0.02, 50)


def blur_gaussian_float(img: np.ndarray) -> np.ndarray:
    """GaussianBlur.blurGaussian(ip, 2, 2, 0.02) on a whole float processor [3P-recall]:
    blurFloat = X pass (blur1Direction xDirection=true) then Y pass, in place, float32."""
    x = blur_lines(img.astype(f32), _KERN, _KSUM)
    y = blur_lines(np.ascontiguousarray(x.T), _KERN, _KSUM)
    return np.ascontiguousarray(y.T)


def clip_rect(x0: int, y0: int, w: int, h: int, width: int, height: int):
    """ImageProcessor.setRoi clips the rectangle to the image [3P-recall]."""
    x1, y1 = min(x0 + w, width), min(y0 + h, height)
    x0, y0 = max(x0, 0), max(y0, 0)
    return x0, y0, max(x1 - x0, 0), max(y1 - y0, 0)


def quantize_u8(blurred: np.ndarray, lo: float, hi: float, quant_mode: int) -> np.ndarray:
    """FloatProcessor.create8BitImage [3P-recall]: float32 arithmetic,
    value = px - min; clamp <0; ivalue = (int)(value*scale + 0.5f); clamp >255.
    Java (int) of NaN is 0 and of +inf saturates (then clamps to 255)."""
    min2, max2 = f32(lo), f32(hi)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        if quant_mode == QUANT_ROUND255:
            scale = f32(255.0) / f32(max2 - min2)
        else:
            scale = f32(256.0) / f32(max2 - min2)
        value = (blurred.astype(f32) - min2).astype(f32)
        value = np.where(value < 0, f32(0), value).astype(f32)
        t = (value * scale).astype(f32)
        if quant_mode == QUANT_ROUND255:
            t = (t + f32(0.5)).astype(f32)
    out = np.zeros(t.shape, dtype=np.uint8)
    finite = np.isfinite(t)
    tv = np.where(finite, t, 0)
    iv = np.trunc(tv).astype(np.int64)
    iv = np.clip(iv, 0, 255)
    out[...] = iv
    out[np.isposinf(t)] = 255
    out[np.isnan(t)] = 0
    return out


def rgb_to_gray8(packed: np.ndarray, weights=(1.0 / 3.0, 1.0 / 3.0, 1.0 / 3.0)) -> np.ndarray:
    """TypeConverter.convertRGBToByte [ImageJ, 3P-recall], the first half of convertToGray32 on a
    ColorProcessor (LST4:519-521): (byte)(r*rw + g*gw + b*bw + 0.5) in double, pixels 0x00RRGGBB."""
    c = packed.astype(np.uint32)
    r = ((c >> 16) & 0xff).astype(np.float64)
    g = ((c >> 8) & 0xff).astype(np.float64)
    b = (c & 0xff).astype(np.float64)
    v = r * weights[0] + g * weights[1] + b * weights[2] + 0.5
    return (v.astype(np.int64) & 0xff).astype(np.uint8)


def preprocess_crop(frame: np.ndarray, x0: int, y0: int, w: int, h: int, p: Params) -> np.ndarray:
    """crop -> [convertToGray32] -> blurGaussian(2,2,0.02) -> 8-bit, i.e. what reaches
    cv::matchTemplate for one window or template (LST4:1346-1347, 1456-1473, 2525-2535).

    u8 + matchIntensity=false: ByteProcessor blur rounds back with (int)(v+0.5f) clamped to
    [0,255] (ByteProcessor.setPixels(int, FloatProcessor) [3P-recall]) == ROUND255 on [0,255].
    u16 + matchIntensity=false dereferences a null Mat in the reference (LST4:2510-2523):
    unsupported here as there."""
    x0, y0, w, h = clip_rect(x0, y0, w, h, frame.shape[1], frame.shape[0])
    crop = frame[y0:y0 + h, x0:x0 + w]
    whole = frame
    if frame.dtype == np.uint32:
        # RGB slice (DOES_RGB, LST4:172): matched on intensity; an 8-bit grey image from here on, whose
        # 0..255 display range survives convertToGray32 like an 8-bit slice's
        if not p.match_intensity:
            # matchIntensity=false: no convertToGray32 (LST4:1456-1468); GaussianBlur.blurGaussian on a ColorProcessor blurs
            # the channels one by one as float images and writes each back with (int)(v+0.5f) clamped to 0..255
            # (ColorProcessor.setPixels(channel, fp)) [ImageJ, 3P-recall]; getBufferedImage -> Java2DFrameUtils.toMat gives
            # OpenCV a CV_8UC3 image (LST4:2525-2535 case 24).  Returned as (h, w, 3), channel order R, G, B (matchTemplate
            # sums over the channels, so the order of the planes does not matter).
            planes = [quantize_u8(blur_gaussian_float(((crop >> np.uint32(sh)) & np.uint32(0xff)).astype(f32)), 0.0, 255.0,
                                  QUANT_ROUND255) for sh in (16, 8, 0)]
            return np.stack(planes, axis=-1)
        crop = rgb_to_gray8(crop, p.rgb_weights)
        if p.range_mode == RANGE_FRAME:
            whole = rgb_to_gray8(frame, p.rgb_weights)
        frame = crop
    if frame.dtype == np.uint8 and not p.match_intensity:
        lo, hi, qm = 0.0, 255.0, QUANT_ROUND255
    else:
        if frame.dtype == np.uint16 and not p.match_intensity:
            raise ValueError("16-bit stacks need matchIntensity=true (reference: null Mat, LST4:2510-2523)")
        qm = p.quant_mode
        if p.range_mode == RANGE_FIXED:
            lo, hi = p.range_lo, p.range_hi
        elif p.range_mode == RANGE_FRAME:
            lo, hi = float(whole.min()), float(whole.max())
        elif p.range_mode == RANGE_DISPLAY and frame.dtype == np.uint8:
            lo, hi = 0.0, 255.0
        else:
            lo, hi = float(crop.min()), float(crop.max())
    blurred = blur_gaussian_float(crop.astype(f32))
    return quantize_u8(blurred, lo, hi, qm)


# --------------------------------------------------------------------------------------
# OpenCV restatements (3P-recall, pinned against cv2 in tests)
# --------------------------------------------------------------------------------------
_HERE = os.path.dirname(os.path.abspath(__file__))
_clib = None


def _load_clib():
    """oracle/_build/liblst_oracle_c.so -- the C restatement of the exact integer
    cross-correlation (oracle/lst_oracle_c.c); built by oracle/Makefile."""
    global _clib
    if _clib is None:
        path = os.path.join(_HERE, "_build", "liblst_oracle_c.so")
        if os.path.exists(path):
            lib = ctypes.CDLL(path)
            lib.lst_oracle_xcorr_u8.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
            lib.lst_oracle_xcorr_u8.restype = None
            _clib = lib
        else:
            _clib = False
    return _clib


def xcorr_exact(win8: np.ndarray, tpl8: np.ndarray) -> np.ndarray:
    """Exact integer cross-correlation sum_{v,u} W[y+v,x+u]*T[v,u] (int64 [Rh,Rw]).
    This is what cv::crossCorr approximates with a float32 DFT (templmatch.cpp)."""
    win8 = np.ascontiguousarray(win8, dtype=np.uint8)
    tpl8 = np.ascontiguousarray(tpl8, dtype=np.uint8)
    sh, sw = win8.shape
    th, tw = tpl8.shape
    rh, rw = sh - th + 1, sw - tw + 1
    lib = _load_clib()
    if lib:
        out = np.zeros((rh, rw), dtype=np.int64)
        lib.lst_oracle_xcorr_u8(win8.ctypes.data, sw, sh, sw, tpl8.ctypes.data, tw, th, out.ctypes.data)
        return out
    # numpy fallback: float64 FFT correlation is accurate to << 0.5 for these magnitudes
    from numpy.fft import rfft2, irfft2
    fh, fw = sh, sw
    a = rfft2(win8.astype(np.float64), (fh, fw))
    b = rfft2(tpl8[::-1, ::-1].astype(np.float64), (fh, fw))
    full = irfft2(a * b, (fh, fw))
    return np.rint(full[th - 1:th - 1 + rh, tw - 1:tw - 1 + rw]).astype(np.int64)


def template_stats(tpl8: np.ndarray):
    """cv::meanStdDev(templ) and the derived constants of common_matchTemplate
    (templmatch.cpp) for a 1-channel template, all in double:
    returns (tsum, tsq, inv_area, templ_mean, templ_norm, flat) where
    templ_norm = sqrt(sdv^2)/sqrt(inv_area) and flat = (sdv^2 < DBL_EPSILON)."""
    t = tpl8.astype(np.int64)
    n = t.size
    tsum = int(t.sum())
    tsq = int((t * t).sum())
    inv_area = 1.0 / float(n)
    mean = float(tsum) / float(n)
    var = float(tsq) / float(n) - mean * mean
    if var < 0.0:
        var = 0.0
    sdv = math.sqrt(var)
    templ_norm = sdv * sdv
    flat = templ_norm < DBL_EPSILON
    templ_norm = math.sqrt(templ_norm)
    templ_norm = templ_norm / math.sqrt(inv_area)
    return tsum, tsq, inv_area, mean, templ_norm, flat


def box_sums(win8: np.ndarray, th: int, tw: int):
    """Window sums of I and I^2 under the template footprint (cv::integral CV_64F is exact
    for u8): int64 [Rh, Rw]."""
    w = win8.astype(np.int64)
    sat = np.zeros((w.shape[0] + 1, w.shape[1] + 1), dtype=np.int64)
    sat[1:, 1:] = w.cumsum(0).cumsum(1)
    sq = np.zeros_like(sat)
    sq[1:, 1:] = (w * w).cumsum(0).cumsum(1)

    def box(s):
        return s[th:, tw:] - s[:-th, tw:] - s[th:, :-tw] + s[:-th, :-tw]
    return box(sat), box(sq)


def ncc_from_sums(corr, wsum, wsq, inv_area: float, templ_mean: float, templ_norm: float, flat: bool):
    """common_matchTemplate, method TM_CCOEFF_NORMED, cn=1 (templmatch.cpp) with the
    *exact* integer correlation instead of the float32 DFT result.  Double arithmetic in
    OpenCV's operation order, stored as float32."""
    corr = np.asarray(corr, dtype=np.float64)
    if flat:
        return np.ones(corr.shape, dtype=f32)
    t = np.asarray(wsum, dtype=np.float64)
    wnd_mean2 = (t * t) * inv_area
    num = corr - t * templ_mean
    wnd_sum2 = np.asarray(wsq, dtype=np.float64)
    diff2 = np.maximum(wnd_sum2 - wnd_mean2, 0.0)
    thr = np.minimum(0.5, 10.0 * FLT_EPSILON * wnd_sum2)
    tt = np.where(diff2 <= thr, 0.0, np.sqrt(diff2) * templ_norm)
    anum = np.abs(num)
    with np.errstate(divide="ignore", invalid="ignore"):
        q = num / tt
    out = np.where(anum < tt, q, np.where(anum < tt * 1.125, np.where(num > 0, 1.0, -1.0), 0.0))
    return out.astype(f32)


def match_from_sums(method: int, corr, wsum, wsq, tpl8: np.ndarray):
    """common_matchTemplate (templmatch.cpp), cn=1, for any of the six methods of LST4:2381
    (0 SQDIFF, 1 SQDIFF_NORMED, 2 CCORR, 3 CCORR_NORMED, 4 CCOEFF, 5 CCOEFF_NORMED) with the
    *exact* integer correlation.  Double arithmetic in OpenCV's order, stored as float32."""
    tsum, tsq, inv_area, mean, norm, flat = template_stats(tpl8)
    if method == 5:
        return ncc_from_sums(corr, wsum, wsq, inv_area, mean, norm, flat)
    corr = np.asarray(corr, dtype=np.float64)
    if method == 2:
        return corr.astype(f32)
    num_type = 0 if method in (2, 3) else (1 if method == 4 else 2)
    normed = method in (1, 3)
    # templSum2 / templNorm as derived from meanStdDev (not from the integer sums)
    var = float(tsq) / float(tpl8.size) - mean * mean
    sdv = math.sqrt(max(var, 0.0))
    sum2 = sdv * sdv + mean * mean
    norm_raw = math.sqrt(sum2) / math.sqrt(inv_area)
    sum2 = sum2 / inv_area
    num = corr
    wnd_mean2 = 0.0
    wnd_sum2 = 0.0
    if num_type == 1:
        t = np.asarray(wsum, dtype=np.float64)
        wnd_mean2 = (t * t) * inv_area
        num = num - t * mean
    if normed or num_type == 2:
        wnd_sum2 = np.asarray(wsq, dtype=np.float64)
        if num_type == 2:
            num = np.maximum(wnd_sum2 - 2.0 * num + sum2, 0.0)
    if normed:
        diff2 = np.maximum(wnd_sum2 - wnd_mean2, 0.0)
        thr = np.minimum(0.5, 10.0 * FLT_EPSILON * wnd_sum2)
        tt = np.where(diff2 <= thr, 0.0, np.sqrt(diff2) * norm_raw)
        anum = np.abs(num)
        with np.errstate(divide="ignore", invalid="ignore"):
            q = num / tt
        other = 1.0 if method == 1 else 0.0
        num = np.where(anum < tt, q, np.where(anum < tt * 1.125, np.where(num > 0, 1.0, -1.0), other))
    return np.asarray(num).astype(f32)


def match_from_sums3(method: int, corr, wsum3, wsq, tpl8: np.ndarray):
    """common_matchTemplate (templmatch.cpp) with cn = 3 (RGB stacks with matchIntensity=false, LST4:2525-2535 case 24):
    corr = sum over the channels of the exact correlations, wsum3[k] = window sums of channel k, wsq = sum over the channels
    of the window sums of squares; tpl8 is (th, tw, 3).  templNorm = sdv0^2 + sdv1^2 + sdv2^2, templSum2 = templNorm +
    mean0^2 + mean1^2 + mean2^2, wndMean2 and num accumulate channel by channel -- OpenCV's operation order, in double."""
    corr = np.asarray(corr, dtype=np.float64)
    if method == 2:
        return corr.astype(f32)
    n = tpl8.shape[0] * tpl8.shape[1]
    inv_area = 1.0 / float(n)
    means, norm = [], 0.0
    for k in range(3):
        t = tpl8[:, :, k].astype(np.int64)
        mean = float(int(t.sum())) / float(n)
        var = float(int((t * t).sum())) / float(n) - mean * mean
        sdv = math.sqrt(max(var, 0.0))
        norm += sdv * sdv
        means.append(mean)
    if method == 5 and norm < DBL_EPSILON:
        return np.ones(corr.shape, dtype=f32)
    sum2 = norm
    for k in range(3):
        sum2 += means[k] * means[k]
    num_type = 0 if method in (2, 3) else (1 if method in (4, 5) else 2)
    normed = method in (1, 3, 5)
    if num_type != 1:
        means = [0.0, 0.0, 0.0]
        norm = sum2
    sum2 = sum2 / inv_area
    norm = math.sqrt(norm) / math.sqrt(inv_area)
    num = corr
    wnd_mean2 = 0.0
    wnd_sum2 = 0.0
    if num_type == 1:
        wnd_mean2 = np.zeros(corr.shape, dtype=np.float64)
        for k in range(3):
            t = np.asarray(wsum3[k], dtype=np.float64)
            wnd_mean2 = wnd_mean2 + t * t
            num = num - t * means[k]
        wnd_mean2 = wnd_mean2 * inv_area
    if normed or num_type == 2:
        wnd_sum2 = np.asarray(wsq, dtype=np.float64)
        if num_type == 2:
            num = np.maximum(wnd_sum2 - 2.0 * num + sum2, 0.0)
    if normed:
        diff2 = np.maximum(wnd_sum2 - wnd_mean2, 0.0)
        thr = np.minimum(0.5, 10.0 * FLT_EPSILON * wnd_sum2)
        tt = np.where(diff2 <= thr, 0.0, np.sqrt(diff2) * norm)
        anum = np.abs(num)
        with np.errstate(divide="ignore", invalid="ignore"):
            q = num / tt
        other = 1.0 if method == 1 else 0.0
        num = np.where(anum < tt, q, np.where(anum < tt * 1.125, np.where(num > 0, 1.0, -1.0), other))
    return np.asarray(num).astype(f32)


def ncc_map_exact(win8: np.ndarray, tpl8: np.ndarray, method: int = TM_CCOEFF_NORMED) -> np.ndarray:
    """matchTemplate(win, tpl, method) with an exact numerator (LST4:2560)."""
    if tpl8.ndim == 3:
        th, tw = tpl8.shape[:2]
        corr, wsq, wsum3 = 0, 0, []
        for k in range(3):
            wk = np.ascontiguousarray(win8[:, :, k])
            corr = corr + xcorr_exact(wk, np.ascontiguousarray(tpl8[:, :, k]))
            s1, s2 = box_sums(wk, th, tw)
            wsum3.append(s1)
            wsq = wsq + s2
        return match_from_sums3(method, corr, wsum3, wsq, tpl8)
    th, tw = tpl8.shape
    corr = xcorr_exact(win8, tpl8)
    wsum, wsq = box_sums(win8, th, tw)
    return match_from_sums(method, corr, wsum, wsq, tpl8)


def ncc_map_cv2(win8: np.ndarray, tpl8: np.ndarray, method: int = TM_CCOEFF_NORMED) -> np.ndarray:
    """The real third-party routine (OpenCV in this image, IPP off, 1 thread)."""
    import cv2
    cv2.setNumThreads(1)
    try:
        cv2.ipp.setUseIPP(False)
    except Exception:
        pass
    return cv2.matchTemplate(np.ascontiguousarray(win8), np.ascontiguousarray(tpl8), method)   # cv2.TM_* == 0..5


def best_loc_first(res: np.ndarray, method: int = TM_CCOEFF_NORMED):
    """minMaxLoc + the min/max choice of LST4:2671-2679: minimum for methods 0 and 1."""
    if method in (0, 1):
        idx = int(np.argmin(res))
        y, x = divmod(idx, res.shape[1])
        return x, y, float(res[y, x])
    return max_loc_first(res)


def max_loc_first(res: np.ndarray):
    """cv::minMaxLoc max: first occurrence in row-major order (SURVEY 2.1 [MEASURED])."""
    idx = int(np.argmax(res))
    y, x = divmod(idx, res.shape[1])
    return x, y, float(res[y, x])


# --------------------------------------------------------------------------------------
# the reference's own arithmetic
# --------------------------------------------------------------------------------------
def subpixel_fit(res: np.ndarray, x: int, y: int) -> Tuple[float, float]:
    """3x3 quadratic fit, LST4:2681-2717.  FloatIndexer.get returns a Java float, so the four-term sum of fxy and the
    differences of fx / fy are float32 operations (one rounding each, left to right) promoted to double only by the
    division by 4.0 / 2.0; fxx and fyy are promoted from the start by the 2.0 * term (LST4:2694-2699)."""
    rows, cols = res.shape
    if x == 0 or x == cols - 1 or y == 0 or y == rows - 1:
        return 0.0, 0.0
    f = lambda yy, xx: np.float32(res[yy, xx])
    r = lambda yy, xx: float(res[yy, xx])
    fxx = r(y, x - 1) - 2.0 * r(y, x) + r(y, x + 1)
    fyy = r(y - 1, x) - 2.0 * r(y, x) + r(y + 1, x)
    with np.errstate(all="ignore"):
        fxy = float(np.float32(np.float32(np.float32(f(y + 1, x + 1) + f(y - 1, x - 1)) - f(y - 1, x + 1)) - f(y + 1, x - 1))) / 4.0
        fx = float(np.float32(f(y, x + 1) - f(y, x - 1))) / 2.0
        fy = float(np.float32(f(y + 1, x) - f(y - 1, x))) / 2.0
    denom = fxy * fxy - fxx * fyy
    if denom == 0.0:
        return 0.0, 0.0
    dx = (fyy * fx - fxy * fy) / denom
    dy = (fxx * fy - fxy * fx) / denom
    if abs(dx) > 1.0 or abs(dy) > 1.0:
        return 0.0, 0.0
    return dx, dy


def do_match_coord_res(win8: np.ndarray, tpl8: np.ndarray, sub_pix: bool, backend: str = "exact",
                       method: int = TM_CCOEFF_NORMED):
    """doMatch_coord_res(src, tpl, 5, subPix, null) on the 8-bit images that reach OpenCV
    (LST4:2502-2563, 2665-2722).  Returns ([x, y, score], result_map)."""
    res = ncc_map_exact(win8, tpl8, method) if backend == "exact" else ncc_map_cv2(win8, tpl8, method)
    x, y, v = best_loc_first(res, method)
    cx, cy = float(x), float(y)
    if sub_pix:
        dx, dy = subpixel_fit(res, x, y)
        cx += dx
        cy += dy
    return [cx, cy, v], res


def do_match_test(tpl8: np.ndarray, backend: str = "exact", method: int = TM_CCOEFF_NORMED) -> float:
    """doMatch_test: matchTemplate(tpl, tpl) -> 1x1 result (LST4:2724-2767)."""
    res = ncc_map_exact(tpl8, tpl8, method) if backend == "exact" else ncc_map_cv2(tpl8, tpl8, method)
    return float(res[0, 0])


def test_match_result(result: float, ref: float, x: float, y: float, search_width: int, tpl_size: int,
                      thresh: Optional[float] = None, method: int = TM_CCOEFF_NORMED) -> bool:
    """testMatchResult (LST4:1225-1255)."""
    ok = True
    dist = min(0.05 * tpl_size, 0.05 * search_width)
    if thresh is None:
        thresh = MATCH_THRESHOLD[method]
    with np.errstate(divide="ignore", invalid="ignore"):
        result, ref = np.float64(result), np.float64(ref)    # Java double division: x/0 = inf, 0/0 = NaN
        if method == 0:
            bad = result / ref > thresh
        elif method == 1:
            bad = result > thresh
        elif method == 2:
            bad = abs((result - ref) / ref) > thresh
        elif method == 4:
            bad = abs(result - ref) / ref > thresh
        else:
            bad = abs(result - ref) > thresh
    if bad:
        ok = False
    if search_width != 0 and (x < dist or y < dist or x > search_width - dist or y > search_width - dist):
        ok = False
    return ok


def jint(v: float) -> int:
    """Java (int) cast of a double: truncation toward zero (LST4:1327)."""
    return int(math.trunc(v))


def calc_displacement(ref_xy: Sequence[Tuple[float, float]], dis: Sequence[Tuple[float, float]],
                      mark_dist: float, spot0: Optional[Tuple[float, float]]):
    """calcDisplacement (LST4:2318-2365).  ref_xy / dis are ordered spot, mark1..mark4;
    ref_xy are the Java ``float`` click coordinates (LST4:95-99).
    Returns (dX_pix, dY_pix, x_abs, y_abs, dL, spot0)."""
    (rsx, rsy), (r1x, r1y), (r2x, r2y), (r3x, r3y), (r4x, r4y) = [(float(f32(a)), float(f32(b))) for a, b in ref_xy]
    (dsx, dsy), (d1x, d1y), (d2x, d2y), (d3x, d3y), (d4x, d4y) = dis
    dX_pix = dsx - d1x
    dY_pix = dsy - d1y
    m1x, m1y = r1x + d1x, r1y + d1y
    m2x, m2y = r2x + d2x, r2y + d2y
    m3x, m3y = r3x + d3x, r3y + d3y
    m4x, m4y = r4x + d4x, r4y + d4y
    a1x, a1y = m4x - m1x, m4y - m1y
    b1x, b1y = m2x - m1x, m2y - m1y
    a2x, a2y = m3x - m2x, m3y - m2y
    b2x, b2y = m3x - m4x, m3y - m4y
    rx, ry = rsx + dsx - m1x, rsy + dsy - m1y
    AX = a1x * (b1y - b2y) + a1y * (b2x - b1x)
    BX = -a1x * b1y + a1y * b1x - rx * (b1y - b2y) - ry * (b2x - b1x)
    CX = -rx * b1y + ry * b1x
    AY = b1x * (a1y - a2y) + b1y * (a2x - a1x)
    BY = -b1x * a1y + b1y * a1x - rx * (a1y - a2y) - ry * (a2x - a1x)
    CY = -rx * a1y + ry * a1x

    def jdiv(a, b):  # Java double division: x/0 -> +-inf or NaN, never raises
        if b == 0.0:
            if a == 0.0 or a != a:
                return float("nan")
            neg = (a < 0) != (math.copysign(1.0, b) < 0)
            return float("-inf") if neg else float("inf")
        return a / b

    def jsqrt(a):
        return math.sqrt(a) if a >= 0 else float("nan")

    if AX == 0.0:
        x_abs = jdiv(CX, BX) * mark_dist
        y_abs = jdiv(CY, BY) * mark_dist
    else:
        DX = jsqrt(BX * BX + 4.0 * AX * CX)
        DY = jsqrt(BY * BY + 4.0 * AY * CY)
        x_abs = jdiv(jdiv(DX - BX, AX), 2.0) * mark_dist
        y_abs = jdiv(jdiv(-DY - BY, AY), 2.0) * mark_dist
    if spot0 is None:
        spot0 = (x_abs, y_abs)
    dX = x_abs - spot0[0]
    dY = y_abs - spot0[1]
    s = dX * dX + dY * dY
    dL = math.sqrt(s) if s >= 0 else float("nan")
    return dX_pix, dY_pix, x_abs, y_abs, dL, spot0


# --------------------------------------------------------------------------------------
# the tracking loop (LST4:486-702 setup, LST4:793-902 loop, LST4:1302-2216 analyseSlice)
# --------------------------------------------------------------------------------------
STATUS_OK = 0
STATUS_SKIP = 1
STATUS_STOP = 2
ORDER = ("spot", "mark1", "mark2", "mark3", "mark4")  # storage order everywhere
MATCH_ORDER = (1, 2, 3, 4, 0)                          # analyseSlice matches mark1..4, then the spot


@dataclass
class TemplateMatch:
    x: float = 0.0          # coord_res[0]: result-map x (+ sub-pixel)
    y: float = 0.0
    score: float = 0.0      # coord_res[2]: double of a float32
    ix: int = 0             # integer peak
    iy: int = 0
    wx: int = 0             # window actually used (xStart, yStart, sWX, sWY)
    wy: int = 0
    ww: int = 0
    wh: int = 0
    s_used: int = 0         # search area used for the accepted/last match (sArea or a doubled one)
    accepted: bool = False
    evaluated: bool = False


@dataclass
class FrameRecord:
    status: int = STATUS_OK
    reject: int = -1         # index (storage order) of the template that rejected the frame, -1 = none
    tm: List[TemplateMatch] = field(default_factory=lambda: [TemplateMatch() for _ in range(5)])
    dX_pix: float = 0.0
    dY_pix: float = 0.0
    x_abs: float = 0.0
    y_abs: float = 0.0
    dL: float = 0.0
    dis: List[Tuple[float, float]] = field(default_factory=lambda: [(0.0, 0.0)] * 5)


class OracleTracker:
    """Headless restatement of the plugin's tracking session for method 5.

    Policy where the reference opens a dialog (LST4:1281-1300): the frame is skipped
    (answer 'Skip the frame' / autoSkip, LST4:2146-2148) and the caller restores the
    displacement state (LST4:832-846).
    """

    def __init__(self, params: Params, backend: str = "exact"):
        self.p = params
        self.backend = backend
        self.rects: List[Tuple[int, int, int, int]] = []
        self.ref_xy: List[Tuple[float, float]] = []
        self.tpl8: List[np.ndarray] = []
        self.mideal: List[float] = []
        self.dis: List[Tuple[float, float]] = [(0.0, 0.0)] * 5
        self.spot0 = None
        self.ref_record: Optional[FrameRecord] = None

    # -- setup: LST4:486-702 ----------------------------------------------------------
    def set_reference(self, frame: np.ndarray, ref_xy: Sequence[Tuple[float, float]]) -> FrameRecord:
        p = self.p
        assert frame.shape == (p.height, p.width)
        half = float(f32(p.templ_size) / f32(2.0))       # float rect_half_size = templSize / 2.0f  (LST4:506)
        self.ref_xy = [(float(f32(x)), float(f32(y))) for x, y in ref_xy]
        self.rects, self.tpl8, self.mideal = [], [], []
        for (rx, ry) in self.ref_xy:
            # new Roi((int)(ref - half), (int)(ref - half), (int)(2*half), (int)(2*half))  LST4:512 (float math)
            x0 = jint(float(f32(f32(rx) - f32(half))))
            y0 = jint(float(f32(f32(ry) - f32(half))))
            side = jint(float(f32(f32(2.0) * f32(half))))
            self.rects.append((x0, y0, side, side))                               # Roi.getBounds(): unclipped
            cx0, cy0, cw, ch = clip_rect(x0, y0, side, side, p.width, p.height)   # crop() clips to the image
            t8 = preprocess_crop(frame, cx0, cy0, cw, ch, p)
            self.tpl8.append(t8)
            self.mideal.append(do_match_test(t8, self.backend, 2 if p.method == 0 else p.method))   # LST4:566
        self.dis = [(0.0, 0.0)] * 5
        self.spot0 = None
        rec = FrameRecord()
        out = calc_displacement(self.ref_xy, self.dis, p.mark_dist, None)          # LST4:702
        rec.dX_pix, rec.dY_pix, rec.x_abs, rec.y_abs, rec.dL, self.spot0 = out
        rec.dis = list(self.dis)
        self.ref_record = rec
        return rec

    # -- window placement: LST4:1325-1445 ---------------------------------------------
    def window_for(self, k: int, idis: Tuple[int, int], s: Optional[int] = None, shrink: bool = False):
        p = self.p
        s = p.s_area if s is None else s
        if p.s_area == 0:
            return 0, 0, p.width, p.height, (False, False, False, False)
        rx, ry, rw, rh = self.rects[k]
        xs = rx + idis[0] - s
        ys = ry + idis[1] - s
        sw = rw + 2 * s
        sh = rh + 2 * s
        lb = ub = rb = bb = False
        if xs < 0:
            xs = 0
            lb = True
        if ys < 0:
            ys = 0
            ub = True
        if xs + sw > p.width:
            if shrink:
                sw = p.width - xs
            else:
                xs = p.width - sw
            rb = True
        if ys + sh > p.height:
            if shrink:
                sh = p.height - ys
            else:
                ys = p.height - sh
            bb = True
        return xs, ys, sw, sh, (lb, rb, bb, ub)

    def _match(self, frame, k, wx, wy, ww, wh, s_used) -> TemplateMatch:
        p = self.p
        cx, cy, cw, ch = clip_rect(wx, wy, ww, wh, p.width, p.height)
        win8 = preprocess_crop(frame, cx, cy, cw, ch, p)
        coord, res = do_match_coord_res(win8, self.tpl8[k], p.sub_pixel, self.backend, p.method)
        ix, iy, _ = best_loc_first(res, p.method)
        tm = TemplateMatch(x=coord[0], y=coord[1], score=coord[2], ix=ix, iy=iy,
                           wx=wx, wy=wy, ww=ww, wh=wh, s_used=s_used, evaluated=True)
        tm.accepted = test_match_result(coord[2], self.mideal[k], coord[0], coord[1], 2 * s_used, p.templ_size,
                                        method=p.method)
        return tm

    # -- analyseSlice: LST4:1302-2216 --------------------------------------------------
    def analyse_slice(self, frame: np.ndarray) -> FrameRecord:
        """One frame.  On STATUS_SKIP the displacement state is left untouched (the
        reference mutates and the caller restores, LST4:820-846 -- same net effect)."""
        p = self.p
        rec = FrameRecord()
        dis = list(self.dis)
        idis = [(jint(dx), jint(dy)) for dx, dy in dis]
        wins = [list(self.window_for(k, idis[k])[:4]) for k in range(5)]
        for k in MATCH_ORDER:
            wx, wy, ww, wh = wins[k]
            tm = self._match(frame, k, wx, wy, ww, wh, p.s_area)
            if not tm.accepted and p.doubling and p.s_area != 0 and k in (1, 0):
                tm = self._doubling(frame, k, idis[k], tm)
                if tm.accepted and k == 1:
                    # mark1 found by doubling: shift the other four windows (LST4:1757-1880)
                    x_shift = tm.x + tm.wx - self.rects[1][0] - dis[1][0]
                    y_shift = tm.y + tm.wy - self.rects[1][1] - dis[1][1]
                    for j in (0, 2, 3, 4):
                        xs = jint(wins[j][0] + x_shift)      # int += double  => (int)(int + double)
                        ys = jint(wins[j][1] + y_shift)
                        sw, sh = wins[j][2], wins[j][3]
                        if xs < 0:
                            xs = 0
                        if ys < 0:
                            ys = 0
                        if xs + sw > p.width:
                            xs = p.width - sw
                        if ys + sh > p.height:
                            ys = p.height - sh
                        wins[j][0], wins[j][1] = xs, ys
            rec.tm[k] = tm
            if not tm.accepted:
                rec.status = STATUS_SKIP
                rec.reject = k
                rec.dis = list(self.dis)
                return rec
            dis[k] = (tm.x + tm.wx - self.rects[k][0], tm.y + tm.wy - self.rects[k][1])   # LST4:1891-1892
        self.dis = dis
        out = calc_displacement(self.ref_xy, self.dis, p.mark_dist, self.spot0)
        rec.dX_pix, rec.dY_pix, rec.x_abs, rec.y_abs, rec.dL, self.spot0 = out
        rec.dis = list(self.dis)
        return rec

    def _doubling(self, frame, k, idis, tm_first) -> TemplateMatch:
        """Search-area doubling (mark1 LST4:1666-1735, spot LST4:2058-2128)."""
        p = self.p
        s_new = p.s_area
        found = False
        lb = rb = bb = ub = False
        tm = tm_first
        while not found and not (lb and rb and bb and ub):
            s_new *= 2
            if k == 0 and p.auto_skip and p.max_s_area != 0 and s_new > p.max_s_area:   # LST4:2066
                break
            xs, ys, sw, sh, (l, r, b, u) = self.window_for(k, idis, s_new, shrink=True)
            lb, rb, bb, ub = lb or l, rb or r, bb or b, ub or u
            tm = self._match(frame, k, xs, ys, sw, sh, s_new)
            found = tm.accepted
        return tm

    def run(self, frames) -> List[FrameRecord]:
        """The driver loop LST4:793-846 over an iterable of frames (after the reference frame)."""
        return [self.analyse_slice(f) for f in frames]
