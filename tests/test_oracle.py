"""The CPU oracle against known answers, against the real OpenCV in this image, and against the
frozen fixtures of tests/golden/ (see make_golden.py for what they pin and what they cannot)."""
import os

import numpy as np
import pytest

from laser_spot_track_b200.synth import StackConfig, SyntheticStack
from oracle import lst_oracle as O

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# ---- ImageJ restatements ----------------------------------------------------------------------
def test_gaussian_kernel_values():
    kern, ksum = O.make_gaussian_kernel(2.0, 0.02, 50)
    # 13 taps, values computed from the recalled algorithm
    want = [0.19968557, 0.17622191, 0.12111542, 0.06482842, 0.02702450, 0.00877357, 0.00219339]
    assert len(kern) == 7
    assert np.allclose(kern, want, atol=5e-9)
    assert abs(float(kern[0]) + 2 * float(kern[1:].astype(np.float64).sum()) - 1.0) < 1e-7
    # kern_sum[i] = tail beyond tap i
    tail = [0.5 - 0.5 * float(kern[0]) - float(kern[1:i + 1].astype(np.float64).sum()) for i in range(7)]
    assert np.allclose(ksum, tail, atol=1e-7)
    g = np.load(os.path.join(G, "blur_kat.npz"))
    assert np.array_equal(kern, g["kern"]) and np.array_equal(ksum, g["ksum"])


def test_blur_vectorised_equals_scalar_convolve_line():
    rng = np.random.default_rng(0)
    kern, ksum = O.make_gaussian_kernel()
    for n in (14, 15, 20, 33, 96):
        a = rng.uniform(0, 65535, size=(5, n)).astype(np.float32)
        got = O.blur_lines(a, kern, ksum)
        want = np.stack([O.convolve_line_scalar(r, kern, ksum) for r in a])
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), n
    # short lines (n < 2*kRadius) take the scalar form by construction
    a = rng.uniform(0, 255, size=(3, 9)).astype(np.float32)
    assert O.blur_lines(a, kern, ksum).shape == (3, 9)


def test_blur_properties():
    # a constant image stays constant to float rounding; edge replication keeps the mean
    img = np.full((30, 40), 123.0, np.float32)
    out = O.blur_gaussian_float(img)
    assert np.abs(out - 123.0).max() < 1e-4
    # impulse response = outer product of the kernel (interior)
    imp = np.zeros((41, 41), np.float32)
    imp[20, 20] = 1.0
    out = O.blur_gaussian_float(imp)
    kern, _ = O.make_gaussian_kernel()
    full = np.concatenate([kern[:0:-1], kern])
    assert np.allclose(out[14:27, 14:27], np.outer(full, full), atol=1e-9)
    assert abs(out.sum() - 1.0) < 1e-5


def test_quantise_conventions():
    v = np.array([[0.0, 0.49, 0.5, 254.5, 255.0, 300.0, -3.0]], np.float32)
    assert O.quantize_u8(v, 0, 255, O.QUANT_ROUND255).tolist() == [[0, 0, 1, 255, 255, 255, 0]]
    assert O.quantize_u8(v, 0, 255, O.QUANT_TRUNC256).tolist() == [[0, 0, 0, 255, 255, 255, 0]]
    # degenerate range: Java (int)NaN == 0, (int)+inf saturates then clamps to 255
    flat = np.array([[5.0, 5.0000005, 4.0]], np.float32)
    assert O.quantize_u8(flat, 5, 5, O.QUANT_ROUND255).tolist() == [[0, 255, 0]]
    g = np.load(os.path.join(G, "blur_kat.npz"))
    assert np.array_equal(O.preprocess_crop(g["img8"], 0, 0, 52, 40, O.Params(52, 40, 14, 4, match_intensity=False)), g["q8"])
    assert np.array_equal(O.preprocess_crop(g["img16"], 0, 0, 47, 33, O.Params(47, 33, 14, 4)), g["q16"])
    assert np.array_equal(O.blur_gaussian_float(g["img16"].astype(np.float32)).view(np.uint32), g["blur16"].view(np.uint32))
    with pytest.raises(ValueError):   # 16-bit + matchIntensity=false: null Mat in the reference
        O.preprocess_crop(g["img16"], 0, 0, 47, 33, O.Params(47, 33, 14, 4, match_intensity=False))


# ---- OpenCV restatements ------------------------------------------------------------------------
def test_xcorr_c_equals_numpy(oracle_clib):
    rng = np.random.default_rng(1)
    win = rng.integers(0, 256, size=(40, 37), dtype=np.uint8)
    tpl = rng.integers(0, 256, size=(11, 9), dtype=np.uint8)
    got = O.xcorr_exact(win, tpl)
    sw = np.lib.stride_tricks.sliding_window_view(win.astype(np.int64), tpl.shape)
    want = np.einsum("yxvu,vu->yx", sw, tpl.astype(np.int64))
    assert np.array_equal(got, want)


def test_ncc_restatement_pinned_by_opencv(oracle_clib):
    """common_matchTemplate restated vs the real cv::matchTemplate: same map within OpenCV's
    float32-DFT noise (SURVEY [MEASURED] <= 6e-5), same integer peak, sub-pixel within 1e-3 px."""
    g = np.load(os.path.join(G, "match_kat.npz"))
    for i in range(4):
        win, tpl = g[f"win{i}"], g[f"tpl{i}"]
        exact = O.ncc_map_exact(win, tpl)
        cvm = O.ncc_map_cv2(win, tpl)
        assert np.array_equal(exact.view(np.uint32), g[f"exactmap{i}"].view(np.uint32))
        assert np.abs(cvm - g[f"cv2map{i}"]).max() <= 1e-6          # this image's OpenCV is what was frozen
        assert np.abs(exact - cvm).max() <= 1e-4
        x, y, v = O.max_loc_first(exact)
        assert (x, y) == (int(g[f"cv2peak{i}"][0]), int(g[f"cv2peak{i}"][1])) == (13, 9)
        assert abs(v - g[f"cv2peak{i}"][2]) <= 1e-4
        a, _ = O.do_match_coord_res(win, tpl, True, "exact")
        b, _ = O.do_match_coord_res(win, tpl, True, "cv2")
        assert abs(a[0] - b[0]) <= 1e-3 and abs(a[1] - b[1]) <= 1e-3


def test_ncc_known_answers(oracle_clib):
    rng = np.random.default_rng(2)
    tpl = rng.integers(0, 256, size=(16, 16), dtype=np.uint8)
    # flat window -> 0 everywhere; flat template -> 1 everywhere (OpenCV semantics)
    for backend in (O.ncc_map_exact, O.ncc_map_cv2):
        assert np.all(backend(np.full((40, 40), 50, np.uint8), tpl) == 0)
        assert np.all(backend(rng.integers(0, 256, size=(40, 40), dtype=np.uint8), np.full((16, 16), 7, np.uint8)) == 1)
    # template cut from the window: peak at the cut offset, score 1 (exact) / ~1 (cv2)
    win = rng.integers(0, 256, size=(60, 70), dtype=np.uint8)
    t2 = np.ascontiguousarray(win[21:37, 33:49])
    x, y, v = O.max_loc_first(O.ncc_map_exact(win, t2))
    assert (x, y) == (33, 21) and abs(v - 1.0) <= 1.2e-7
    assert O.do_match_test(t2, "exact") in (1.0, float(np.float32(1.0) - np.float32(2 ** -24)))
    # two equal maxima -> the first in row-major order (cv::minMaxLoc)
    w2 = rng.integers(0, 30, size=(50, 80), dtype=np.uint8)
    w2[10:26, 40:56] = tpl
    w2[10:26, 5:21] = tpl
    import cv2
    m = O.ncc_map_exact(w2, tpl)
    assert O.max_loc_first(m)[:2] == (5, 10)
    assert cv2.minMaxLoc(m)[3] == (5, 10)


def test_subpixel_fit_kats():
    # sampled paraboloid: exact recovery of the vertex
    yy, xx = np.mgrid[0:9, 0:9].astype(np.float64)
    for (cx, cy) in [(4.3, 3.8), (4.0, 4.0), (3.6, 4.45)]:
        res = (1.0 - 0.01 * (xx - cx) ** 2 - 0.02 * (yy - cy) ** 2 + 0.005 * (xx - cx) * (yy - cy)).astype(np.float32)
        x, y, _ = O.max_loc_first(res)
        dx, dy = O.subpixel_fit(res, x, y)
        assert abs(x + dx - cx) < 2e-5 and abs(y + dy - cy) < 2e-5
    res = np.zeros((5, 5), np.float32)
    res[0, 2] = 1.0
    assert O.subpixel_fit(res, 2, 0) == (0.0, 0.0)              # peak on the border (LST4:2686-2691)
    assert O.subpixel_fit(np.ones((5, 5), np.float32), 2, 2) == (0.0, 0.0)   # denom == 0 (LST4:2701-2703)
    res = np.zeros((5, 5), np.float32)                            # |dx| > 1 -> rejected (LST4:2707-2710)
    res[2, 2], res[2, 3], res[2, 1] = 1.0, 0.999, 0.0
    res[1, 2] = res[3, 2] = 0.5
    dx, dy = O.subpixel_fit(res, 2, 2)
    assert abs(dx) <= 1.0 and abs(dy) <= 1.0


def test_match_result_thresholds():
    # |score - mideal| <= 0.2 and peak not within min(0.05*T, 0.05*2s) of the map edge (LST4:1225-1255)
    assert O.test_match_result(0.85, 1.0, 30.0, 30.0, 64, 32)
    assert not O.test_match_result(0.79, 1.0, 30.0, 30.0, 64, 32)
    assert not O.test_match_result(0.95, 1.0, 1.5, 30.0, 64, 32)       # d = min(1.6, 3.2) = 1.6
    assert O.test_match_result(0.95, 1.0, 1.7, 30.0, 64, 32)
    assert not O.test_match_result(0.95, 1.0, 30.0, 62.5, 64, 32)
    assert O.test_match_result(0.95, 1.0, 0.0, 0.0, 0, 32)             # sArea == 0: no edge test


# ---- the loop, against the frozen fixtures --------------------------------------------------------
GOLDEN_TRACKS = ["track_a_u8.npz", "track_small_u16.npz"] + [f"track_method{m}_u8.npz" for m in range(5)]


@pytest.mark.parametrize("name", GOLDEN_TRACKS)
def test_tracking_loop_golden(oracle_clib, name):
    g = np.load(os.path.join(G, name))
    w, h, T, s, n = [int(v) for v in g["cfg"]]
    cfg = StackConfig(w, h, str(g["dtype"]), T, s, 24 if name.startswith("track_small") else 200)
    st = SyntheticStack(cfg)
    frames = [st.frame(i) for i in range(n + 1)]
    assert [int(f.astype(np.int64).sum()) for f in frames] == g["frame_crc"].tolist(), "synthetic generator changed"
    op = O.Params(w, h, T, s, match_intensity=bool(g["match_intensity"]), method=int(g["method"]))
    tr = O.OracleTracker(op)
    ref = tr.set_reference(frames[0], st.ref_xy)
    assert np.array_equal(np.array(tr.rects), g["rects"]) and tr.mideal == g["mideal"].tolist()
    assert [ref.dX_pix, ref.dY_pix, ref.x_abs, ref.y_abs, ref.dL] == g["ref_out"].tolist()
    recs = tr.run(frames[1:])
    for i, r in enumerate(recs):
        assert r.status == g["status"][i]
        assert [(t.ix, t.iy) for t in r.tm] == [tuple(p) for p in g["peaks"][i]]
        assert [t.score for t in r.tm] == g["scores"][i].tolist()
        assert [(t.x, t.y) for t in r.tm] == [tuple(p) for p in g["pos"][i]]
        assert [r.dX_pix, r.dY_pix, r.x_abs, r.y_abs, r.dL] == g["out"][i].tolist()
    if int(g["method"]) != 5:
        return
    # sanity against ground truth: tracked spot displacement follows the synthetic motion
    t = st.truth(n) - st.truth(0)
    assert abs(recs[-1].dX_pix - (t[0, 0] - t[1, 0])) < 0.15 and abs(recs[-1].dY_pix - (t[0, 1] - t[1, 1])) < 0.15


@pytest.mark.parametrize("method", [0, 1, 2, 3, 4, 5])
def test_all_methods_pinned_against_cv2(oracle_clib, method):
    """common_matchTemplate restated for the six methods of the dialog (LST4:2381),
    pinned against the OpenCV in this image.  Tolerance relative to the map's scale: cv2's numerator
    is a float32 DFT."""
    import cv2
    rng = np.random.default_rng(40 + method)
    for shape, tshape, blur in [((90, 100), (32, 32), 0), ((90, 100), (32, 32), 3), ((64, 70), (20, 33), 0)]:
        win = rng.integers(0, 256, shape, dtype=np.uint8)
        if blur:
            win = cv2.GaussianBlur(win, (0, 0), blur)
        tpl = win[20:20 + tshape[0], 30:30 + tshape[1]].copy() if blur == 0 and tshape == (32, 32) else \
            rng.integers(0, 256, tshape, dtype=np.uint8)
        a = O.ncc_map_exact(win, tpl, method)
        b = O.ncc_map_cv2(win, tpl, method)
        scale = max(1.0, float(np.abs(b).max()))
        assert float(np.abs(a.astype(np.float64) - b).max()) / scale < 2e-4
        assert O.best_loc_first(a, method)[:2] == O.best_loc_first(b, method)[:2]


def test_match_result_all_methods():
    """testMatchResult's per-method acceptance (LST4:1225-1255) with matchThreshold of LST4:147."""
    T = O.test_match_result
    assert T(0.09e6, 1.0e6, 30, 30, 64, 32, method=0) and not T(0.11e6, 1.0e6, 30, 30, 64, 32, method=0)
    assert T(0.09, 123.0, 30, 30, 64, 32, method=1) and not T(0.11, 0.0, 30, 30, 64, 32, method=1)
    assert T(0.96e6, 1.0e6, 30, 30, 64, 32, method=2) and not T(1.06e6, 1.0e6, 30, 30, 64, 32, method=2)
    assert T(0.96, 1.0, 30, 30, 64, 32, method=3) and not T(0.94, 1.0, 30, 30, 64, 32, method=3)
    assert T(0.81e6, 1.0e6, 30, 30, 64, 32, method=4) and not T(0.79e6, 1.0e6, 30, 30, 64, 32, method=4)
    assert not T(0.96e6, 1.0e6, 1.0, 30, 64, 32, method=2)           # edge test is method-independent
    # Java double division by a zero reference: x/0 -> inf (> thresh) for x > 0, 0/0 -> NaN (not > thresh)
    assert not T(5.0, 0.0, 30, 30, 64, 32, method=0) and T(0.0, 0.0, 30, 30, 64, 32, method=0)


def test_rgb_to_gray8_is_imagej_unweighted_mean():
    """convertRGBToByte with the default weights (1/3 each) and with the 'weighted' ones."""
    px = np.array([[0x000000, 0xffffff, 0xff0000, 0x010101], [0x102030, 0x0000ff, 0xfefdfc, 0x808182]], np.uint32)
    g = O.rgb_to_gray8(px)
    assert g.tolist() == [[0, 255, 85, 1], [32, 85, 253, 129]]
    gw = O.rgb_to_gray8(px, (0.299, 0.587, 0.114))
    assert gw.tolist() == [[0, 255, 76, 1], [int(16 * 0.299 + 32 * 0.587 + 48 * 0.114 + 0.5), 29, 253, 129]]


@pytest.mark.parametrize("method", [0, 1, 2, 3, 4, 5])
def test_methods_golden(oracle_clib, method):
    """Frozen vectors of SURVEY 8f-N3: cv2's own result map, its min/max location and the 1x1 self match
    for every method, on fixed 8-bit inputs (tests/golden/match_methods_kat.npz)."""
    g = np.load(os.path.join(G, "match_methods_kat.npz"))
    win, tpl = g["win"], g["tpl"]
    exact = O.ncc_map_exact(win, tpl, method)
    assert np.array_equal(exact.view(np.uint32), g[f"exactmap{method}"].view(np.uint32))
    cvm = g[f"cv2map{method}"]
    scale = max(1.0, float(np.abs(cvm).max()))
    assert float(np.abs(exact.astype(np.float64) - cvm).max()) / scale < 1e-4
    x, y, v = O.best_loc_first(exact, method)
    assert (x, y) == (int(g[f"cv2best{method}"][0]), int(g[f"cv2best{method}"][1]))
    assert abs(v - g[f"cv2best{method}"][2]) / scale < 1e-4
    assert O.do_match_test(tpl, "exact", method) == g[f"self{method}"][0]


def test_three_channel_match_restatement_vs_cv2():
    """common_matchTemplate with cn = 3 (RGB stacks, matchIntensity=false, LST4:2525-2535 case 24) restated with an exact
    numerator, pinned against the real cv2.matchTemplate on CV_8UC3 images: maps within OpenCV's float32-DFT noise
    (relative to the map's scale for the unnormalised methods), same extremum."""
    import cv2
    rng = np.random.default_rng(5)
    for trial in range(3):
        win = cv2.GaussianBlur(rng.integers(0, 256, (70, 90, 3), dtype=np.uint8), (0, 0), 2)
        tpl = win[20:52, 30:62].copy()
        for method in range(6):
            a = O.ncc_map_exact(win, tpl, method)
            b = O.ncc_map_cv2(win, tpl, method)
            scale = max(1.0, float(np.abs(b).max())) if method in (0, 2, 4) else 1.0
            assert float(np.abs(a - b).max()) / scale <= 1e-4, (trial, method)
            assert O.best_loc_first(a, method)[:2] == O.best_loc_first(b, method)[:2]
            if method in (0, 1, 3, 5):   # the template was cut from the window at (30, 20)
                assert O.best_loc_first(a, method)[:2] == (30, 20)
    flat = np.full((16, 16, 3), 9, np.uint8)
    assert np.all(O.ncc_map_exact(win, flat, 5) == 1.0) and np.all(O.ncc_map_cv2(win, flat, 5) == 1.0)


def test_rgb_three_channel_preprocess_planes():
    """matchIntensity=false on a ColorProcessor: every channel is blurred like an 8-bit image of its own (ByteProcessor rule:
    (int)(v+0.5f) clamped), so the planes equal the 8-bit path applied to each channel."""
    rng = np.random.default_rng(2)
    rgb = rng.integers(0, 1 << 24, (60, 80), dtype=np.uint32)
    p = O.Params(80, 60, 16, 8)
    p.match_intensity = False
    got = O.preprocess_crop(rgb, 5, 7, 50, 40, p)
    assert got.shape == (40, 50, 3)
    for k, sh in enumerate((16, 8, 0)):
        ch = ((rgb >> sh) & 0xff).astype(np.uint8)
        assert np.array_equal(got[:, :, k], O.preprocess_crop(ch, 5, 7, 50, 40, p))
