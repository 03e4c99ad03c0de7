"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle on
the same seeded inputs.  Integer/byte work must be bit-exact; float32 scores are compared
bit-exactly against the exact oracle and within 1e-4 against OpenCV (cv2) whose DFT numerator
carries ~5e-5 noise; sub-pixel positions within 1e-3 px against cv2."""
import numpy as np
import pytest

from laser_spot_track_b200 import LaserSpotTrack4, native as N
from laser_spot_track_b200.synth import CONFIGS, StackConfig, SyntheticStack
from oracle import lst_oracle as O

import util

pytestmark = pytest.mark.gpu

SCORE_TOL_CV2 = 1e-4      # north_star: correlation scores within 1e-4
SUBPIX_TOL_CV2 = 1e-3     # north_star: sub-pixel positions / displacements within 1e-3 px


def _tracker(cfg, op, **kw):
    bits = {"u8": 8, "u16": 16, "f32": 32}[cfg.dtype]
    return LaserSpotTrack4(cfg.width, cfg.height, bits, method=op.method, templSize=op.templ_size, sArea=op.s_area,
                           markDist=op.mark_dist, subPixel=op.sub_pixel, matchIntensity=op.match_intensity,
                           doubling=op.doubling, maxSArea=op.max_s_area, range_mode=op.range_mode,
                           range_lo=op.range_lo, range_hi=op.range_hi, quant_mode=op.quant_mode, **kw)


@pytest.fixture(scope="module")
def stack_a(built_lib, oracle_clib):
    cfg = CONFIGS["A"]
    st = SyntheticStack(cfg)
    return cfg, st, [st.frame(i) for i in range(0, 61)]


def test_device_present(built_lib):
    assert N.load().lst_device_count() >= 1


@pytest.mark.parametrize("dtype,mi,range_mode,quant", [
    ("u8", False, O.RANGE_DISPLAY, O.QUANT_ROUND255),
    ("u8", True, O.RANGE_DISPLAY, O.QUANT_ROUND255),
    ("u8", True, O.RANGE_CROP, O.QUANT_ROUND255),
    ("u8", True, O.RANGE_CROP, O.QUANT_TRUNC256),
    ("u16", True, O.RANGE_DISPLAY, O.QUANT_ROUND255),
    ("u16", True, O.RANGE_FIXED, O.QUANT_ROUND255),
    ("f32", True, O.RANGE_DISPLAY, O.QUANT_ROUND255),
    ("u16", True, O.RANGE_FRAME, O.QUANT_ROUND255),
    ("f32", True, O.RANGE_FRAME, O.QUANT_TRUNC256),
    ("u8", True, O.RANGE_FRAME, O.QUANT_ROUND255),
])
def test_preprocess_bit_exact(built_lib, dtype, mi, range_mode, quant):
    """crop -> float -> blur -> 8-bit: every byte equal to the oracle's, incl. windows touching the borders."""
    cfg = StackConfig(320, 240, dtype, 32, 24, 4)
    st = SyntheticStack(cfg)
    fr = st.frame(1)
    op = util.oracle_params(cfg, match_intensity=mi, range_mode=range_mode, quant_mode=quant,
                            range_lo=100.0, range_hi=60000.0 if dtype == "u16" else 200.0)
    with _tracker(cfg, op) as t:
        for (x0, y0, w, h) in [(0, 0, 80, 80), (240, 160, 80, 80), (37, 41, 96, 64), (100, 50, 160, 160),
                               (0, 0, 320, 240), (300, 0, 20, 33), (5, 200, 150, 14)]:
            got = t.preprocess(fr, x0, y0, w, h)
            want = O.preprocess_crop(fr, x0, y0, w, h, op)
            assert got.shape == want.shape
            bad = np.argwhere(got != want)
            assert bad.size == 0, f"{dtype} rect {(x0, y0, w, h)}: {len(bad)} bytes differ, first {bad[:3]}, got {got[tuple(bad[0])]} want {want[tuple(bad[0])]}"


@pytest.mark.parametrize("S,T", [(96, 32), (160, 48), (64, 14), (75, 33), (200, 120), (150, 127)])
def test_match_single_map(built_lib, oracle_clib, S, T):
    """matchTemplate + minMaxLoc + sub-pixel on 8-bit inputs: full score map."""
    rng = np.random.default_rng(S * 1000 + T)
    base = rng.integers(0, 256, size=(S + 40, S + 40), dtype=np.uint8)
    import cv2
    base = cv2.GaussianBlur(base, (0, 0), 2.0)
    win = np.ascontiguousarray(base[:S, :S])
    tpl = np.ascontiguousarray(base[11:11 + T, 17:17 + T]) if S - T > 17 else np.ascontiguousarray(base[3:3 + T, 5:5 + T])
    with LaserSpotTrack4(640, 480, 8) as t:
        (res, m) = t.doMatch_coord_res(win, tpl, 5, True, want_map=True)
        exact = O.ncc_map_exact(win, tpl)
        assert m.shape == exact.shape
        assert np.array_equal(m.view(np.uint32), exact.view(np.uint32)), f"score map differs from the exact oracle: max abs {np.abs(m - exact).max()}"
        want, _ = O.do_match_coord_res(win, tpl, True, "exact")
        assert res == want, (res, want)
        cvm = O.ncc_map_cv2(win, tpl)
        assert np.abs(m - cvm).max() <= SCORE_TOL_CV2
        wcv, _ = O.do_match_coord_res(win, tpl, True, "cv2")
        assert abs(res[2] - wcv[2]) <= SCORE_TOL_CV2
        assert abs(res[0] - wcv[0]) <= SUBPIX_TOL_CV2 and abs(res[1] - wcv[1]) <= SUBPIX_TOL_CV2
        # doMatch_test: 1x1 self match
        assert t.doMatch_test(tpl) == O.do_match_test(tpl, "exact")


def test_match_edge_cases(built_lib, oracle_clib):
    with LaserSpotTrack4(640, 480, 8) as t:
        rng = np.random.default_rng(5)
        tpl = rng.integers(0, 256, size=(20, 20), dtype=np.uint8)
        # flat window -> every score 0 -> first position wins
        res, m = t.doMatch_coord_res(np.full((50, 60), 77, np.uint8), tpl, 5, True, want_map=True)
        assert np.all(m == 0) and res == [0.0, 0.0, 0.0]
        # flat template -> every score 1
        res, m = t.doMatch_coord_res(rng.integers(0, 256, size=(50, 60), dtype=np.uint8), np.full((20, 20), 9, np.uint8), 5, True, want_map=True)
        assert np.all(m == 1) and res == [0.0, 0.0, 1.0]
        # two exact copies of the template -> first in row-major order
        win = rng.integers(0, 40, size=(70, 90), dtype=np.uint8)
        win[30:50, 50:70] = tpl
        win[30:50, 10:30] = tpl
        res = t.doMatch_coord_res(win, tpl, 5, False)
        want, _ = O.do_match_coord_res(win, tpl, False, "exact")
        assert res == want and (res[0], res[1]) == (10.0, 30.0)
        # window == template size -> 1x1 map, peak on the border -> no sub-pixel shift
        res = t.doMatch_coord_res(tpl, tpl, 5, True)
        assert res[0] == 0.0 and res[1] == 0.0 and res[2] == O.do_match_test(tpl)


@pytest.mark.parametrize("mi,max_batch", [(False, 0), (True, 7), (False, 1)])
def test_track_config_a(stack_a, mi, max_batch):
    """Config A (640x480 8-bit, T=32, s=32): the whole per-frame record, bit-exact vs the exact oracle."""
    cfg, st, frames = stack_a
    op = util.oracle_params(cfg, match_intensity=mi)
    otr = O.OracleTracker(op)
    oref = otr.set_reference(frames[0], st.ref_xy)
    want = otr.run(frames[1:])
    with _tracker(cfg, op, max_batch=max_batch) as t:
        info = t.set_reference(frames[0], st.ref_xy)
        assert [tuple(r) for r in info.rect] == [tuple(r) for r in otr.rects]
        assert list(info.mideal) == otr.mideal
        assert (info.ref_record.x_abs, info.ref_record.y_abs, info.ref_record.dL) == (oref.x_abs, oref.y_abs, oref.dL)
        got = t.track(np.stack(frames[1:]))
        util.compare_records(got, want, label=f"A mi={mi}")
        assert np.array_equal(t.get_state(), np.array([v for d in otr.dis for v in d]))
        s = t.stats()
        assert s.frames == len(want) and s.kernel_launches > 0


def test_track_config_a_vs_cv2(stack_a):
    """Same stack against the OpenCV-backed oracle (the reference's real third-party routine)."""
    cfg, st, frames = stack_a
    op = util.oracle_params(cfg, match_intensity=False)
    otr = O.OracleTracker(op, backend="cv2")
    otr.set_reference(frames[0], st.ref_xy)
    want = otr.run(frames[1:])
    with _tracker(cfg, op) as t:
        t.set_reference(frames[0], st.ref_xy)
        got = t.track(np.stack(frames[1:]))
    near_ties = 0
    for i, (g, w) in enumerate(zip(got, want)):
        assert g.status == w.status
        for k in range(5):
            a, b = g.tm[k], w.tm[k]
            if (a.ix, a.iy) != (b.ix, b.iy):     # only allowed for a documented near-tie
                near_ties += 1
                assert abs(a.score - b.score) <= SCORE_TOL_CV2, f"frame {i} tpl {k}: peaks differ and scores are not tied"
                continue
            assert abs(a.score - b.score) <= SCORE_TOL_CV2
            assert abs(a.x - b.x) <= SUBPIX_TOL_CV2 and abs(a.y - b.y) <= SUBPIX_TOL_CV2
        if w.status == 0:
            assert abs(g.dX_pix - w.dX_pix) <= 2 * SUBPIX_TOL_CV2 and abs(g.dY_pix - w.dY_pix) <= 2 * SUBPIX_TOL_CV2
    assert near_ties <= 2, f"{near_ties} near-tie peak disagreements with OpenCV"


def test_track_u16_small(built_lib, oracle_clib):
    """16-bit stack, T=48, s=56 (config B's template/search geometry on a small slice)."""
    cfg = StackConfig(480, 400, "u16", 48, 56, 24)
    st = SyntheticStack(cfg)
    frames = [st.frame(i) for i in range(0, 17)]
    op = util.oracle_params(cfg)
    otr = O.OracleTracker(op)
    otr.set_reference(frames[0], st.ref_xy)
    want = otr.run(frames[1:])
    with _tracker(cfg, op) as t:
        t.set_reference(frames[0], st.ref_xy)
        got = t.track_slices(frames[1:])
        util.compare_records(got, want, label="u16")


@pytest.mark.parametrize("dtype", ["u16", "f32"])
@pytest.mark.parametrize("supplied", [False, True])
def test_track_range_frame(built_lib, oracle_clib, dtype, supplied):
    """LST_RANGE_FRAME (the crop inherits the whole slice's min/max): records bit-exact against the oracle,
    with the ranges reduced on the device and with the ranges the caller announces (slice_proc.getMin()/getMax()) --
    the second with pinned slices, so that only the search regions cross PCIe."""
    cfg = StackConfig(480, 400, dtype, 48, 56, 24)
    st = SyntheticStack(cfg)
    frames = [st.frame(i) for i in range(0, 13)]
    op = util.oracle_params(cfg, range_mode=O.RANGE_FRAME)
    otr = O.OracleTracker(op)
    otr.set_reference(frames[0], st.ref_xy)
    want = otr.run(frames[1:])
    with _tracker(cfg, op, max_batch=5) as t:
        if supplied:
            t.set_frame_ranges([(float(frames[0].min()), float(frames[0].max()))])
        t.set_reference(frames[0], st.ref_xy)
        if supplied:
            view, ptr = _pinned_stack(frames[1:])
            try:
                t.set_frame_ranges([(float(f.min()), float(f.max())) for f in frames[1:]])
                got = t.track(view)
                util.compare_records(got, want, label=f"range FRAME supplied {dtype}")
            finally:
                N.load().lst_free_pinned(ptr)
        else:
            got = t.track_slices(frames[1:])
            util.compare_records(got, want, label=f"range FRAME reduced {dtype}")


def test_rejects_and_skips(stack_a):
    """A frame whose spot is blanked is rejected and skipped; state is restored (LST4:832-846)."""
    cfg, st, frames = stack_a
    frames = [f.copy() for f in frames[:12]]
    frames[5][:] = frames[5].mean().astype(np.uint8)      # nothing to match in slice 5
    c = st.truth(8)[0].astype(int)
    frames[8][c[1] - 40:c[1] + 40, c[0] - 40:c[0] + 40] = 30   # spot erased in slice 8
    op = util.oracle_params(cfg, match_intensity=False)
    otr = O.OracleTracker(op)
    otr.set_reference(frames[0], st.ref_xy)
    want = otr.run(frames[1:])
    assert [w.status for w in want].count(1) >= 2
    with _tracker(cfg, op) as t:
        t.set_reference(frames[0], st.ref_xy)
        got = t.track(np.stack(frames[1:]))
        util.compare_records(got, want, label="skips")


def test_unsupported_like_reference(built_lib):
    from laser_spot_track_b200 import LstError
    with pytest.raises(LstError):
        LaserSpotTrack4(640, 480, 16, matchIntensity=False)     # null Mat in the reference, LST4:2510-2523
    with pytest.raises(LstError):
        LaserSpotTrack4(640, 480, 8, method=6)
    with pytest.raises(LstError):
        LaserSpotTrack4(640, 480, 12)
    with LaserSpotTrack4(640, 480, 8) as t:
        with pytest.raises(LstError):
            t.track(np.zeros((1, 480, 640), np.uint8))            # no reference yet


def _pinned_stack(frames):
    """Copy slices into one lst_alloc_pinned block; returns (numpy view, raw pointer)."""
    import ctypes as C
    lib = N.load()
    a = np.stack(frames)
    ptr = C.c_void_p()
    assert lib.lst_alloc_pinned(a.nbytes, C.byref(ptr)) == 0
    view = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(a.nbytes,)).view(a.dtype).reshape(a.shape)
    view[...] = a
    return view, ptr


@pytest.mark.parametrize("max_batch,roi_mode", [(4, 0), (16, 0), (7, 2)])
def test_roi_ingest_pinned_equals_whole_slice(stack_a, max_batch, roi_mode):
    """Pinned host slices take the ROI ingest (only the search regions cross PCIe); results must be
    identical to the oracle and to the whole-slice ingest, including when the windows outrun the
    speculation margin and are read straight from host memory."""
    cfg, st, frames = stack_a
    op = util.oracle_params(cfg, match_intensity=True, range_mode=O.RANGE_CROP)
    otr = O.OracleTracker(op)
    otr.set_reference(frames[0], st.ref_xy)
    want = otr.run(frames[1:41])
    view, ptr = _pinned_stack(frames[1:41])
    try:
        with _tracker(cfg, op, max_batch=max_batch, ingest_mode=roi_mode) as t:   # 0: 3-D DMA, 2: gather kernel
            t.set_reference(frames[0], st.ref_xy)
            got = t.track(view)
            util.compare_records(got, want, label="roi")
            if roi_mode == 0:
                assert t.stats().h2d_bytes < 0.6 * view.nbytes, "ROI ingest should move a fraction of the stack"
            # scattered slices (one pointer each, reversed order in memory): runs of length 1
            t.set_state([0.0] * 10)
            got = t.track_host_ptrs([view[i].ctypes.data for i in range(view.shape[0])])
            util.compare_records(got, want, label="roi ptrs")
        with _tracker(cfg, op, max_batch=max_batch, ingest_mode=1) as t:
            t.set_reference(frames[0], st.ref_xy)
            util.compare_records(t.track(view), want, label="whole")
            assert t.stats().h2d_bytes >= view.nbytes
    finally:
        N.load().lst_free_pinned(ptr)


def test_roi_ingest_margin_exceeded(built_lib, oracle_clib):
    """Fast target: the search windows leave the gathered regions -> zero-copy fallback, same results."""
    cfg = StackConfig(640, 480, "u16", 32, 40, 30, spot_travel=2.0)
    st = SyntheticStack(cfg)
    frames = [st.frame(i) for i in range(0, 25)]
    op = util.oracle_params(cfg)
    otr = O.OracleTracker(op)
    otr.set_reference(frames[0], st.ref_xy)
    want = otr.run(frames[1:])
    assert abs(want[-1].dis[0][0]) > 40
    view, ptr = _pinned_stack(frames[1:])
    try:
        with _tracker(cfg, op, max_batch=6) as t:
            t.set_reference(frames[0], st.ref_xy)
            got = t.track(view)
            util.compare_records(got, want, label="fallback")
            assert t.stats().roi_fallback_tasks > 0
    finally:
        N.load().lst_free_pinned(ptr)


@pytest.mark.parametrize("name", ["track_a_u8.npz", "track_small_u16.npz"] + [f"track_method{m}_u8.npz" for m in range(5)])
def test_golden_fixtures_on_gpu(built_lib, name):
    """The CUDA path against the committed fixtures (no oracle at run time)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name))
    w, h, T, s, n = [int(v) for v in g["cfg"]]
    cfg = StackConfig(w, h, str(g["dtype"]), T, s, 24 if name.startswith("track_small") else 200)
    st = SyntheticStack(cfg)
    frames = [st.frame(i) for i in range(n + 1)]
    bits = {"u8": 8, "u16": 16}[cfg.dtype]
    with LaserSpotTrack4(w, h, bits, method=int(g["method"]), templSize=T, sArea=s, matchIntensity=bool(g["match_intensity"])) as t:
        info = t.set_reference(frames[0], st.ref_xy)
        assert np.array_equal(np.array([list(r) for r in info.rect]), g["rects"])
        assert list(info.mideal) == g["mideal"].tolist()
        got = t.track(np.stack(frames[1:]))
        for i, r in enumerate(got):
            assert r.status == g["status"][i]
            assert [(m.ix, m.iy) for m in r.tm] == [tuple(p) for p in g["peaks"][i]]
            assert [m.score for m in r.tm] == g["scores"][i].tolist()
            assert [(m.x, m.y) for m in r.tm] == [tuple(p) for p in g["pos"][i]]
            assert [r.dX_pix, r.dY_pix, r.x_abs, r.y_abs, r.dL] == g["out"][i].tolist()


def test_track_whole_slice_search(built_lib, oracle_clib):
    """sArea == 0: every template is searched over the whole slice (LST4:1316-1320); no edge test."""
    cfg = StackConfig(232, 184, "u16", 24, 0, 8)
    st = SyntheticStack(cfg)
    frames = [st.frame(i) for i in range(0, 5)]
    op = util.oracle_params(cfg)
    otr = O.OracleTracker(op)
    otr.set_reference(frames[0], st.ref_xy)
    want = otr.run(frames[1:])
    assert all(w.tm[0].ww == cfg.width and w.tm[0].wh == cfg.height for w in want)
    with _tracker(cfg, op) as t:
        t.set_reference(frames[0], st.ref_xy)
        got = t.track(np.stack(frames[1:]))
        util.compare_records(got, want, label="sArea=0")


def test_large_template_regime(built_lib, oracle_clib):
    """Config D geometry (128x128 templates, 512x512 ROIs) on a 1024x1024 slice: one frame against the
    exact oracle, peaks/scores/sub-pixel against OpenCV within the north_star tolerances."""
    cfg = StackConfig(1024, 1024, "u8", 128, 192, 4)
    st = SyntheticStack(cfg)
    frames = [st.frame(i) for i in range(0, 3)]
    op = util.oracle_params(cfg, match_intensity=False)
    ocv = O.OracleTracker(op, backend="cv2")
    ocv.set_reference(frames[0], st.ref_xy)
    want_cv = ocv.run(frames[1:])
    with _tracker(cfg, op) as t:
        t.set_reference(frames[0], st.ref_xy)
        got = t.track(np.stack(frames[1:]))
        for g, w in zip(got, want_cv):
            assert g.status == w.status == 0
            for k in range(5):
                assert (g.tm[k].ix, g.tm[k].iy) == (w.tm[k].ix, w.tm[k].iy)
                assert abs(g.tm[k].score - w.tm[k].score) <= SCORE_TOL_CV2
                assert abs(g.tm[k].x - w.tm[k].x) <= SUBPIX_TOL_CV2 and abs(g.tm[k].y - w.tm[k].y) <= SUBPIX_TOL_CV2
        # one window bit-exact against the exact oracle (the C correlation takes a few seconds)
        m = got[0].tm[0]
        win8 = O.preprocess_crop(frames[1], m.wx, m.wy, m.ww, m.wh, op)
        ex, _ = O.do_match_coord_res(win8, ocv.tpl8[0], True, "exact")
        assert [m.x, m.y, m.score] == ex


def test_batch_size_invariance_full_size(built_lib):
    """BASELINE config B at full slice size (1920x1080 16-bit, T=48, s=56): results do not depend on
    how frames are batched (property test; the comparison with the oracle at this size is
    tests/test_gpu_configs.py::test_config_b_full_size_vs_oracle)."""
    cfg = StackConfig(1920, 1080, "u16", 48, 56, 48, motion="periodic", period=48)
    st = SyntheticStack(cfg)
    frames = np.stack([st.frame(i) for i in range(0, 25)])
    outs = []
    for mb in (1, 5, 24):
        with LaserSpotTrack4(1920, 1080, 16, templSize=48, sArea=56, max_batch=mb) as t:
            t.set_reference(frames[0], st.ref_xy)
            recs = t.track(frames[1:])
            outs.append([(r.status, [(m.ix, m.iy, m.wx, m.wy, m.score, m.x, m.y) for m in r.tm],
                          r.dX_pix, r.dY_pix, r.x_abs, r.y_abs, r.dL) for r in recs])
            assert all(r.status == 0 for r in recs)
    assert outs[0] == outs[1] == outs[2]
    # accuracy sanity against the synthetic ground truth
    t24 = st.truth(24) - st.truth(0)
    assert abs(outs[0][-1][2] - (t24[0, 0] - t24[1, 0])) < 0.2


def test_search_area_doubling_on_gpu(built_lib, oracle_clib):
    """SURVEY 8f-N1 on the device: a scene jump larger than sArea is recovered by doubling the search
    area (mark1: LST4:1666-1882 with the shift of the other four windows; spot: LST4:2058-2128).
    Windows of different sizes share one device pass."""
    cfg = StackConfig(400, 320, "u8", 24, 10, 12)
    st = SyntheticStack(cfg)
    frames = [st.frame(i) for i in range(0, 9)]
    for i in range(4, 9):
        frames[i] = np.roll(frames[i], (-14, 17), axis=(0, 1))
    c = st.truth(7)[0].astype(int) + np.array([17, -14])
    patch = frames[7][c[1] - 12:c[1] + 12, c[0] - 12:c[0] + 12].copy()
    frames[7][c[1] - 12:c[1] + 12, c[0] - 12:c[0] + 12] = frames[7][c[1] - 12:c[1] + 12, c[0] + 40:c[0] + 64]
    frames[7][c[1] - 12:c[1] + 12, c[0] + 13:c[0] + 37] = patch
    op = util.oracle_params(cfg, match_intensity=False, doubling=True, max_s_area=0)
    seq = O.OracleTracker(op)
    seq.set_reference(frames[0], st.ref_xy)
    want = seq.run(frames[1:])
    assert any(w.tm[1].s_used > cfg.s_area and w.status == 0 for w in want)
    for mb in (1, 8):
        with _tracker(cfg, op, max_batch=mb) as t:
            t.set_reference(frames[0], st.ref_xy)
            got = t.track(np.stack(frames[1:]))
            util.compare_records(got, want, label=f"doubling mb={mb}")


@pytest.mark.parametrize("n_ctx", [2, 3])
def test_multi_context_sharding(stack_a, n_ctx):
    """lst_multi_*: one context + host thread per device, contiguous chunks, chunk-boundary state
    speculated and verified.  Uses every visible GPU (cycling), so it also runs on one."""
    from laser_spot_track_b200 import LaserSpotTrack4Multi
    cfg, st, frames = stack_a
    op = util.oracle_params(cfg, match_intensity=False)
    otr = O.OracleTracker(op)
    otr.set_reference(frames[0], st.ref_xy)
    want = otr.run(frames[1:])
    ndev = N.load().lst_device_count()
    devices = [i % ndev for i in range(n_ctx)]
    with LaserSpotTrack4Multi(cfg.width, cfg.height, 8, devices, templSize=op.templ_size, sArea=op.s_area,
                              matchIntensity=False, max_batch=16) as t:
        info = t.set_reference(frames[0], st.ref_xy)
        assert list(info.mideal) == otr.mideal
        got = t.track(np.stack(frames[1:31]))
        util.compare_records(got, want[:30], label=f"multi{n_ctx} a")
        got2 = t.track(np.stack(frames[31:]), first_frame_index=31)     # the state carries over between calls
        util.compare_records(got2, want[30:], label=f"multi{n_ctx} b")
        assert np.array_equal(t.get_state(), np.array([v for d in otr.dis for v in d]))


def test_both_match_engines_identical(stack_a):
    """lst_ncc_tc (tcgen05.mma kind::i8, default) and lst_ncc_match (IDP.4A) return identical records."""
    cfg, st, frames = stack_a
    op = util.oracle_params(cfg, match_intensity=False)
    otr = O.OracleTracker(op)
    otr.set_reference(frames[0], st.ref_xy)
    want = otr.run(frames[1:25])
    for engine in (1, 2):
        with _tracker(cfg, op, ncc_engine=engine, max_batch=9) as t:
            t.set_reference(frames[0], st.ref_xy)
            got = t.track(np.stack(frames[1:25]))
            util.compare_records(got, want, label=f"engine {engine}")
    # odd template / window sizes through both engines (match_single uses IDP.4A for the map; tracking uses auto)
    cfg2 = StackConfig(300, 260, "u8", 33, 21, 10)
    st2 = SyntheticStack(cfg2)
    fr2 = [st2.frame(i) for i in range(0, 6)]
    op2 = util.oracle_params(cfg2, match_intensity=False)
    o2 = O.OracleTracker(op2)
    o2.set_reference(fr2[0], st2.ref_xy)
    w2 = o2.run(fr2[1:])
    for engine in (1, 2):
        with _tracker(cfg2, op2, ncc_engine=engine) as t:
            t.set_reference(fr2[0], st2.ref_xy)
            util.compare_records(t.track(np.stack(fr2[1:])), w2, label=f"odd sizes engine {engine}")


@pytest.mark.parametrize("method", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("S,T", [(96, 32), (75, 33), (150, 127)])
def test_match_single_other_methods(built_lib, oracle_clib, method, S, T):
    """the other five methods of the dialog (LST4:2381), score map bit-exact against
    the exact oracle, min/max choice of LST4:2671-2679, and close to cv2 (relative to the map's scale:
    OpenCV's float32 DFT numerator carries ~1e-6 relative noise on the raw-valued methods)."""
    rng = np.random.default_rng(S * 1000 + T)
    import cv2
    base = cv2.GaussianBlur(rng.integers(0, 256, size=(S + 40, S + 40), dtype=np.uint8), (0, 0), 2.0)
    win = np.ascontiguousarray(base[:S, :S])
    tpl = np.ascontiguousarray(base[11:11 + T, 17:17 + T]) if S - T > 17 else np.ascontiguousarray(base[3:3 + T, 5:5 + T])
    with LaserSpotTrack4(640, 480, 8) as t:
        res, m = t.doMatch_coord_res(win, tpl, method, True, want_map=True)
        exact = O.ncc_map_exact(win, tpl, method)
        assert np.array_equal(m.view(np.uint32), exact.view(np.uint32)), f"method {method}: max abs {np.abs(m - exact).max()}"
        want, _ = O.do_match_coord_res(win, tpl, True, "exact", method)
        assert res == want, (res, want)
        cvm = O.ncc_map_cv2(win, tpl, method)
        scale = max(1.0, float(np.abs(cvm).max()))
        assert float(np.abs(m.astype(np.float64) - cvm).max()) / scale <= SCORE_TOL_CV2
        assert t.doMatch_test(tpl, method) == O.do_match_test(tpl, "exact", method)
        # without the map (the tracking kernels' path): same peak
        assert t.doMatch_coord_res(win, tpl, method, True) == want


@pytest.mark.parametrize("method", [0, 1, 2, 3, 4])
def test_track_other_methods(stack_a, method):
    """Whole records for the other five methods, both match engines, against the oracle."""
    cfg, st, frames = stack_a
    op = util.oracle_params(cfg, match_intensity=False, method=method)
    otr = O.OracleTracker(op)
    otr.set_reference(frames[0], st.ref_xy)
    want = otr.run(frames[1:21])
    for engine in (1, 2):
        with _tracker(cfg, op, ncc_engine=engine, max_batch=8) as t:
            info = t.set_reference(frames[0], st.ref_xy)
            assert list(info.mideal) == otr.mideal
            util.compare_records(t.track(np.stack(frames[1:21])), want, label=f"method {method} engine {engine}")


def _rgb_stack(cfg, n):
    """An RGB stack (ColorProcessor pixels 0x00RRGGBB) whose grey value carries the synthetic scene."""
    st = SyntheticStack(cfg)
    frames = []
    for i in range(n):
        g = st.frame(i).astype(np.uint32)
        r, b = (g * 3 // 4 + 17) & 0xff, (255 - g // 2) & 0xff
        frames.append((r << 16) | (g << 8) | b)
    return st, frames


@pytest.mark.parametrize("weights", [None, (0.299, 0.587, 0.114)])
def test_rgb_stack_matched_on_intensity(built_lib, oracle_clib, weights):
    """DOES_RGB (LST4:172) with matchIntensity=true: convertToGray32 of a ColorProcessor,
    then the 8-bit path.  Blurred 8-bit crops and whole records against the oracle."""
    cfg = StackConfig(320, 240, "u8", 32, 20, 14)
    st, frames = _rgb_stack(cfg, 14)
    kw = dict(match_intensity=True)
    if weights:
        kw["rgb_weights"] = weights
    op = util.oracle_params(cfg, **kw)
    otr = O.OracleTracker(op)
    otr.set_reference(frames[0], st.ref_xy)
    want = otr.run(frames[1:])
    assert sum(w.status == 0 for w in want) >= 10
    extra = dict(rgb_weights=weights) if weights else {}
    with LaserSpotTrack4(cfg.width, cfg.height, 24, templSize=cfg.templ_size, sArea=cfg.s_area, matchIntensity=True,
                         max_batch=5, **extra) as t:
        for (x0, y0, w, h) in [(0, 0, 72, 72), (101, 57, 75, 91), (248, 168, 72, 72)]:
            got8 = t.preprocess(frames[3], x0, y0, w, h)
            assert np.array_equal(got8, O.preprocess_crop(frames[3], x0, y0, w, h, op))
        info = t.set_reference(frames[0], st.ref_xy)
        assert list(info.mideal) == otr.mideal
        util.compare_records(t.track(np.stack(frames[1:])), want, label="RGB")
        # (h, w, 3) uint8 arrays are accepted as well
        rgb3 = np.stack([(frames[5] >> 16) & 0xff, (frames[5] >> 8) & 0xff, frames[5] & 0xff], axis=-1).astype(np.uint8)
        assert np.array_equal(t.preprocess(rgb3, 10, 10, 80, 80), O.preprocess_crop(frames[5], 10, 10, 80, 80, op))


def _rgb3_stack(cfg, n):
    """An RGB stack whose channels carry different pictures: the scene, its negative and a shifted copy -- the sums over the
    channels then differ from any single-channel match."""
    st = SyntheticStack(cfg)
    frames = []
    for i in range(n):
        g = st.frame(i).astype(np.uint32)
        r = (255 - g) & 0xff
        b = np.roll(g, (3, -2), axis=(0, 1)) // 2 + (np.arange(cfg.width, dtype=np.uint32)[None, :] * 5) % 64
        frames.append((r << 16) | (g << 8) | (b & 0xff))
    return st, frames


@pytest.mark.parametrize("method", [5, 0, 1, 3, 4])
def test_rgb_stack_three_channel_match(built_lib, oracle_clib, method):
    """SURVEY 8f-N4, LST4:2525-2535 case 24 with matchIntensity=false on an RGB stack: the ColorProcessor is blurred channel by
    channel and OpenCV matches a CV_8UC3 image (correlation and norms summed over the channels).  Blurred planes byte-exact,
    records bit-exact against the exact oracle and within the north_star tolerances against cv2's 3-channel matchTemplate."""
    cfg = StackConfig(320, 240, "u8", 32, 20, 14)
    st, frames = _rgb3_stack(cfg, 12)
    op = util.oracle_params(cfg, match_intensity=False, method=method)
    otr = O.OracleTracker(op)
    otr.set_reference(frames[0], st.ref_xy)
    want = otr.run(frames[1:])
    assert sum(w.status == 0 for w in want) >= 8
    ocv = O.OracleTracker(op, backend="cv2")
    ocv.set_reference(frames[0], st.ref_xy)
    want_cv = ocv.run(frames[1:])
    with LaserSpotTrack4(cfg.width, cfg.height, 24, templSize=cfg.templ_size, sArea=cfg.s_area, matchIntensity=False,
                         method=method, max_batch=5) as t:
        for (x0, y0, w, h) in [(0, 0, 72, 72), (101, 57, 75, 91), (248, 168, 72, 72)]:
            got8 = t.preprocess(frames[3], x0, y0, w, h)
            assert got8.shape == (h, w, 3)
            assert np.array_equal(got8, O.preprocess_crop(frames[3], x0, y0, w, h, op))
        info = t.set_reference(frames[0], st.ref_xy)
        assert list(info.mideal) == otr.mideal
        got = t.track(np.stack(frames[1:]))
        util.compare_records(got, want, label=f"RGB 3-channel method {method}")
        if method in (1, 3, 5):   # the normalised maps: scores on cv2's own scale
            for i, (g, w) in enumerate(zip(got, want_cv)):
                assert g.status == w.status
                for k in range(5):
                    a, e = g.tm[k], w.tm[k]
                    if not e.evaluated:
                        continue
                    assert abs(a.score - e.score) <= SCORE_TOL_CV2, f"frame {i} tpl {k}"
                    if (a.ix, a.iy) == (e.ix, e.iy):
                        assert abs(a.x - e.x) <= SUBPIX_TOL_CV2 and abs(a.y - e.y) <= SUBPIX_TOL_CV2, f"frame {i} tpl {k}"


def test_rgb_three_channel_match_wide_maps_and_whole_slice(built_lib, oracle_clib):
    """The 3-channel match with result maps wider than one column tile of lst_ncc_rgb (search area 70: 141 x 141 maps) and
    as a whole-slice search (sArea = 0: the five templates share the three blurred planes of the slice)."""
    for s_area, (wd, ht), n in ((70, (480, 360), 6), (0, (320, 240), 4)):
        cfg = StackConfig(wd, ht, "u8", 24, s_area, n)
        st, frames = _rgb3_stack(cfg, n)
        op = util.oracle_params(cfg, match_intensity=False)
        otr = O.OracleTracker(op)
        otr.set_reference(frames[0], st.ref_xy)
        want = otr.run(frames[1:])
        with LaserSpotTrack4(cfg.width, cfg.height, 24, templSize=cfg.templ_size, sArea=cfg.s_area, matchIntensity=False,
                             max_batch=4) as t:
            info = t.set_reference(frames[0], st.ref_xy)
            assert list(info.mideal) == otr.mideal
            util.compare_records(t.track(np.stack(frames[1:])), want, label=f"RGB 3-channel sArea {s_area}")


def test_lanes_on_one_device_with_device_resident_slices(stack_a):
    """Several lanes on one GPU (the device listed more than once in lst_multi_create) over slices that
    already lie in device memory: records identical to the sequential oracle, stats summed over lanes."""
    import ctypes as C
    from laser_spot_track_b200 import LaserSpotTrack4Multi
    cfg, st, frames = stack_a
    op = util.oracle_params(cfg, match_intensity=False)
    otr = O.OracleTracker(op)
    otr.set_reference(frames[0], st.ref_xy)
    want = otr.run(frames[1:])
    lib = N.load()
    with _tracker(cfg, op) as single:
        stack = np.ascontiguousarray(np.stack(frames[1:]))
        fb = stack[0].nbytes
        dev = single.device_alloc(stack.nbytes)
        single._check(lib.lst_device_upload(single._ctx, C.c_void_p(dev), stack.ctypes.data, stack.nbytes))
        ptrs = [dev + i * fb for i in range(len(stack))]
        with LaserSpotTrack4Multi(cfg.width, cfg.height, 8, [0, 0, 0], templSize=op.templ_size, sArea=op.s_area,
                                  markDist=op.mark_dist, matchIntensity=False, max_batch=7) as m:
            m.set_reference(frames[0], st.ref_xy)
            got = m.track_device_ptrs(ptrs)
            util.compare_records(got, want, label="3 lanes, device slices")
            s = m.stats()
            assert s.frames >= len(want) and s.kernel_launches > 0
            assert np.array_equal(m.get_state(), np.array([v for d in otr.dis for v in d]))
        single.device_free(dev)


@pytest.mark.parametrize("dtype", ["u8", "u16", "f32"])
def test_preprocess_random_rectangles_odd_pitch(built_lib, dtype):
    """The blur kernel fetches crop rows as aligned 128-bit words and realigns them: a slice whose row
    pitch is odd gives every row another alignment, and random rectangles give every crop another start.
    Every byte must still equal the oracle's (incl. crops touching any border, widths from 14 pixels up,
    the min/max range of the 16-bit / float conventions, and wide crops that the kernel tiles)."""
    rng = np.random.default_rng({"u8": 1, "u16": 2, "f32": 3}[dtype])
    W, H = 263, 211
    if dtype == "u8":
        fr = rng.integers(0, 256, (H, W)).astype(np.uint8)
    elif dtype == "u16":
        fr = rng.integers(0, 65536, (H, W)).astype(np.uint16)
    else:
        fr = (rng.standard_normal((H, W)) * 1000.0).astype(np.float32)
    import cv2
    fr = cv2.GaussianBlur(fr.astype(np.float32), (0, 0), 1.5).astype(fr.dtype)     # some spatial structure
    cfg = StackConfig(W, H, dtype, 32, 24, 1)
    op = util.oracle_params(cfg, match_intensity=True)
    with _tracker(cfg, op) as t:
        rects = [(0, 0, W, H), (0, 0, 14, 14), (W - 14, H - 14, 14, 14), (1, 0, 262, 40), (3, 7, 200, 14)]
        for _ in range(40):
            w, h = int(rng.integers(14, 230)), int(rng.integers(14, 200))
            rects.append((int(rng.integers(0, W - w + 1)), int(rng.integers(0, H - h + 1)), w, h))
        for (x0, y0, w, h) in rects:
            got = t.preprocess(fr, x0, y0, w, h)
            want = O.preprocess_crop(fr, x0, y0, w, h, op)
            bad = np.argwhere(got != want)
            assert bad.size == 0, f"{dtype} rect {(x0, y0, w, h)}: {len(bad)} bytes differ, first {bad[:3]}"


def test_pageable_slices_are_packed_by_reader_threads(stack_a):
    """Pageable host slices (what a JNI shim holding Java arrays hands over) cannot be gathered by the copy
    engine: reader threads pack the rows of the search regions into pinned staging (ingest_mode 0 and 3);
    ingest_mode 1 still copies whole slices.  Same records either way, a fraction of the bytes."""
    cfg, st, frames = stack_a
    op = util.oracle_params(cfg, match_intensity=False)
    otr = O.OracleTracker(op)
    otr.set_reference(frames[0], st.ref_xy)
    want = otr.run(frames[1:])
    stack = np.stack(frames[1:])                      # ordinary (pageable) numpy memory
    sent = {}
    for mode in (0, 1, 3):
        with _tracker(cfg, op, max_batch=16, ingest_mode=mode) as t:
            t.set_reference(frames[0], st.ref_xy)
            util.compare_records(t.track(stack), want, label=f"pageable, ingest_mode {mode}")
            sent[mode] = t.stats().h2d_bytes
    assert sent[0] == sent[3] and sent[0] < 0.6 * sent[1]
    assert sent[1] >= stack.nbytes


@pytest.mark.parametrize("method", [0, 1, 2, 3, 4, 5])
def test_match_methods_golden_on_gpu(built_lib, method):
    """The CUDA score maps and peaks against the frozen vectors of tests/golden/match_methods_kat.npz
    (exact maps bit for bit; cv2's own maps within its float32-DFT noise) -- no oracle at run time."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "match_methods_kat.npz"))
    win, tpl = g["win"], g["tpl"]
    with LaserSpotTrack4(640, 480, 8) as t:
        res, m = t.doMatch_coord_res(win, tpl, method, False, want_map=True)
        assert np.array_equal(m.view(np.uint32), g[f"exactmap{method}"].view(np.uint32))
        cvm = g[f"cv2map{method}"]
        scale = max(1.0, float(np.abs(cvm).max()))
        assert float(np.abs(m.astype(np.float64) - cvm).max()) / scale < SCORE_TOL_CV2
        assert (res[0], res[1]) == (g[f"cv2best{method}"][0], g[f"cv2best{method}"][1])
        assert t.doMatch_test(tpl, method) == g[f"self{method}"][0]


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6, 7, 8])
def test_random_geometries_both_engines(built_lib, oracle_clib, seed):
    """Random template sizes (14..90), search areas and slice sizes: the tensor-core engine sees different
    numbers of K-steps, column blocks and partial blocks, windows clipped by the borders, result maps
    split over several CTAs (R > 128); both engines must reproduce the oracle's records."""
    rng = np.random.default_rng(1000 + seed)
    T = int(rng.integers(14, 91))
    s = int(rng.integers(6, 80))
    W = int(rng.integers(max(260, 3 * T + 2 * s + 40), 520))
    H = int(rng.integers(max(220, 3 * T + 2 * s + 40), 440))
    dtype = ["u8", "u16"][seed % 2]
    cfg = StackConfig(W, H, dtype, T, s, 6)
    st = SyntheticStack(cfg)
    frames = [st.frame(i) for i in range(0, 6)]
    op = util.oracle_params(cfg, match_intensity=(dtype == "u16") or bool(seed % 3 == 0))
    otr = O.OracleTracker(op)
    otr.set_reference(frames[0], st.ref_xy)
    want = otr.run(frames[1:])
    for engine in (1, 2):
        with _tracker(cfg, op, ncc_engine=engine, max_batch=3) as t:
            t.set_reference(frames[0], st.ref_xy)
            util.compare_records(t.track(np.stack(frames[1:])), want, label=f"seed {seed} T={T} s={s} {W}x{H} {dtype} engine {engine}")
