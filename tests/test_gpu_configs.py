"""The named BASELINE.json configurations at their full geometry against the oracle, and the
multi-lane / sharded paths on a drifting stack whose chunk boundaries are mis-speculated."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import util  # noqa: E402
from laser_spot_track_b200 import LaserSpotTrack4, LaserSpotTrack4Multi, native as N  # noqa: E402
from laser_spot_track_b200.sharded import ShardedTracker, chunk_bounds  # noqa: E402
from laser_spot_track_b200.synth import StackConfig, SyntheticStack  # noqa: E402
from oracle import lst_oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu

# north_star tolerances against the real OpenCV routine (float32 DFT noise ~5e-5)
SCORE_TOL_CV2 = 1e-4
SUBPIX_TOL_CV2 = 1e-3


def _tracker(cfg, op, **kw):
    return LaserSpotTrack4(cfg.width, cfg.height, {"u8": 8, "u16": 16, "f32": 32}[cfg.dtype], method=op.method,
                           templSize=op.templ_size, sArea=op.s_area, markDist=op.mark_dist, subPixel=op.sub_pixel,
                           matchIntensity=op.match_intensity, range_mode=op.range_mode, quant_mode=op.quant_mode,
                           doubling=op.doubling, **kw)


def _compare_cv2(got, want, label):
    near_ties = 0
    for i, (g, w) in enumerate(zip(got, want)):
        assert g.status == w.status, f"{label} frame {i}: status"
        for k in range(5):
            a, b = g.tm[k], w.tm[k]
            assert (a.wx, a.wy, a.ww, a.wh) == (b.wx, b.wy, b.ww, b.wh), f"{label} frame {i} tpl {k}: window"
            assert abs(a.score - b.score) <= SCORE_TOL_CV2, f"{label} frame {i} tpl {k}: score {a.score} vs {b.score}"
            if (a.ix, a.iy) != (b.ix, b.iy):     # only a documented near-tie may move the integer peak
                near_ties += 1
                continue
            assert abs(a.x - b.x) <= SUBPIX_TOL_CV2 and abs(a.y - b.y) <= SUBPIX_TOL_CV2, f"{label} frame {i} tpl {k}: sub-pixel"
        if w.status == 0 and near_ties == 0:
            assert abs(g.dX_pix - w.dX_pix) <= 2 * SUBPIX_TOL_CV2 and abs(g.dY_pix - w.dY_pix) <= 2 * SUBPIX_TOL_CV2
    assert near_ties <= 1, f"{label}: {near_ties} near-tie peak disagreements with OpenCV"


def test_config_b_full_size_vs_oracle(built_lib, oracle_clib):
    """BASELINE configs[1] at 1920x1080 16-bit, T=48, s=56, 24 tracked slices of a drifting scene (the spot travels
    25 px, so the integer displacements -- and with them the windows -- change many times): every field of every
    record bit-exact against the exact oracle, and within the north_star tolerances of the OpenCV-backed oracle."""
    cfg = StackConfig(1920, 1080, "u16", 48, 56, 25, spot_travel=0.9)
    st = SyntheticStack(cfg)
    frames = [st.frame(i) for i in range(25)]
    op = util.oracle_params(cfg)
    ex = O.OracleTracker(op)
    ex.set_reference(frames[0], st.ref_xy)
    want = ex.run(frames[1:])
    assert sum(1 for a, b in zip(want, want[1:]) if [O.jint(v) for d in a.dis for v in d] != [O.jint(v) for d in b.dis for v in d]) >= 5
    cv = O.OracleTracker(op, backend="cv2")
    cv.set_reference(frames[0], st.ref_xy)
    want_cv = cv.run(frames[1:])
    for mb in (0, 5):
        with _tracker(cfg, op, max_batch=mb) as t:
            info = t.set_reference(frames[0], st.ref_xy)
            assert list(info.mideal) == ex.mideal
            got = t.track_slices(frames[1:])
            util.compare_records(got, want, label=f"config B full size, max_batch {mb}")
            _compare_cv2(got, want_cv, "config B vs cv2")
            assert np.array_equal(t.get_state(), np.array([v for d in ex.dis for v in d]))


def test_config_d_2048_vs_oracle(built_lib, oracle_clib):
    """BASELINE configs[3] as SURVEY 8 states it: 2048x2048 8-bit, 128x128 templates, 512x512 search windows.  Three
    tracked slices against the OpenCV-backed oracle (tolerances), two windows bit-exact against the exact oracle."""
    cfg = StackConfig(2048, 2048, "u8", 128, 192, 8, spot_travel=0.3)
    st = SyntheticStack(cfg)
    frames = [st.frame(i) for i in range(4)]
    op = util.oracle_params(cfg, match_intensity=False)
    cv = O.OracleTracker(op, backend="cv2")
    cv.set_reference(frames[0], st.ref_xy)
    want_cv = cv.run(frames[1:])
    with _tracker(cfg, op) as t:
        t.set_reference(frames[0], st.ref_xy)
        got = t.track_slices(frames[1:])
        assert all(g.status == 0 for g in got)
        _compare_cv2(got, want_cv, "config D vs cv2")
        for fi, k in ((0, 0), (2, 3)):
            m = got[fi].tm[k]
            assert (m.ww, m.wh) == (512, 512)
            win8 = O.preprocess_crop(frames[1 + fi], m.wx, m.wy, m.ww, m.wh, op)
            ex, _ = O.do_match_coord_res(win8, cv.tpl8[k], True, "exact")
            assert [m.x, m.y, m.score] == ex, f"config D frame {fi} tpl {k}"


def _exact_scores_3x3(win8, tpl8, ix, iy):
    """TM_CCOEFF_NORMED of the nine positions around (ix, iy) with the oracle's exact arithmetic."""
    th, tw = tpl8.shape
    corr = np.zeros((3, 3), dtype=np.int64)
    s = np.zeros((3, 3), dtype=np.int64)
    q = np.zeros((3, 3), dtype=np.int64)
    t64 = tpl8.astype(np.int64)
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            patch = win8[iy + dy:iy + dy + th, ix + dx:ix + dx + tw].astype(np.int64)
            corr[dy + 1, dx + 1] = int((patch * t64).sum())
            s[dy + 1, dx + 1] = int(patch.sum())
            q[dy + 1, dx + 1] = int((patch * patch).sum())
    return O.match_from_sums(5, corr, s, q, tpl8)


def test_config_c_full_frame_vs_oracle(built_lib, oracle_clib):
    """BASELINE configs[2]: 4096x3072 16-bit, 64x64 templates, sArea = 0 => all five templates are searched over the
    whole slice (LST4:1316-1320, 1451-1454).  One tracked slice: the blurred 8-bit slice byte-exact against the oracle,
    peaks / scores / sub-pixel positions within the north_star tolerances of OpenCV's matchTemplate on the same image,
    and the peak score bit-exact against the exact arithmetic at that position."""
    cfg = StackConfig(4096, 3072, "u16", 64, 0, 4, blob_variety=0.12)    # distinct blobs: with identical ones any template fits any blob
    st = SyntheticStack(cfg)
    f0, f1 = st.frame(0), st.frame(1)
    op = util.oracle_params(cfg)
    cv = O.OracleTracker(op, backend="cv2")
    cv.set_reference(f0, st.ref_xy)
    want = cv.run([f1])
    with _tracker(cfg, op) as t:
        t.set_reference(f0, st.ref_xy)
        got = t.track_slices([f1])
        g, w = got[0], want[0]
        assert g.status == w.status == 0
        win8 = O.preprocess_crop(f1, 0, 0, cfg.width, cfg.height, op)
        dev8 = t.preprocess(f1, 0, 0, cfg.width, cfg.height)
        assert np.array_equal(dev8, win8), "blurred 8-bit slice differs from the oracle"
        for k in range(5):
            a, b = g.tm[k], w.tm[k]
            assert (a.wx, a.wy, a.ww, a.wh) == (0, 0, cfg.width, cfg.height)
            assert (a.ix, a.iy) == (b.ix, b.iy), f"tpl {k}: peak {(a.ix, a.iy)} vs OpenCV {(b.ix, b.iy)}"
            assert abs(a.score - b.score) <= SCORE_TOL_CV2
            assert abs(a.x - b.x) <= SUBPIX_TOL_CV2 and abs(a.y - b.y) <= SUBPIX_TOL_CV2
            nb = _exact_scores_3x3(win8, cv.tpl8[k], a.ix, a.iy)
            assert float(nb[1, 1]) == a.score, f"tpl {k}: peak score {a.score!r} vs exact {float(nb[1, 1])!r}"
            dx, dy = O.subpixel_fit(nb, 1, 1)
            assert (a.x, a.y) == (a.ix + dx, a.iy + dy), f"tpl {k}: sub-pixel fit"
        truth = st.truth(1)
        for k in range(5):      # every template found its own blob
            assert abs(g.tm[k].x + cfg.templ_size / 2 - truth[k, 0]) < 1.5 and abs(g.tm[k].y + cfg.templ_size / 2 - truth[k, 1]) < 1.5
        assert abs(g.x_abs - w.x_abs) <= 1e-3 and abs(g.y_abs - w.y_abs) <= 1e-3


def _drift_stack():
    # the spot drifts 38 px and the scene jitters: the state at every chunk boundary differs from the state at call time
    cfg = StackConfig(640, 480, "u8", 32, 32, 61, spot_travel=1.2)
    st = SyntheticStack(cfg)
    return cfg, st, [st.frame(i) for i in range(61)]


def test_multi_lanes_retrack_on_drifting_stack(built_lib, oracle_clib, monkeypatch):
    """lst_multi_* with no lead-in: every chunk starts from the state at call time, which is wrong at every boundary of
    a drifting stack -- the boundary check must catch it, re-track those chunks and still return the sequential
    records (the re-track path had only ever run in the zero-mis-speculation case)."""
    cfg, st, frames = _drift_stack()
    op = util.oracle_params(cfg, match_intensity=False)
    otr = O.OracleTracker(op)
    otr.set_reference(frames[0], st.ref_xy)
    want = otr.run(frames[1:])
    assert [O.jint(v) for d in want[19].dis for v in d] != [0] * 10
    ndev = N.load().lst_device_count()
    for lead_in, expect_retrack in (("0", True), ("2", False)):
        monkeypatch.setenv("LST_MULTI_LEAD_IN", lead_in)
        with LaserSpotTrack4Multi(cfg.width, cfg.height, 8, [i % ndev for i in range(3)], templSize=op.templ_size,
                                  sArea=op.s_area, matchIntensity=False, max_batch=8) as m:
            m.set_reference(frames[0], st.ref_xy)
            got = m.track(np.stack(frames[1:]))
            util.compare_records(got, want, label=f"3 lanes, lead-in {lead_in}")
            assert np.array_equal(m.get_state(), np.array([v for d in otr.dis for v in d]))
            if expect_retrack:
                assert m.retracked >= 1, "no chunk was re-tracked although every boundary guess was wrong"


def test_multi_lanes_with_doubling_and_lost_mark(built_lib, oracle_clib, monkeypatch):
    """The same with search-area doubling on and a scene jump that loses mark1 right after a chunk boundary: the shift of
    the other four windows uses mark1's un-truncated displacement (LST4:1757), so the boundary states must agree exactly
    there."""
    cfg = StackConfig(640, 480, "u8", 32, 16, 31, spot_travel=0.5)
    st = SyntheticStack(cfg)
    frames = [st.frame(i) for i in range(31)]
    for i in range(11, 31):                      # the whole scene jumps by more than sArea after slice 10
        frames[i] = np.roll(frames[i], (9, 21), axis=(0, 1))
    op = util.oracle_params(cfg, match_intensity=False, doubling=True)
    otr = O.OracleTracker(op)
    otr.set_reference(frames[0], st.ref_xy)
    want = otr.run(frames[1:])
    assert any(w.tm[1].s_used > cfg.s_area for w in want), "the jump did not trigger the doubling path"
    ndev = N.load().lst_device_count()
    for lead_in in ("0", "2"):
        monkeypatch.setenv("LST_MULTI_LEAD_IN", lead_in)
        with LaserSpotTrack4Multi(cfg.width, cfg.height, 8, [i % ndev for i in range(3)], templSize=op.templ_size,
                                  sArea=op.s_area, matchIntensity=False, doubling=True, max_batch=8) as m:
            m.set_reference(frames[0], st.ref_xy)
            got = m.track(np.stack(frames[1:]))
            util.compare_records(got, want, label=f"doubling, 3 lanes, lead-in {lead_in}")


class _ReplayComm:
    """Runs the two 'ranks' of a ShardedTracker one after the other in this process.  Rank 0 goes first: it never needs
    rank 1's values (its own chunk starts from the true state).  Rank 1 then sees rank 0's gathered values."""

    def __init__(self, rank, world, board):
        self.rank, self.world, self.board = rank, world, board

    def all_gather(self, vals):
        self.board[self.rank] = list(vals)
        return [self.board.get(r, list(vals)) for r in range(self.world)]

    def broadcast(self, vals, src, n=10):
        if self.rank == src:
            self.board[("b", src)] = list(vals)
            return list(vals)
        return list(self.board.get(("b", src), [0.0] * n))     # rank 0 does not use what rank 1 would broadcast


def test_sharded_tracker_retrack_on_gpu(built_lib, oracle_clib):
    """ShardedTracker (the N > 1 path of bench.py) on the GPU with a wrong boundary guess: rank 1 starts its chunk from
    the reference state, must notice at the boundary exchange, re-track, and return the sequential records."""
    cfg, st, frames = _drift_stack()
    op = util.oracle_params(cfg, match_intensity=False)
    otr = O.OracleTracker(op)
    otr.set_reference(frames[0], st.ref_xy)
    want = otr.run(frames[1:])
    bounds = chunk_bounds(len(frames) - 1, 2)
    board, out = {}, {}
    for rank in (0, 1):
        with _tracker(cfg, op, max_batch=8) as t:
            t.set_reference(frames[0], st.ref_xy)
            sh = ShardedTracker(t, _ReplayComm(rank, 2, board))
            lo, hi = bounds[rank]
            recs, end = sh.track_chunk(np.stack(frames[1 + lo:1 + hi]), 1 + lo, [0.0] * 10)
            out[rank] = (list(recs), list(end), sh.retracked)
    got = out[0][0] + out[1][0]
    util.compare_records(got, want, label="sharded, 2 ranks")
    assert out[1][2] == 1, "rank 1 speculated the reference state and must have re-tracked"
    assert out[1][1] == [v for d in otr.dis for v in d]


def _rec_key(r):
    return (int(r.frame), r.status, [(m.ix, m.iy, m.wx, m.wy, m.score, m.x, m.y, m.accepted) for m in r.tm], list(r.dis),
            (r.dX_pix, r.dY_pix, r.x_abs, r.y_abs, r.dL))


def test_submit_fetch_override(built_lib, oracle_clib):
    """lst_submit / lst_fetch / lst_override.  Batches are queued while earlier ones run (LST_WOULD_BLOCK when
    the queue is full), the records come back in slice order and equal lst_track's; an override drops everything after
    the overridden slice and restarts the recurrence from the caller's state."""
    import ctypes as C
    cfg, st, frames = _drift_stack()
    op = util.oracle_params(cfg, match_intensity=False)
    otr = O.OracleTracker(op)
    otr.set_reference(frames[0], st.ref_xy)
    want = otr.run(frames[1:])
    lib = N.load()
    stack = np.ascontiguousarray(np.stack(frames[1:]))
    ptr = C.c_void_p()
    assert lib.lst_alloc_pinned(stack.nbytes, C.byref(ptr)) == 0
    try:
        view = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(stack.nbytes,)).view(stack.dtype).reshape(stack.shape)
        view[...] = stack
        addr = [view[i].ctypes.data for i in range(len(view))]
        with _tracker(cfg, op, max_batch=8) as t:
            t.set_reference(frames[0], st.ref_xy)
            got, nxt, blocked = [], 0, 0
            while len(got) < len(addr):
                if nxt < len(addr):
                    n = min(7, len(addr) - nxt)
                    if t.submit(addr[nxt:nxt + n], 1 + nxt):
                        nxt += n
                        continue
                    blocked += 1
                got += t.fetch(5, wait=True)
            assert blocked >= 1, "the queue never filled up: nothing was in flight while the caller went on"
            assert t.fetch(5) == []
            util.compare_records(got, want, label="submit/fetch")
            assert [int(r.frame) for r in got] == list(range(1, len(addr) + 1))
            # override: keep slice 20's state, drop what follows (some of it finished, some queued), resume from 21
            t.set_state([0.0] * 10)
            assert t.submit(addr[:30], 1) and t.submit(addr[30:45], 31)
            first = t.fetch(12, wait=True)
            resume = t.override(20, list(want[19].dis[k][j] for k in range(5) for j in range(2)))
            assert resume == 21
            rest = []
            while True:
                part = t.fetch(64)
                if not part:
                    break
                rest += part
            kept = first + rest
            assert [int(r.frame) for r in kept] == list(range(1, 21)), "records after the overridden slice must be dropped"
            assert t.submit(addr[20:], 21)
            tail = []
            while len(tail) < len(addr) - 20:
                tail += t.fetch(64, wait=True)
            util.compare_records(kept + tail, want, label="after override")
    finally:
        lib.lst_free_pinned(ptr)
