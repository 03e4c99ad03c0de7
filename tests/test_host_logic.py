"""Host logic of the library (no GPU): window placement, template rectangles, calcDisplacement and
the recurrence resolver -- the resolver is driven with the oracle standing in for the device and
must reproduce the sequential loop of LST4:793-846 bit for bit."""
import ctypes as C

import numpy as np
import pytest

from laser_spot_track_b200 import LaserSpotTrack4, native as N
from laser_spot_track_b200.synth import CONFIGS, StackConfig, SyntheticStack
from oracle import lst_oracle as O

import util


def test_template_rect_java_float_math(built_lib):
    lib = N.load()
    out = (C.c_int32 * 4)()
    for T in (32, 33, 48, 120, 127):
        p = N.default_params(2000, 2000, N.PIX_U8, templ_size=T)
        tr = O.OracleTracker(O.Params(2000, 2000, T, 10))
        for (x, y) in [(320.0, 240.0), (100.4, 77.6), (15.2, 9.9), (1000.5, 3.49)]:
            assert lib.lst_host_template_rect(C.byref(p), x, y, out) == 0
            half = float(np.float32(T) / np.float32(2))
            want = (O.jint(float(np.float32(np.float32(x) - np.float32(half)))),
                    O.jint(float(np.float32(np.float32(y) - np.float32(half)))),
                    O.jint(float(np.float32(2) * np.float32(half))))
            assert tuple(out)[:3] == want and out[3] == want[2]


def test_window_placement_clamp_by_shift(built_lib):
    lib = N.load()
    out = (C.c_int32 * 4)()
    W, H, T, s = 640, 480, 32, 32
    p = N.default_params(W, H, N.PIX_U8, templ_size=T, s_area=s)
    tr = O.OracleTracker(O.Params(W, H, T, s))
    rng = np.random.default_rng(1)
    for _ in range(300):
        rect = (int(rng.integers(-5, W)), int(rng.integers(-5, H)), T, T)
        dx, dy = float(rng.uniform(-40, 40)), float(rng.uniform(-40, 40))
        tr.rects = [rect] * 5
        want = tr.window_for(0, (O.jint(dx), O.jint(dy)))[:4]
        r = (C.c_int32 * 4)(*rect)
        assert lib.lst_host_window_for(C.byref(p), r, dx, dy, out) == 0
        assert tuple(out) == tuple(want)
    # negative displacement truncates toward zero: (int)-0.7 == 0  (LST4:1327)
    r = (C.c_int32 * 4)(300, 200, T, T)
    lib.lst_host_window_for(C.byref(p), r, -0.7, -1.2, out)
    assert tuple(out) == (300 - s, 200 - 1 - s, T + 2 * s, T + 2 * s)
    # sArea == 0: whole slice (LST4:1316-1320)
    p0 = N.default_params(W, H, N.PIX_U8, templ_size=T, s_area=0)
    lib.lst_host_window_for(C.byref(p0), r, 3.0, 4.0, out)
    assert tuple(out) == (0, 0, W, H)


def test_calc_displacement_kats(built_lib):
    """Inverse-bilinear solve: spot at (u, v) of the mark quadrilateral -> (u, v) * markDist."""
    L = 200.0
    m1, m2, m3, m4 = (100.0, 300.0), (100.0, 100.0), (300.0, 100.0), (300.0, 300.0)   # u: m1->m4, v: m1->m2
    for (u, v) in [(0.5, 0.5), (0.25, 0.75), (0.0, 0.0), (1.0, 1.0), (0.9, 0.1)]:
        sx, sy = m1[0] + u * L, m1[1] - v * L
        ref = [sx, sy, *m1, *m2, *m3, *m4]
        got = LaserSpotTrack4.calcDisplacement(ref, [0.0] * 10, 100.0)
        want = O.calc_displacement([(sx, sy), m1, m2, m3, m4], [(0.0, 0.0)] * 5, 100.0, None)[:5]
        assert got == list(want)
        assert abs(got[2] - 100 * u) < 1e-9 and abs(got[3] - 100 * v) < 1e-9 and got[4] == 0.0
    # perspective-skewed quadrilateral (AX != 0 branch) and a displaced state, against the oracle bit for bit
    rng = np.random.default_rng(3)
    for _ in range(200):
        ref = [float(np.float32(v)) for v in (200 + rng.uniform(-30, 30), 200 + rng.uniform(-30, 30),
                                               100 + rng.uniform(-9, 9), 300 + rng.uniform(-9, 9),
                                               100 + rng.uniform(-9, 9), 100 + rng.uniform(-9, 9),
                                               300 + rng.uniform(-9, 9), 100 + rng.uniform(-9, 9),
                                               300 + rng.uniform(-9, 9), 300 + rng.uniform(-9, 9))]
        dis = [float(v) for v in rng.uniform(-3, 3, 10)]
        spot0 = (float(rng.uniform(40, 60)), float(rng.uniform(40, 60)))
        got = LaserSpotTrack4.calcDisplacement(ref, dis, 100.0, spot0)
        want = O.calc_displacement(list(zip(ref[0::2], ref[1::2])), list(zip(dis[0::2], dis[1::2])), 100.0, spot0)[:5]
        assert got == list(want)
    # exact parallelogram: AX == 0 branch (LST4:2346-2348)
    ref = [150.0, 250.0, 100.0, 300.0, 120.0, 100.0, 320.0, 100.0, 300.0, 300.0]
    got = LaserSpotTrack4.calcDisplacement(ref, [0.0] * 10, 50.0)
    want = O.calc_displacement(list(zip(ref[0::2], ref[1::2])), [(0.0, 0.0)] * 5, 50.0, None)[:5]
    assert got == list(want)


@pytest.fixture(scope="module")
def stack_a(oracle_clib):
    cfg = CONFIGS["A"]
    st = SyntheticStack(cfg)
    return cfg, st, [st.frame(i) for i in range(0, 49)]


@pytest.mark.parametrize("max_batch", [1, 5, 16, 64])
def test_resolver_reproduces_sequential_loop(built_lib, stack_a, max_batch):
    cfg, st, frames = stack_a
    op = util.oracle_params(cfg, match_intensity=False)
    seq = O.OracleTracker(op)
    seq.set_reference(frames[0], st.ref_xy)
    want = seq.run(frames[1:])
    assert len({(w.tm[0].wx, w.tm[0].wy) for w in want}) >= 4, "window must move several times for this test to mean anything"
    tr = O.OracleTracker(op)
    tr.set_reference(frames[0], st.ref_xy)
    got, dis, stt, dev = util.host_resolve(op, tr, frames[1:], N.PIX_U8, max_batch)
    util.compare_records(got, want, label=f"max_batch={max_batch}")
    assert dis == [v for d in seq.dis for v in d]
    assert stt.frames == len(want) and stt.tasks == dev.n_tasks and stt.passes == dev.calls
    if max_batch == 1:
        assert stt.respeculated == 0


def test_resolver_skips_and_restores_state(built_lib, stack_a):
    cfg, st, frames = stack_a
    frames = [f.copy() for f in frames[:20]]
    frames[4][:] = 31                                   # mark1 rejects -> frame skipped
    c = st.truth(9)[0].astype(int)
    frames[9][c[1] - 40:c[1] + 40, c[0] - 40:c[0] + 40] = 30   # spot erased -> spot rejects
    op = util.oracle_params(cfg, match_intensity=False)
    seq = O.OracleTracker(op)
    seq.set_reference(frames[0], st.ref_xy)
    want = seq.run(frames[1:])
    assert want[3].status == 1 and want[3].reject == 1 and want[8].status == 1 and want[8].reject == 0
    tr = O.OracleTracker(op)
    tr.set_reference(frames[0], st.ref_xy)
    got, dis, _, _ = util.host_resolve(op, tr, frames[1:], N.PIX_U8, 8)
    util.compare_records(got, want, label="skips")
    assert dis == [v for d in seq.dis for v in d]


def test_resolver_search_area_doubling(built_lib, oracle_clib):
    """a jump larger than sArea is recovered by doubling (mark1: LST4:1666-1882 incl. the
    shift of the other windows; spot: LST4:2058-2128)."""
    cfg = StackConfig(400, 320, "u8", 24, 10, 12)
    st = SyntheticStack(cfg)
    frames = [st.frame(i) for i in range(0, 9)]
    # whole-scene jump of (+17, -14) px from slice 4 on, spot-only jump of +25 px in slice 7
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
    assert any(w.tm[1].s_used > cfg.s_area and w.status == 0 for w in want), "mark1 doubling not exercised"
    for mb in (1, 3, 8):
        tr = O.OracleTracker(op)
        tr.set_reference(frames[0], st.ref_xy)
        got, dis, stt, _ = util.host_resolve(op, tr, frames[1:], N.PIX_U8, mb)
        util.compare_records(got, want, label=f"doubling mb={mb}")
        assert dis == [v for d in seq.dis for v in d]
