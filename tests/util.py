"""Helpers shared by the tests: oracle-backed stand-in for the device (CPU tests only) and
record comparison."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from laser_spot_track_b200 import native as N
from oracle import lst_oracle as O


def oracle_params(cfg, **kw) -> O.Params:
    p = O.Params(cfg.width, cfg.height, cfg.templ_size, cfg.s_area)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def native_params(op: O.Params, pixel_type: int, max_batch: int = 0) -> N.Params:
    return N.default_params(op.width, op.height, pixel_type, method=op.method, templ_size=op.templ_size, s_area=op.s_area,
                            mark_dist=op.mark_dist, sub_pixel=int(op.sub_pixel),
                            match_intensity=int(op.match_intensity), range_mode=op.range_mode,
                            range_lo=op.range_lo, range_hi=op.range_hi, quant_mode=op.quant_mode,
                            doubling=int(op.doubling), auto_skip=int(op.auto_skip), max_s_area=op.max_s_area,
                            max_batch=max_batch)


def pix_type(dtype) -> int:
    return {np.dtype(np.uint8): N.PIX_U8, np.dtype(np.uint16): N.PIX_U16, np.dtype(np.float32): N.PIX_F32}[np.dtype(dtype)]


class OracleDevice:
    """Stands in for the GPU behind lst_host_resolve: evaluates (frame, template, window) tasks
    with the oracle.  CPU tests of the resolver only."""

    def __init__(self, tracker: O.OracleTracker, frames, frame_base: int = 0):
        self.tr = tracker
        self.frames = frames
        self.frame_base = frame_base
        self.calls = 0
        self.n_tasks = 0
        self.fn = N.COMPUTE_FN(self._compute)

    def _compute(self, user, tasks, n, results):
        self.calls += 1
        self.n_tasks += n
        try:
            for i in range(n):
                t = tasks[i]
                tm = self.tr._match(self.frames[t.frame - self.frame_base], t.tpl, t.wx, t.wy, t.ww, t.wh, t.s_used)
                r = results[i]
                r.x, r.y, r.score, r.ix, r.iy, r.accepted = tm.x, tm.y, tm.score, tm.ix, tm.iy, int(tm.accepted)
            return 0
        except Exception as e:  # pragma: no cover
            print("oracle device failed:", e)
            return -3


def host_resolve(op: O.Params, tracker: O.OracleTracker, frames, pixel_type, max_batch, dis0=None, first_index=1,
                 frame_base=0):
    """Run the library's resolver on CPU with the oracle as the device."""
    lib = N.load()
    p = native_params(op, pixel_type, max_batch)
    rect = ((C.c_int32 * 4) * 5)()
    for k in range(5):
        for j in range(4):
            rect[k][j] = tracker.rects[k][j]
    ref = (C.c_float * 10)(*[v for xy in tracker.ref_xy for v in xy])
    spot0 = (C.c_double * 2)(*tracker.spot0)
    dis = (C.c_double * 10)(*(dis0 if dis0 is not None else [0.0] * 10))
    n = len(frames)
    out = (N.Record * n)()
    st = N.Stats()
    dev = OracleDevice(tracker, frames, frame_base)
    rc = lib.lst_host_resolve(C.byref(p), C.byref(rect), ref, spot0, dis, first_index, n, dev.fn, None, out, C.byref(st))
    assert rc == 0, rc
    return out, list(dis), st, dev


def same_float(a: float, b: float, tol: float = 0.0) -> bool:
    if math.isnan(a) and math.isnan(b):
        return True
    if a == b:
        return True
    return abs(a - b) <= tol


def compare_records(got, want, *, score_tol=0.0, pos_tol=0.0, out_tol=0.0, label=""):
    """got: lst_record array; want: list of oracle FrameRecord.  tol 0 => bit-exact."""
    assert len(got) == len(want)
    for i, (g, w) in enumerate(zip(got, want)):
        where = f"{label} frame {i}"
        assert g.status == w.status, f"{where}: status {g.status} != {w.status}"
        assert g.reject == w.reject, f"{where}: reject {g.reject} != {w.reject}"
        for k in range(5):
            a, b = g.tm[k], w.tm[k]
            if not b.evaluated:
                continue
            assert (a.wx, a.wy, a.ww, a.wh) == (b.wx, b.wy, b.ww, b.wh), f"{where} tpl {k}: window {(a.wx, a.wy, a.ww, a.wh)} != {(b.wx, b.wy, b.ww, b.wh)}"
            assert (a.ix, a.iy) == (b.ix, b.iy), f"{where} tpl {k}: peak {(a.ix, a.iy)} != {(b.ix, b.iy)}"
            assert same_float(a.score, b.score, score_tol), f"{where} tpl {k}: score {a.score!r} != {b.score!r}"
            assert same_float(a.x, b.x, pos_tol) and same_float(a.y, b.y, pos_tol), f"{where} tpl {k}: pos {(a.x, a.y)} != {(b.x, b.y)}"
            assert bool(a.accepted) == bool(b.accepted), f"{where} tpl {k}: accepted"
        for j in range(5):
            gd = (g.dis[2 * j], g.dis[2 * j + 1])
            assert same_float(gd[0], w.dis[j][0], pos_tol) and same_float(gd[1], w.dis[j][1], pos_tol), f"{where}: dis[{j}] {gd} != {w.dis[j]}"
        if w.status == 0:
            for name in ("dX_pix", "dY_pix", "x_abs", "y_abs", "dL"):
                assert same_float(getattr(g, name), getattr(w, name), out_tol), f"{where}: {name} {getattr(g, name)!r} != {getattr(w, name)!r}"


class CpuResolverTracker:
    """Tracker-shaped CPU stand-in for the sharding tests: the library's resolver (C++) with the
    oracle as the device.  Same set_state/get_state/track surface as LaserSpotTrack4."""

    def __init__(self, op: O.Params, ref_frame, ref_xy, pixel_type, max_batch=8):
        self.op, self.pixel_type, self.max_batch = op, pixel_type, max_batch
        self.otr = O.OracleTracker(op)
        self.otr.set_reference(ref_frame, ref_xy)
        self.state = [0.0] * 10
        self.tasks = 0

    def set_state(self, dis):
        self.state = [float(v) for v in dis]

    def get_state(self):
        return list(self.state)

    def track(self, frames, first_frame_index=1):
        frames = list(frames)
        out, dis, st, dev = host_resolve(self.op, self.otr, frames, self.pixel_type, self.max_batch,
                                         dis0=self.state, first_index=first_frame_index)
        self.state = dis
        self.tasks += st.tasks
        return out
