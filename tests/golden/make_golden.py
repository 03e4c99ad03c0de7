#!/usr/bin/env python
"""Regenerates tests/golden/*.npz from the oracle (run from the repo root:
``python tests/golden/make_golden.py``).

PARITY UNPINNED by the reference: it ships no vectors and cannot run here (no JVM).  These
fixtures freeze (a) the oracle's own outputs on seeded synthetic stacks, so that any later change
to the oracle or to the CUDA path is caught, and (b) outputs of the real third-party routine
available in this image -- cv2.matchTemplate / cv2.minMaxLoc (OpenCV 4.13; the reference pins
OpenCV 4.6.0 through javacv 1.5.8) -- which the oracle's restatement is checked against.
Assumptions frozen with the vectors: float->8-bit range convention RANGE_DISPLAY + QUANT_ROUND255
(oracle/lst_oracle.py), ImageJ GaussianBlur restated from recollection.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from laser_spot_track_b200.synth import StackConfig, SyntheticStack  # noqa: E402
from oracle import lst_oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def records_to_arrays(recs):
    n = len(recs)
    a = dict(status=np.zeros(n, np.int32), reject=np.zeros(n, np.int32), peaks=np.zeros((n, 5, 2), np.int32),
             windows=np.zeros((n, 5, 4), np.int32), scores=np.zeros((n, 5)), pos=np.zeros((n, 5, 2)),
             accepted=np.zeros((n, 5), np.int32), out=np.zeros((n, 5)), dis=np.zeros((n, 10)))
    for i, r in enumerate(recs):
        a["status"][i], a["reject"][i] = r.status, r.reject
        for k in range(5):
            t = r.tm[k]
            a["peaks"][i, k] = (t.ix, t.iy)
            a["windows"][i, k] = (t.wx, t.wy, t.ww, t.wh)
            a["scores"][i, k] = t.score
            a["pos"][i, k] = (t.x, t.y)
            a["accepted"][i, k] = int(t.accepted)
        a["out"][i] = (r.dX_pix, r.dY_pix, r.x_abs, r.y_abs, r.dL)
        a["dis"][i] = [v for d in r.dis for v in d]
    return a


def track_fixture(name, cfg, n, **kw):
    st = SyntheticStack(cfg)
    frames = [st.frame(i) for i in range(n + 1)]
    op = O.Params(cfg.width, cfg.height, cfg.templ_size, cfg.s_area, **kw)
    tr = O.OracleTracker(op, "exact")
    ref = tr.set_reference(frames[0], st.ref_xy)
    recs = tr.run(frames[1:])
    a = records_to_arrays(recs)
    np.savez_compressed(os.path.join(HERE, name), cfg=np.array([cfg.width, cfg.height, cfg.templ_size, cfg.s_area, n]),
                        method=int(op.method),
                        dtype=cfg.dtype, ref_xy=np.array(st.ref_xy), rects=np.array(tr.rects), mideal=np.array(tr.mideal),
                        ref_out=np.array([ref.dX_pix, ref.dY_pix, ref.x_abs, ref.y_abs, ref.dL]),
                        match_intensity=int(op.match_intensity), frame_crc=np.array([int(f.astype(np.int64).sum()) for f in frames]), **a)
    print(name, "status", a["status"].tolist())


def match_fixture():
    """8-bit window/template pairs with cv2's own result map, peak and the exact map."""
    import cv2
    cv2.setNumThreads(1)
    rng = np.random.default_rng(20260119)
    out = {}
    for i, (S, T) in enumerate([(96, 32), (160, 48), (70, 21), (64, 14)]):
        base = cv2.GaussianBlur(rng.integers(0, 256, size=(S + 30, S + 30), dtype=np.uint8), (0, 0), 2.0)
        win = np.ascontiguousarray(base[:S, :S])
        tpl = np.ascontiguousarray(base[9:9 + T, 13:13 + T])
        out[f"win{i}"], out[f"tpl{i}"] = win, tpl
        out[f"cv2map{i}"] = O.ncc_map_cv2(win, tpl)
        out[f"exactmap{i}"] = O.ncc_map_exact(win, tpl)
        mn, mx, mnl, mxl = cv2.minMaxLoc(out[f"cv2map{i}"])
        out[f"cv2peak{i}"] = np.array([mxl[0], mxl[1], mx])
    np.savez_compressed(os.path.join(HERE, "match_kat.npz"), **out)
    print("match_kat ok")


def methods_fixture():
    """the other five methods of the dialog: cv2's own maps and min/max locations on fixed
    8-bit inputs, plus the oracle's exact maps."""
    import cv2
    cv2.setNumThreads(1)
    rng = np.random.default_rng(20260120)
    base = cv2.GaussianBlur(rng.integers(0, 256, size=(130, 130), dtype=np.uint8), (0, 0), 2.0)
    win = np.ascontiguousarray(base[:100, :110])
    tpl = np.ascontiguousarray(base[31:63, 40:72])
    out = dict(win=win, tpl=tpl)
    for m in range(6):
        cvm = O.ncc_map_cv2(win, tpl, m)
        out[f"cv2map{m}"] = cvm
        out[f"exactmap{m}"] = O.ncc_map_exact(win, tpl, m)
        mn, mx, mnl, mxl = cv2.minMaxLoc(cvm)
        out[f"cv2best{m}"] = np.array([mnl[0], mnl[1], mn] if m in (0, 1) else [mxl[0], mxl[1], mx])
        out[f"self{m}"] = np.array([O.do_match_test(tpl, "exact", m), O.do_match_test(tpl, "cv2", m)])
    np.savez_compressed(os.path.join(HERE, "match_methods_kat.npz"), **out)
    print("match_methods_kat ok")


def blur_fixture():
    rng = np.random.default_rng(7)
    img8 = rng.integers(0, 256, size=(40, 52), dtype=np.uint8)
    img16 = rng.integers(0, 65536, size=(33, 47), dtype=np.uint16)
    kern, ksum = O.make_gaussian_kernel()
    np.savez_compressed(os.path.join(HERE, "blur_kat.npz"), kern=kern, ksum=ksum, img8=img8, img16=img16,
                        blur8=O.blur_gaussian_float(img8.astype(np.float32)),
                        blur16=O.blur_gaussian_float(img16.astype(np.float32)),
                        q8=O.preprocess_crop(img8, 0, 0, 52, 40, O.Params(52, 40, 14, 4, match_intensity=False)),
                        q16=O.preprocess_crop(img16, 0, 0, 47, 33, O.Params(47, 33, 14, 4)))
    print("blur_kat ok")


if __name__ == "__main__":
    track_fixture("track_a_u8.npz", StackConfig(640, 480, "u8", 32, 32, 200), 40, match_intensity=False)
    track_fixture("track_small_u16.npz", StackConfig(480, 400, "u16", 48, 56, 24), 12)
    for m in (0, 1, 2, 3, 4):
        track_fixture(f"track_method{m}_u8.npz", StackConfig(640, 480, "u8", 32, 32, 200), 16, match_intensity=False, method=m)
    match_fixture()
    methods_fixture()
    blur_fixture()
