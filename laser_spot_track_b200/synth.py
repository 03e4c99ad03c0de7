"""Seeded synthetic stacks for the tracking hot path.

A frame is a smooth background ramp + Gaussian read noise + five Gaussian blobs: four
marks on the corners of a square centred in the frame and one moving spot.  Every frame
is generated from a counter-based RNG keyed by (seed, frame index), so any shard of a
sharded run can regenerate exactly its own frames.

Two motion models:
* ``survey``   -- bounded random-walk jitter of the whole scene, a slow per-mark wobble
                  and a saturating-exponential spot path; used by the
                  parity tests (integer displacements change several times).
* ``periodic`` -- every trajectory is a sum of sinusoids with period ``period`` frames, so
                  a ring of ``period`` distinct frames can be cycled seamlessly by the
                  benchmark (frame i == frame i mod period).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

SEED = 20260119


@dataclass
class StackConfig:
    width: int
    height: int
    dtype: str = "u8"            # "u8" | "u16" | "f32"
    templ_size: int = 32
    s_area: int = 32
    n_frames: int = 200
    motion: str = "survey"       # "survey" | "periodic"
    period: int = 64
    seed: int = SEED
    noise: float = 0.01          # read noise sigma, fraction of full scale
    spot_travel: float = 0.45    # spot travel as a fraction of s_area (survey motion)
    spot_amp: float = 3.0        # spot amplitude in px (periodic motion)
    blob_variety: float = 0.0    # > 0: the five blobs get different, anisotropic widths -- needed when every template is
                                 # searched over the whole slice (sArea = 0), where identical blobs are interchangeable

    @property
    def full_scale(self) -> float:
        return {"u8": 255.0, "u16": 65535.0, "f32": 1.0}[self.dtype]

    @property
    def np_dtype(self):
        return {"u8": np.uint8, "u16": np.uint16, "f32": np.float32}[self.dtype]


# the named BASELINE.json configurations
CONFIGS = {
    "A": StackConfig(640, 480, "u8", 32, 32, 200),
    "B": StackConfig(1920, 1080, "u16", 48, 56, 10000),
    "C": StackConfig(4096, 3072, "u16", 64, 0, 5000, blob_variety=0.12),
    "D": StackConfig(2048, 2048, "u8", 128, 192, 2000),
    "E": StackConfig(1920, 1080, "u16", 48, 56, 100000),
}


class SyntheticStack:
    def __init__(self, cfg: StackConfig):
        self.cfg = cfg
        c = cfg
        side = 0.5 * min(c.width, c.height)
        cx, cy = c.width / 2.0, c.height / 2.0
        h = side / 2.0
        # marks: m1 bottom-left, m2 top-left, m3 top-right, m4 bottom-right (image y points down).
        # u runs along m1->m4, v along m1->m2; the root selection of LST4:2350-2353 needs
        # (m4-m1) x (m2-m1) < 0, i.e. this orientation.
        self.base = np.array([[cx, cy],
                              [cx - h, cy + h], [cx - h, cy - h], [cx + h, cy - h], [cx + h, cy + h]], dtype=np.float64)
        # click points = true centres rounded to integers (reference frame = frame 0)
        n = c.n_frames if c.motion == "survey" else c.period
        self.n_traj = n
        rng = np.random.Generator(np.random.Philox(key=c.seed ^ 0x5EED))
        t = np.arange(n, dtype=np.float64)
        off = np.zeros((n, 5, 2))
        s = max(c.s_area, 8)
        if c.motion == "survey":
            steps = rng.normal(0.0, 0.15, size=(n, 2))
            steps[0] = 0
            walk = np.zeros((n, 2))
            lim = s / 4.0
            for i in range(1, n):
                walk[i] = np.clip(walk[i - 1] + steps[i], -lim, lim)
            off += walk[:, None, :]
            periods = [37.0, 53.0, 71.0, 89.0]
            for k in range(4):
                off[:, 1 + k, 0] += 0.5 * np.sin(2 * math.pi * t / periods[k])
                off[:, 1 + k, 1] += 0.5 * np.sin(2 * math.pi * t / periods[k] + 1.0)
            tau = max(n / 5.0, 1.0)
            d = c.spot_travel * s * (1.0 - np.exp(-t / tau))
            ang = 0.6
            off[:, 0, 0] += d * math.cos(ang)
            off[:, 0, 1] += d * math.sin(ang)
        else:
            P = float(c.period)
            w = 2 * math.pi * t / P
            scene = np.stack([0.8 * np.sin(w), 0.6 * np.sin(2 * w)], axis=1)
            off += scene[:, None, :]
            for k in range(4):
                off[:, 1 + k, 0] += 0.4 * np.sin((k + 1) * w)
                off[:, 1 + k, 1] += 0.4 * (1 - np.cos((k + 1) * w))
            off[:, 0, 0] += c.spot_amp * np.sin(w)
            off[:, 0, 1] += c.spot_amp * (1 - np.cos(w)) * 0.5
        self.offsets = off
        self.ref_xy = [(float(round(x)), float(round(y))) for x, y in self.base]
        yy, xx = np.mgrid[0:c.height, 0:c.width]
        gx, gy = 0.03 / c.width, 0.02 / c.height
        self._bg = (0.12 + gx * xx + gy * yy).astype(np.float32)

    def truth(self, i: int) -> np.ndarray:
        """True blob centres [5,2] (x,y) for frame i, order spot, mark1..4."""
        return self.base + self.offsets[i % self.n_traj]

    def frame(self, i: int) -> np.ndarray:
        c = self.cfg
        j = i % self.n_traj if c.motion == "periodic" else i
        rng = np.random.Generator(np.random.Philox(key=c.seed, counter=[j, 0, 0, 0]))
        img = self._bg + rng.standard_normal(self._bg.shape, dtype=np.float32) * np.float32(c.noise)
        sig = max(c.templ_size / 8.0, 1.5)
        rad = int(math.ceil(4 * sig))
        for k, (x, y) in enumerate(self.truth(j)):
            sx, sy = sig * (1.0 + c.blob_variety * k), sig * (1.0 + c.blob_variety * (4 - k) * 0.7)
            x0, x1 = max(int(x) - rad, 0), min(int(x) + rad + 1, c.width)
            y0, y1 = max(int(y) - rad, 0), min(int(y) + rad + 1, c.height)
            if x1 <= x0 or y1 <= y0:
                continue
            xs = np.arange(x0, x1, dtype=np.float64) - x
            ys = np.arange(y0, y1, dtype=np.float64) - y
            blob = 0.70 * np.exp(-(ys[:, None] ** 2 / (2 * sy * sy) + xs[None, :] ** 2 / (2 * sx * sx)))
            img[y0:y1, x0:x1] += blob.astype(np.float32)
        fs = c.full_scale
        if c.dtype == "f32":
            return img.astype(np.float32)
        q = np.floor(np.clip(img, 0.0, 1.0) * np.float32(fs) + np.float32(0.5))
        return q.astype(c.np_dtype)

    def frames(self, start: int, stop: int):
        for i in range(start, stop):
            yield self.frame(i)
