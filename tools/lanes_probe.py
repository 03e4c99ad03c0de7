"""Scratch probe: device-resident throughput of config B with 1, 2, 3 lanes on one GPU."""
import ctypes as C, sys, time
import numpy as np
sys.path.insert(0, ".")
from laser_spot_track_b200 import LaserSpotTrack4, LaserSpotTrack4Multi, native as N
from laser_spot_track_b200.synth import StackConfig, CONFIGS

c = CONFIGS["B"]
ring = 64
cfg = StackConfig(c.width, c.height, c.dtype, c.templ_size, c.s_area, n_frames=ring, motion="periodic", period=ring)
from laser_spot_track_b200.synth import SyntheticStack
st = SyntheticStack(cfg)
lib = N.load()
fb = cfg.width * cfg.height * 2
single = LaserSpotTrack4(cfg.width, cfg.height, 16, templSize=cfg.templ_size, sArea=cfg.s_area, max_batch=1024)
dev_ring = single.device_alloc(ring * fb)
host = np.stack([st.frame(i) for i in range(ring)])
single._check(lib.lst_device_upload(single._ctx, C.c_void_p(dev_ring), host.ctypes.data, ring * fb))
F = 10240
idx = [(1 + j) % ring for j in range(F)]
ptrs = (C.c_void_p * F)(*[dev_ring + i * fb for i in idx])
single.set_reference(st.frame(0), st.ref_xy)
out = (N.Record * F)()
for _ in range(2):
    single.track_device_ptrs(ptrs, 1, out)
t0 = time.perf_counter()
for _ in range(3):
    single.track_device_ptrs(ptrs, 1, out)
t1 = time.perf_counter()
ref = [(r.status, r.dX_pix, r.dY_pix, r.dL) for r in out]
print("1 lane", 3 * F / (t1 - t0))
for lanes in (2, 3, 4):
    with LaserSpotTrack4Multi(cfg.width, cfg.height, 16, [0] * lanes, templSize=cfg.templ_size, sArea=cfg.s_area, max_batch=1024) as m:
        m.set_reference(st.frame(0), st.ref_xy)
        o2 = (N.Record * F)()
        m.set_state(single.get_state() * 0)
        for _ in range(2):
            m.track_device_ptrs(ptrs, 1, o2)
        t0 = time.perf_counter()
        for _ in range(3):
            m.track_device_ptrs(ptrs, 1, o2)
        t1 = time.perf_counter()
        same = [(r.status, r.dX_pix, r.dY_pix, r.dL) for r in o2] == ref
        print(lanes, "lanes", 3 * F / (t1 - t0), "identical", same, "retracked", m.retracked)
