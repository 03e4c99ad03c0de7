"""Diagnosis of the JPEG slice-file leg on a GPU box: time lst_track_files over JPEG files for several reader-thread counts."""
import os, sys, time, tempfile, shutil
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, cv2
from laser_spot_track_b200 import LaserSpotTrack4
from laser_spot_track_b200.synth import SyntheticStack
import bench

cfg = bench.bench_config("B", 64)
st = SyntheticStack(cfg)
d = tempfile.mkdtemp(prefix="jl_", dir="/dev/shm")
files = []
for i in range(32):
    g = (st.frame(i) >> 8).astype(np.uint8)
    p = os.path.join(d, f"a{i}.jpg")
    cv2.imwrite(p, np.stack([g, g, g], -1), [cv2.IMWRITE_JPEG_QUALITY, 92, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420])
    files.append(p)
paths = [files[i % 32] for i in range(1, 1025)]
print("cores", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for nt in (1, 4, 16):
    for mb in (1024, 128):
        os.environ["LST_FILE_THREADS"] = str(nt)
        with LaserSpotTrack4(cfg.width, cfg.height, 24, templSize=cfg.templ_size, sArea=cfg.s_area, max_batch=mb) as t:
            t.set_reference_file(files[0], st.ref_xy)
            t.track_files(paths[:64])
            t.set_state([0.0] * 10)
            t.reset_stats()
            n = 1024 if nt > 1 else 128
            t0 = time.perf_counter()
            r = t.track_files(paths[:n])
            dt = time.perf_counter() - t0
            s = t.stats()
            print(f"threads {nt} batch {mb}: {n / dt:.0f} files/s, {dt * 1e3 / n * nt:.2f} ms per file-thread, passes {s.passes} h2d {s.h2d_bytes / n:.0f} B/slice "
                  f"fallback {s.roi_fallback_tasks} host_ms {s.host_ms:.0f} wait_ms {s.device_wait_ms:.0f} ok {sum(1 for x in r if x.status == 0)}")
shutil.rmtree(d, ignore_errors=True)
