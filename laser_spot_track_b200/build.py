"""In-tree build of liblst_b200.so (nvcc, sm_100a only).  ``python -m laser_spot_track_b200.build``."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "liblst_b200.so")
SOURCES = ["lst_kernels.cu", "lst_ncc_tc.cu", "lst_runtime.cu", "lst_host.cpp", "lst_tiff.cpp", "lst_jpeg.cpp"]
HEADERS = ["lst_math.h", "lst_host.h", "lst_tiff.h", "lst_jpeg.h", "lst_jpeg_math.h", "lst_device.h", "lst_preprocess_win.cuh", os.path.join("..", "..", "include", "lst_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",                      # one IEEE rounding per float/double op (Java / OpenCV parity)
    "-Xcompiler", "-fPIC,-ffp-contract=off,-O2,-Wall",
    "-shared", "-cudart", "static",
]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          [os.path.join(CSRC, s) for s in SOURCES] + ["-lz", "-ldl", "-o", LIB]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building liblst_b200.so")
    if verbose:
        sys.stderr.write(r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
