"""laser_spot_track_b200 -- B200-native (sm_100a) tracking hot path of Laser Spot Track 4.

Only what the path needs: ``csrc/`` (CUDA kernels + the C ABI of include/lst_b200.h),
``native`` (ctypes stub of that ABI), ``plugin`` (host-side mirror of the reference plugin's
interface), ``sharded`` (frame sharding over several GPUs) and ``synth`` (seeded synthetic stacks).
"""
from . import native  # noqa: F401
from .plugin import LaserSpotTrack4, LaserSpotTrack4Multi, LstError, results_rows, results_table, tiff_probe, tiff_read  # noqa: F401

__all__ = ["LaserSpotTrack4", "LaserSpotTrack4Multi", "LstError", "results_rows", "results_table", "tiff_probe", "tiff_read", "native"]
