"""multiviewstitch_b200 — B200-native (sm_100a) non-rigid alignment hot path of MultiViewStitch.

The product is the C-ABI shared library libmvs_b200.so (include/mvs_b200.h, sources in csrc/);
this package is its Python host-side mirror of the reference's operator interface
(api.SRTSolver, api.Deformation, api.Model2Depth, api.PartRecognition, api.Alignment).
"""
from ._capi import Camera, GNParams, AlignParams, MvsError, lib, declared_symbols  # noqa: F401
from . import api  # noqa: F401

__all__ = ["api", "Camera", "GNParams", "AlignParams", "MvsError", "lib", "declared_symbols"]
