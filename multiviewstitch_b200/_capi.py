"""ctypes binding of include/mvs_b200.h (the C-ABI drop-in boundary).

Fails loudly: a missing shared library raises ImportError with the build command, and every
non-zero status from the library raises MvsError carrying mvs_last_error().  There is no CPU
fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("MVS_B200_LIB", PKG / "libmvs_b200.so"))
HEADER = PKG.parent / "include" / "mvs_b200.h"

MVS_OK = 0
ERR_NAMES = {-1: "INVALID_ARG", -2: "CUDA", -3: "NO_DEVICE", -4: "STATE", -5: "CAPACITY",
             -6: "DEGENERATE", -7: "COMM"}


class MvsError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"mvs_b200 status {status} ({ERR_NAMES.get(status, '?')}): {msg}")
        self.status = status


class Camera(C.Structure):
    """mvs_camera [ref: Camera/Camera.h:44-49]"""
    _fields_ = [("K", C.c_double * 9), ("R", C.c_double * 9), ("t", C.c_double * 3),
                ("w", C.c_int32), ("h", C.c_int32)]


class AlignParams(C.Structure):
    _fields_ = [("dist_thres", C.c_double), ("ground_ratio", C.c_double),
                ("remove_ground", C.c_int32), ("local_align", C.c_int32),
                ("neck_label", C.c_int32), ("reserved", C.c_int32)]


class OutlierParams(C.Structure):
    """mvs_outlier_params [ref: Processor/Processor.cpp:177-269]"""
    _fields_ = [("rounds", C.c_int32), ("iter_num", C.c_int32), ("pixel_err", C.c_double),
                ("adapt_ratio", C.c_double), ("seed", C.c_uint64)]


class GNParams(C.Structure):
    _fields_ = [("gn_iters", C.c_int32), ("pcg_iters", C.c_int32), ("corr_mode", C.c_int32),
                ("outer_rounds", C.c_int32), ("w_rot", C.c_double), ("w_reg", C.c_double),
                ("w_fit", C.c_double), ("lam", C.c_double), ("pcg_tol", C.c_double),
                ("max_corr_dist2", C.c_double), ("min_corr_cos", C.c_double)]


class ArapParams(C.Structure):
    """mvs_arap_params [ref: Deformation/Deformation.cpp:398 deform(5, 1e-4)]"""
    _fields_ = [("iters", C.c_int32), ("cg_iters", C.c_int32), ("tol", C.c_double),
                ("cg_tol", C.c_double)]


class PairJob(C.Structure):
    """mvs_pair_job [ref: Processor/Processor.cpp:629-826 one iteration of the pair loop]"""
    _fields_ = [("depth_a", C.POINTER(C.c_float)), ("cam_a", C.POINTER(Camera)),
                ("depth_b", C.POINTER(C.c_float)), ("cam_b", C.POINTER(Camera)),
                ("matches", C.POINTER(C.c_int32)), ("n_matches", C.c_int64), ("seed", C.c_uint64),
                ("kept_idx", C.POINTER(C.c_int32)), ("template_xyz", C.POINTER(C.c_double)),
                ("facets", C.POINTER(C.c_int32)), ("deformed_xyz", C.POINTER(C.c_double)),
                ("render", C.POINTER(C.c_float))]


class PairParams(C.Structure):
    _fields_ = [("min_dsp", C.c_double), ("max_dsp", C.c_double), ("smooth_threshold", C.c_double),
                ("outliers", OutlierParams), ("n_nodes", C.c_int64), ("gn", GNParams),
                ("workers", C.c_int32), ("reserved", C.c_int32)]


class PairResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("reserved", C.c_int32), ("scale", C.c_double),
                ("R", C.c_double * 9), ("t", C.c_double * 3), ("err", C.c_double),
                ("residual", C.c_double), ("n_kept", C.c_int64), ("n_template", C.c_int64),
                ("n_facets", C.c_int64), ("n_target", C.c_int64)]


def declared_symbols() -> list[str]:
    """Every function the header declares (used by the symbol-export test)."""
    txt = HEADER.read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mvs_[a-z0-9_]+)\s*\(", txt)))


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python -m multiviewstitch_b200.build` "
                "(nvcc, sm_100a). There is no CPU fallback.")
        _lib = C.CDLL(str(LIB_PATH))
        _lib.mvs_last_error.restype = C.c_char_p
    return _lib


def check(status: int) -> None:
    if status != MVS_OK:
        raise MvsError(status, lib().mvs_last_error().decode(errors="replace"))
