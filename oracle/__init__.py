"""oracle — TEST INFRASTRUCTURE ONLY.

ctypes front-end of the CPU oracle (oracle/*.cpp, see oracle/oracle.h for what it restates and
how it is pinned).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this package; the product (multiviewstitch_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "liboracle.so"
_lib = None


def build(force: bool = False) -> Path:
    """Compile the oracle with the committed Makefile (g++ only)."""
    srcs = list(_HERE.glob("*.cpp")) + list(_HERE.glob("*.h")) + [_HERE / "Makefile"]
    stale = (not _LIB_PATH.exists()) or any(s.stat().st_mtime > _LIB_PATH.stat().st_mtime for s in srcs)
    if force or stale:
        r = subprocess.run(["make", "-C", str(_HERE), "-j8"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    return _LIB_PATH


class Camera(C.Structure):
    _fields_ = [("K", C.c_double * 9), ("R", C.c_double * 9), ("t", C.c_double * 3),
                ("w", C.c_int32), ("h", C.c_int32)]


class GNParams(C.Structure):
    _fields_ = [("gn_iters", C.c_int32), ("pcg_iters", C.c_int32), ("corr_mode", C.c_int32),
                ("outer_rounds", C.c_int32), ("w_rot", C.c_double), ("w_reg", C.c_double),
                ("w_fit", C.c_double), ("lam", C.c_double), ("pcg_tol", C.c_double),
                ("max_corr_dist2", C.c_double), ("min_corr_cos", C.c_double)]


def gn_params(gn_iters=10, pcg_iters=20, corr_mode=2, outer_rounds=1, w_rot=1.0, w_reg=10.0,
              w_fit=100.0, lam=1e-6, pcg_tol=0.0, max_corr_dist2=0.0, min_corr_cos=-2.0):
    return GNParams(gn_iters, pcg_iters, corr_mode, outer_rounds, w_rot, w_reg, w_fit, lam,
                    pcg_tol, max_corr_dist2, min_corr_cos)


def make_camera(K, R, t, w, h) -> Camera:
    c = Camera()
    c.K[:] = np.asarray(K, np.float64).reshape(9).tolist()
    c.R[:] = np.asarray(R, np.float64).reshape(9).tolist()
    c.t[:] = np.asarray(t, np.float64).reshape(3).tolist()
    c.w, c.h = int(w), int(h)
    return c


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_LIB_PATH))
        _lib.orc_srt_residual.restype = C.c_double
        _lib.orc_srt_scale.restype = C.c_double
        _lib.orc_srt_ransac.restype = C.c_double
        for name in ("orc_uniform_sampling", "orc_sample_fps"):
            if hasattr(_lib, name):
                getattr(_lib, name).restype = C.c_int64
    return _lib


def _p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct)) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


# ----------------------------------------------------------------------------- NN
def knn(q, t, k, method="brute"):
    q = _f32(q).reshape(-1, 3)
    t = _f32(t).reshape(-1, 3)
    idx = np.empty((len(q), k), np.int32)
    d2 = np.empty((len(q), k), np.float32)
    fn = lib().orc_knn_brute if method == "brute" else lib().orc_knn_kdtree
    fn(_p(q, C.c_float), C.c_int64(len(q)), _p(t, C.c_float), C.c_int64(len(t)), C.c_int(k),
       _p(idx, C.c_int32), _p(d2, C.c_float))
    return idx, d2


def part_recog(tmpl, labels, query):
    tmpl = _f64(tmpl).reshape(-1, 3)
    query = _f64(query).reshape(-1, 3)
    labels = _i32(labels)
    out = np.empty(len(query), np.int32)
    lib().orc_part_recog(_p(tmpl, C.c_double), C.c_int64(len(tmpl)), _p(labels, C.c_int32),
                         _p(query, C.c_double), C.c_int64(len(query)), _p(out, C.c_int32))
    return out


# ----------------------------------------------------------------------------- camera / SRT
def project(cam: Camera, p):
    p = _f64(p).reshape(3)
    u, v = C.c_int32(), C.c_int32()
    lib().orc_project(C.byref(cam), _p(p, C.c_double), C.byref(u), C.byref(v))
    return u.value, v.value


def unproject(cam: Camera, u, v, d):
    out = np.empty(3, np.float64)
    lib().orc_unproject(C.byref(cam), C.c_int(u), C.c_int(v), C.c_double(d), _p(out, C.c_double))
    return out


def srt_residual(p, q, cam1, cam2, scale, R, t):
    p, q = _f64(p).reshape(-1, 3), _f64(q).reshape(-1, 3)
    R, t = _f64(R).reshape(9), _f64(t).reshape(3)
    return lib().orc_srt_residual(_p(p, C.c_double), _p(q, C.c_double), C.c_int64(len(p)),
                                  C.byref(cam1), C.byref(cam2), C.c_double(scale),
                                  _p(R, C.c_double), _p(t, C.c_double))


def srt_scale(p, q):
    p, q = _f64(p).reshape(-1, 3), _f64(q).reshape(-1, 3)
    return lib().orc_srt_scale(_p(p, C.c_double), _p(q, C.c_double), C.c_int64(len(p)))


def srt_estimate(p, q):
    p, q = _f64(p).reshape(-1, 3), _f64(q).reshape(-1, 3)
    s = C.c_double()
    R = np.empty(9, np.float64)
    t = np.empty(3, np.float64)
    lib().orc_srt_estimate(_p(p, C.c_double), _p(q, C.c_double), C.c_int64(len(p)), C.byref(s),
                           _p(R, C.c_double), _p(t, C.c_double))
    return s.value, R.reshape(3, 3), t


def srt_ransac(p, q, cam1, cam2, triples):
    p, q = _f64(p).reshape(-1, 3), _f64(q).reshape(-1, 3)
    triples = _i32(triples).reshape(-1, 3)
    s = C.c_double()
    R = np.empty(9, np.float64)
    t = np.empty(3, np.float64)
    best = C.c_int32()
    errs = np.empty(len(triples), np.float64)
    e = lib().orc_srt_ransac(_p(p, C.c_double), _p(q, C.c_double), C.c_int64(len(p)),
                             C.byref(cam1), C.byref(cam2), _p(triples, C.c_int32),
                             C.c_int(len(triples)), C.byref(s), _p(R, C.c_double),
                             _p(t, C.c_double), C.byref(best), _p(errs, C.c_double))
    return s.value, R.reshape(3, 3), t, e, best.value, errs


def draw_triples(n, iter_num, seed):
    out = np.empty((iter_num, 3), np.int32)
    lib().orc_draw_triples(C.c_int64(n), C.c_int(iter_num), C.c_uint64(seed), _p(out, C.c_int32))
    return out


def svd3(S):
    S = _f64(S).reshape(9)
    U, V = np.empty(9), np.empty(9)
    sig = np.empty(3)
    lib().orc_svd3(_p(S, C.c_double), _p(U, C.c_double), _p(sig, C.c_double), _p(V, C.c_double))
    return U.reshape(3, 3), sig, V.reshape(3, 3)


def sym_eig3(A):
    A = _f64(A).reshape(9)
    w, Q = np.empty(3), np.empty(9)
    lib().orc_sym_eig3(_p(A, C.c_double), _p(w, C.c_double), _p(Q, C.c_double))
    return w, Q.reshape(3, 3)
