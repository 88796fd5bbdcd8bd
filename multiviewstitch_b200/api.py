"""Host-side mirror of the reference's operator interface for the hot path, over the C ABI.

Class and method names follow the reference (SRTSolver.SetInput / EstimateTransform /
EstimateTransformRansac / ResidualError, PartRecognition.PartRecog, Deformation.Deform,
Model2Depth.SetInput/Run, Alignment.Align ...) so that parity tests read like calls into the
reference.  Everything here is marshalling: numpy arrays in, ctypes calls into libmvs_b200.so,
numpy arrays out.  No arithmetic of the path happens in Python.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._capi import (AlignParams, ArapParams, Camera, GNParams, MvsError, OutlierParams,  # noqa: F401
                    PairJob, PairParams, PairResult, check, lib)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _ptr(a, ct):
    if a is None:
        return None
    return a.ctypes.data_as(C.POINTER(ct))


def _dev(ptr):
    """device pointer given as int (e.g. torch.Tensor.data_ptr())"""
    return C.c_void_p(int(ptr))


def make_camera(K, R, t, w, h) -> Camera:
    """[ref: Camera/Camera.h:9-16 SetK/SetRT/SetW/SetH]"""
    c = Camera()
    c.K[:] = _f64(K).reshape(9).tolist()
    c.R[:] = _f64(R).reshape(9).tolist()
    c.t[:] = _f64(t).reshape(3).tolist()
    c.w, c.h = int(w), int(h)
    return c


def device_count() -> int:
    n = C.c_int(0)
    check(lib().mvs_device_count(C.byref(n)))
    return n.value


class Context:
    """mvs_ctx: device + stream + profiling/launch accounting."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self._h = C.c_void_p()
        check(lib().mvs_ctx_create(C.byref(self._h), C.c_int(device),
                                   C.c_void_p(stream) if stream else None))

    def close(self):
        if self._h:
            lib().mvs_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def sync(self):
        check(lib().mvs_ctx_sync(self._h))

    def prof_enable(self, on=True):
        check(lib().mvs_prof_enable(self._h, C.c_int(1 if on else 0)))

    def prof_reset(self):
        check(lib().mvs_prof_reset(self._h))

    def prof(self) -> dict:
        need = C.c_int64(0)
        check(lib().mvs_prof_names(self._h, None, C.c_int64(0), C.byref(need)))
        buf = C.create_string_buffer(max(1, need.value))
        check(lib().mvs_prof_names(self._h, buf, C.c_int64(len(buf)), C.byref(need)))
        out = {}
        for name in [s for s in buf.value.decode().split("\n") if s]:
            ms, n = C.c_double(0), C.c_int64(0)
            check(lib().mvs_prof_get(self._h, name.encode(), C.byref(ms), C.byref(n)))
            out[name] = (ms.value, n.value)
        return out

    def launch_count(self) -> int:
        n = C.c_int64(0)
        check(lib().mvs_launch_count(self._h, C.byref(n)))
        return n.value

    def launch_count_reset(self):
        check(lib().mvs_launch_count_reset(self._h))

    # ---- multi-GPU
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        check(lib().mvs_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, nranks: int, rank: int, uid: bytes):
        check(lib().mvs_comm_init(self._h, C.c_int(nranks), C.c_int(rank), uid))


def _ctxh(ctx):
    return ctx.handle if ctx is not None else None


# ============================================================================ exact kNN
class NNIndex:
    """Exact kNN index (replaces cv::flann::Index + knnSearch on the path)."""

    def __init__(self, ctx: Context | None = None):
        self._ctx = ctx
        self._h = C.c_void_p()
        check(lib().mvs_nn_index_create(_ctxh(ctx), C.byref(self._h)))

    def close(self):
        if self._h:
            lib().mvs_nn_index_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def build(self, pts):
        pts = np.asarray(pts)
        if pts.dtype == np.float64:
            a = _f64(pts).reshape(-1, 3)
            check(lib().mvs_nn_index_build_host_f64(self._h, _ptr(a, C.c_double), C.c_int64(len(a))))
        else:
            a = _f32(pts).reshape(-1, 3)
            check(lib().mvs_nn_index_build_host(self._h, _ptr(a, C.c_float), C.c_int64(len(a))))
        return self

    def build_dev(self, ptr_xyzw: int, n: int):
        check(lib().mvs_nn_index_build_dev(self._h, _dev(ptr_xyzw), C.c_int64(n)))
        return self

    def knn(self, q, k=1):
        q = np.asarray(q)
        nq = q.reshape(-1, 3).shape[0]
        idx = np.empty((nq, k), np.int32)
        d2 = np.empty((nq, k), np.float32)
        if q.dtype == np.float64:
            a = _f64(q).reshape(-1, 3)
            check(lib().mvs_nn_knn_host_f64(self._h, _ptr(a, C.c_double), C.c_int64(nq), C.c_int(k),
                                            _ptr(idx, C.c_int32), _ptr(d2, C.c_float)))
        else:
            a = _f32(q).reshape(-1, 3)
            check(lib().mvs_nn_knn_host(self._h, _ptr(a, C.c_float), C.c_int64(nq), C.c_int(k),
                                        _ptr(idx, C.c_int32), _ptr(d2, C.c_float)))
        return idx, d2

    def knn_dev(self, q_ptr: int, nq: int, k: int, out_idx_ptr: int, out_d2_ptr: int):
        check(lib().mvs_nn_knn_dev(self._h, _dev(q_ptr), C.c_int64(nq), C.c_int(k),
                                   _dev(out_idx_ptr) if out_idx_ptr else None,
                                   _dev(out_d2_ptr) if out_d2_ptr else None))


def nn_search(q, t, k=1, ctx: Context | None = None):
    q = _f32(q).reshape(-1, 3)
    t = _f32(t).reshape(-1, 3)
    idx = np.empty((len(q), k), np.int32)
    d2 = np.empty((len(q), k), np.float32)
    check(lib().mvs_nn_search(_ctxh(ctx), _ptr(q, C.c_float), C.c_int64(len(q)),
                              _ptr(t, C.c_float), C.c_int64(len(t)), C.c_int(k),
                              _ptr(idx, C.c_int32), _ptr(d2, C.c_float)))
    return idx, d2


class PartRecognition:
    """[ref: PartRecognition/PartRecognition.h:32-49]"""

    def __init__(self, ctx: Context | None = None):
        self._ctx = ctx
        self.points = None
        self.parts = None

    def SetInput(self, points):
        self.points = _f64(points).reshape(-1, 3)

    def SetLabels(self, parts):
        """The reference loads these from Template/part/parts (LoadParts); file I/O is the caller's."""
        self.parts = _i32(parts)

    def GetLabels(self):
        return self.parts

    def PartRecog(self, points_):
        q = _f64(points_).reshape(-1, 3)
        out = np.empty(len(q), np.int32)
        check(lib().mvs_part_recog(_ctxh(self._ctx), _ptr(self.points, C.c_double),
                                   C.c_int64(len(self.points)), _ptr(self.parts, C.c_int32),
                                   _ptr(q, C.c_double), C.c_int64(len(q)), _ptr(out, C.c_int32)))
        return out


# ============================================================================ SRT solver
class SRTSolver:
    """[ref: Solver/SRTSolver.h:8-39]"""

    def __init__(self, iter_num: int = 100, ctx: Context | None = None):
        self._ctx = ctx
        self._h = C.c_void_p()
        check(lib().mvs_srt_create(_ctxh(ctx), C.byref(self._h)))
        check(lib().mvs_srt_set_iteration_num(self._h, C.c_int(iter_num)))
        self._keep = None

    def close(self):
        if self._h:
            lib().mvs_srt_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def SetInput(self, p, q, cam1: Camera | None = None, cam2: Camera | None = None):
        """matches as two n x 3 arrays (first, second). float64 -> double planes on the device;
        float32 n x 4 -> float4 layout."""
        p = np.asarray(p)
        self._n = len(p)
        if p.dtype == np.float32 and p.ndim == 2 and p.shape[1] == 4:
            p4, q4 = _f32(p), _f32(q)
            check(lib().mvs_srt_set_input_f32(self._h, _ptr(p4, C.c_float), _ptr(q4, C.c_float),
                                              C.c_int64(len(p4)),
                                              C.byref(cam1) if cam1 else None,
                                              C.byref(cam2) if cam2 else None))
            return
        pd, qd = _f64(p).reshape(-1, 3), _f64(q).reshape(-1, 3)
        check(lib().mvs_srt_set_input(self._h, _ptr(pd, C.c_double), _ptr(qd, C.c_double),
                                      C.c_int64(len(pd)), C.byref(cam1) if cam1 else None,
                                      C.byref(cam2) if cam2 else None))

    def SetInputDev(self, p_ptr: int, q_ptr: int, n: int, cam1=None, cam2=None):
        self._n = n
        check(lib().mvs_srt_set_input_dev_f32(self._h, _dev(p_ptr), _dev(q_ptr), C.c_int64(n),
                                              C.byref(cam1) if cam1 else None,
                                              C.byref(cam2) if cam2 else None))

    def SetIterationNum(self, it: int):
        check(lib().mvs_srt_set_iteration_num(self._h, C.c_int(it)))

    def SetPrintFlag(self, on: bool):
        check(lib().mvs_srt_set_print(self._h, C.c_int(1 if on else 0)))

    def EstimateTransform(self):
        s = C.c_double()
        R = np.empty(9, np.float64)
        t = np.empty(3, np.float64)
        check(lib().mvs_srt_estimate(self._h, C.byref(s), _ptr(R, C.c_double), _ptr(t, C.c_double)))
        return s.value, R.reshape(3, 3), t

    def EstimateTransformRansac(self, triples=None, seed: int = 0, full=False):
        s = C.c_double()
        R = np.empty(9, np.float64)
        t = np.empty(3, np.float64)
        be, bi = C.c_double(), C.c_int32()
        tr = _i32(triples).reshape(-1, 3) if triples is not None else None
        if tr is not None:
            self.SetIterationNum(len(tr))
        check(lib().mvs_srt_estimate_ransac(self._h, _ptr(tr, C.c_int32), C.c_uint64(seed),
                                            C.byref(s), _ptr(R, C.c_double), _ptr(t, C.c_double),
                                            C.byref(be), C.byref(bi)))
        if full:
            return s.value, R.reshape(3, 3), t, be.value, bi.value
        return s.value, R.reshape(3, 3), t

    def EstimateTransformArray(self, triples=None, seed: int = 0):
        """[ref: Solver/SRTSolver.cpp:256-270 EstimateTransform(double&, double[3][3], double[3])]
        the deprecated array overload: RANSAC path, outputs copied into plain arrays."""
        s = C.c_double()
        R = ((C.c_double * 3) * 3)()
        t = (C.c_double * 3)()
        tr = _i32(triples).reshape(-1, 3) if triples is not None else None
        if tr is not None:
            self.SetIterationNum(len(tr))
        check(lib().mvs_srt_estimate_array(self._h, _ptr(tr, C.c_int32), C.c_uint64(seed),
                                           C.byref(s), R, t))
        return s.value, np.array([[R[i][j] for j in range(3)] for i in range(3)]), np.array(t[:])

    def Inliers(self, scale, R, t, pixel_thr):
        """[ref: Processor/Processor.cpp:210-239] -> (flags uint8, err_all sum, n_inliers)"""
        R, t = _f64(R).reshape(9), _f64(t).reshape(3)
        flags = np.empty(self._n, np.uint8)
        es, ni = C.c_double(), C.c_int64()
        check(lib().mvs_srt_inliers(self._h, C.c_double(scale), _ptr(R, C.c_double),
                                    _ptr(t, C.c_double), C.c_double(pixel_thr),
                                    _ptr(flags, C.c_uint8), C.byref(es), C.byref(ni)))
        return flags, es.value, ni.value

    def ResidualError(self, scale, R, t):
        R, t = _f64(R).reshape(9), _f64(t).reshape(3)
        e = C.c_double()
        check(lib().mvs_srt_residual(self._h, C.c_double(scale), _ptr(R, C.c_double),
                                     _ptr(t, C.c_double), C.byref(e)))
        return e.value


def remove_outliers(p, q, cam1: Camera, cam2: Camera, rounds=3, iter_num=200, pixel_err=55.0,
                    adapt_ratio=0.6, seed=0, ctx: Context | None = None):
    """[ref: Processor/Processor.cpp:177-269 RemoveOutliers] -> (kept match indices, err, inlier_ratio)"""
    p, q = _f64(p).reshape(-1, 3), _f64(q).reshape(-1, 3)
    prm = OutlierParams(rounds, iter_num, pixel_err, adapt_ratio, seed)
    keep = np.empty(len(p), np.int32)
    nk = C.c_int64()
    err, ir = C.c_double(), C.c_double()
    check(lib().mvs_remove_outliers(_ctxh(ctx), _ptr(p, C.c_double), _ptr(q, C.c_double),
                                    C.c_int64(len(p)), C.byref(cam1), C.byref(cam2), C.byref(prm),
                                    _ptr(keep, C.c_int32), C.byref(nk), C.byref(err), C.byref(ir)))
    return keep[:nk.value].copy(), err.value, ir.value


TRIPLES_FN = C.CFUNCTYPE(None, C.c_int64, C.c_int, C.POINTER(C.c_int32), C.c_void_p)


def remove_outliers_samples(p, q, cam1: Camera, cam2: Camera, draw, rounds=3, iter_num=200,
                            pixel_err=55.0, adapt_ratio=0.6, ctx: Context | None = None):
    """RemoveOutliers with the RANSAC samples of every round supplied by `draw`: a TRIPLES_FN callback
    (size, iter_num, out_triples, user) or a Python callable (size, iter_num) -> (iter_num, 3) int
    array.  The reference draws them from rand(); handing that stream over reproduces its result."""
    p, q = _f64(p).reshape(-1, 3), _f64(q).reshape(-1, 3)
    prm = OutlierParams(rounds, iter_num, pixel_err, adapt_ratio, 0)
    keep = np.empty(len(p), np.int32)
    nk = C.c_int64()
    err, ir = C.c_double(), C.c_double()
    if not isinstance(draw, TRIPLES_FN):
        fn = draw

        def _draw(size, iters, out, user):
            np.ctypeslib.as_array(out, shape=(iters, 3))[:] = np.asarray(fn(size, iters), np.int32)
        draw = TRIPLES_FN(_draw)
    check(lib().mvs_remove_outliers_samples(_ctxh(ctx), _ptr(p, C.c_double), _ptr(q, C.c_double),
                                            C.c_int64(len(p)), C.byref(cam1), C.byref(cam2),
                                            C.byref(prm), draw, None, _ptr(keep, C.c_int32),
                                            C.byref(nk), C.byref(err), C.byref(ir)))
    return keep[:nk.value].copy(), err.value, ir.value


def srt_draw_triples(n: int, iter_num: int, seed: int):
    out = np.empty((iter_num, 3), np.int32)
    check(lib().mvs_srt_draw_triples(C.c_int64(n), C.c_int(iter_num), C.c_uint64(seed),
                                     _ptr(out, C.c_int32)))
    return out


# ============================================================================ deformation
def gn_params(**kw) -> GNParams:
    """mvs_gn_params with the library defaults, overridden by keyword."""
    p = GNParams()
    check(lib().mvs_gn_default_params(C.byref(p)))
    for k, v in kw.items():
        if k == "lambda_":
            k = "lam"
        if not hasattr(p, k):
            raise TypeError(f"unknown GN parameter {k}")
        setattr(p, k, v)
    return p


def arap_params(**kw) -> ArapParams:
    """mvs_arap_params with the library defaults (5 iterations, tolerance 1e-4), overridden by keyword."""
    p = ArapParams()
    check(lib().mvs_arap_default_params(C.byref(p)))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise TypeError(f"unknown ARAP parameter {k}")
        setattr(p, k, v)
    return p


class Deformation:
    """[ref: Deformation/Deformation.h:224-252]  Deformation(points, normals, facets);
    UniformSampling(); Deform(tpts, tnormals, projLenErr, projDistErr) / Deform(tpts, tnormals,
    threshold); the CGAL Polyhedron() accessor is replaced by Vertices()/Normals()/Facets()."""

    def __init__(self, points, normals, facets, ctx: Context | None = None):
        self._ctx = ctx
        self._h = C.c_void_p()
        points = np.asarray(points)
        f = _i32(facets).reshape(-1) if facets is not None else None
        nf = 0 if f is None else len(f) // 3
        if points.dtype == np.float32:
            p = _f32(points).reshape(-1, 3)
            nr = _f32(normals).reshape(-1, 3) if normals is not None else None
            check(lib().mvs_deform_create_f32(_ctxh(ctx), _ptr(p, C.c_float), _ptr(nr, C.c_float),
                                              C.c_int64(len(p)), _ptr(f, C.c_int32), C.c_int64(nf),
                                              C.byref(self._h)))
        else:
            p = _f64(points).reshape(-1, 3)
            nr = _f64(normals).reshape(-1, 3) if normals is not None else None
            check(lib().mvs_deform_create(_ctxh(ctx), _ptr(p, C.c_double), _ptr(nr, C.c_double),
                                          C.c_int64(len(p)), _ptr(f, C.c_int32), C.c_int64(nf),
                                          C.byref(self._h)))
        self.n = len(p)
        self.nf = nf
        self._keep = []

    def close(self):
        if self._h:
            lib().mvs_deform_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- nodes / graph
    def UniformSampling(self, suppress_k: int = 16):
        check(lib().mvs_deform_uniform_sampling(self._h, C.c_int(suppress_k)))
        return self.Nodes()

    def SampleFPS(self, n_nodes: int, first_vertex: int = 0):
        check(lib().mvs_deform_sample_fps(self._h, C.c_int64(n_nodes), C.c_int64(first_vertex)))
        return self.Nodes()

    def SetNodes(self, ids):
        ids = _i32(ids)
        check(lib().mvs_deform_set_nodes(self._h, _ptr(ids, C.c_int32), C.c_int64(len(ids))))

    def Nodes(self):
        m = C.c_int64(0)
        check(lib().mvs_deform_get_nodes(self._h, None, C.c_int64(0), C.byref(m)))
        out = np.empty(m.value, np.int32)
        check(lib().mvs_deform_get_nodes(self._h, _ptr(out, C.c_int32), C.c_int64(m.value), C.byref(m)))
        return out

    def BuildGraph(self, node_k: int = 4, graph_k: int = 8):
        check(lib().mvs_deform_build_graph(self._h, C.c_int(node_k), C.c_int(graph_k)))
        self._graph_k = graph_k

    def Weights(self):
        idx = np.empty((self.n, 4), np.int32)
        w = np.empty((self.n, 4), np.float32)
        check(lib().mvs_deform_get_weights(self._h, _ptr(idx, C.c_int32), _ptr(w, C.c_float)))
        return idx, w

    def KNearestNeighbor(self, K: int = 8):
        """[ref: Deformation/Deformation.cpp:108-153] node graph, (K+1) nearest incl. self."""
        self.BuildGraph(4, K)
        out = np.empty((len(self.Nodes()), K + 1), np.int32)
        check(lib().mvs_deform_get_graph(self._h, _ptr(out, C.c_int32)))
        return out

    # ---- targets
    def SetTarget(self, tpts, tnormals=None):
        tpts = np.asarray(tpts)
        if tpts.dtype == np.float32:
            p = _f32(tpts).reshape(-1, 3)
            nr = _f32(tnormals).reshape(-1, 3) if tnormals is not None else None
            check(lib().mvs_deform_set_target_f32(self._h, _ptr(p, C.c_float), _ptr(nr, C.c_float),
                                                  C.c_int64(len(p))))
        else:
            p = _f64(tpts).reshape(-1, 3)
            nr = _f64(tnormals).reshape(-1, 3) if tnormals is not None else None
            check(lib().mvs_deform_set_target(self._h, _ptr(p, C.c_double), _ptr(nr, C.c_double),
                                              C.c_int64(len(p))))

    def SetTargetDepth(self, inv_depth, cam: "Camera", min_dsp: float, max_dsp: float) -> int:
        """Target = the valid pixels of an inverse-depth image, unprojected and compacted in pixel
        order on the device (unproject_depth + boolean compaction + SetTarget without the host
        round trip; identical points and order).  Returns the number of target points."""
        a = _f32(inv_depth).reshape(cam.h, cam.w)
        nt = C.c_int64(0)
        check(lib().mvs_deform_set_target_depth(self._h, _ptr(a, C.c_float), C.byref(cam),
                                                C.c_double(min_dsp), C.c_double(max_dsp),
                                                C.byref(nt)))
        return int(nt.value)

    def SetTargetDev(self, tpts_ptr: int, tnormals_ptr: int | None, nt: int):
        check(lib().mvs_deform_set_target_dev_f32(self._h, _dev(tpts_ptr),
                                                  _dev(tnormals_ptr) if tnormals_ptr else None,
                                                  C.c_int64(nt)))

    def SetTargets(self, targets, weights=None):
        t = _f64(targets).reshape(-1, 3)
        w = _f32(weights) if weights is not None else None
        check(lib().mvs_deform_set_targets(self._h, _ptr(t, C.c_double), _ptr(w, C.c_float),
                                           C.c_int64(len(t))))

    def NodeTargets(self, projLenErr=100.0, projDistErr=100.0, smooth_sweeps=2):
        m = len(self.Nodes())
        c = np.empty((m, 3), np.float64)
        v = np.empty(m, np.int32)
        check(lib().mvs_deform_node_targets(self._h, C.c_double(projLenErr), C.c_double(projDistErr),
                                            C.c_int(smooth_sweeps), _ptr(c, C.c_double),
                                            _ptr(v, C.c_int32)))
        return c, v

    def NodeTargetsThreshold(self, threshold: float):
        """Correspondence stage of Deform(tpts, tnormals, threshold) [ref: Deformation.cpp:186-203]."""
        m = len(self.Nodes())
        c = np.empty((m, 3), np.float64)
        v = np.empty(m, np.int32)
        check(lib().mvs_deform_node_targets_threshold(self._h, C.c_double(threshold),
                                                      _ptr(c, C.c_double), _ptr(v, C.c_int32)))
        return c, v

    # ---- Deform
    def Deform(self, tpts, tnormals, *args, params: GNParams | None = None):
        """Deform(tpts, tnormals, threshold) [ref: Deformation.cpp:155-230] or
        Deform(tpts, tnormals, projLenErr, projDistErr) [ref: :232-402]."""
        if tpts is not None:
            self.SetTarget(tpts, tnormals)
        pp = C.byref(params) if params is not None else None
        if len(args) == 1:
            check(lib().mvs_deform_run_threshold(self._h, C.c_double(args[0]), pp))
        elif len(args) == 2:
            check(lib().mvs_deform_run(self._h, C.c_double(args[0]), C.c_double(args[1]), pp))
        else:
            raise TypeError("Deform(tpts, tnormals, threshold) or Deform(tpts, tnormals, "
                            "projLenErr, projDistErr)")

    def DeformARAP(self, tpts, tnormals, projLenErr=100.0, projDistErr=100.0,
                   params: ArapParams | None = None):
        """Deform(tpts, tnormals, projLenErr, projDistErr) with the as-rigid-as-possible solver the
        reference runs through CGAL [ref: Deformation.cpp:232-402, CGAL calls :256-260, :383-400]."""
        if tpts is not None:
            self.SetTarget(tpts, tnormals)
        pp = C.byref(params) if params is not None else None
        check(lib().mvs_deform_run_arap(self._h, C.c_double(projLenErr), C.c_double(projDistErr), pp))

    def ARAP(self, control_xyz, params: ArapParams | None = None):
        """ARAP solve with the sampled nodes as control vertices moved to control_xyz (m x 3);
        returns (energy, iterations run).  Vertices()/Normals() hold the result."""
        c = _f64(control_xyz).reshape(-1, 3)
        if len(c) != len(self.Nodes()):
            raise ValueError("control_xyz must have one row per node")
        e = C.c_double(0.0)
        it = C.c_int32(0)
        pp = C.byref(params) if params is not None else None
        check(lib().mvs_deform_arap(self._h, _ptr(c, C.c_double), pp, C.byref(e), C.byref(it)))
        return e.value, it.value

    def GNBegin(self, params: GNParams):
        check(lib().mvs_deform_gn_begin(self._h, C.byref(params)))

    def GNStep(self):
        check(lib().mvs_deform_gn_step(self._h))

    def GNEnd(self):
        check(lib().mvs_deform_gn_end(self._h))

    def Solve(self, params: GNParams):
        """gn_begin + gn_iters x gn_step + gn_end in one library call (no host round trips)."""
        check(lib().mvs_deform_solve(self._h, C.byref(params)))

    def Correspondences(self):
        """Target index matched to every vertex (original order) by the last GN step."""
        out = np.empty(self.n, np.int32)
        check(lib().mvs_deform_get_correspondences(self._h, _ptr(out, C.c_int32)))
        return out

    def Energy(self):
        e = np.zeros(3, np.float64)
        check(lib().mvs_deform_get_energy(self._h, _ptr(e, C.c_double)))
        return e

    def NodeState(self):
        x = np.empty((len(self.Nodes()), 12), np.float32)
        check(lib().mvs_deform_get_node_state(self._h, _ptr(x, C.c_float)))
        return x

    def SetNodeState(self, x):
        x = _f32(x).reshape(-1, 12)
        check(lib().mvs_deform_set_node_state(self._h, _ptr(x, C.c_float)))

    def System(self):
        m = len(self.Nodes())
        nnz = C.c_int64(0)
        check(lib().mvs_deform_get_system(self._h, None, None, None, None, None, C.byref(nnz)))
        rp = np.empty(m + 1, np.int32)
        col = np.empty(nnz.value, np.int32)
        B = np.empty((nnz.value, 4, 4), np.float32)
        rot = np.empty((m, 9, 9), np.float32)
        rhs = np.empty((m, 12), np.float32)
        check(lib().mvs_deform_get_system(self._h, _ptr(rp, C.c_int32), _ptr(col, C.c_int32),
                                          _ptr(B, C.c_float), _ptr(rot, C.c_float),
                                          _ptr(rhs, C.c_float), C.byref(nnz)))
        return rp, col, B, rot, rhs

    def Stats(self) -> dict:
        out = (C.c_int64 * 6)()
        check(lib().mvs_deform_get_stats(self._h, out))
        return dict(zip(("vertices", "nodes", "pattern_blocks", "tuple_runs", "facets",
                         "target_points"), [int(v) for v in out]))

    def SetShardInterleaved(self, rank: int, nranks: int, chunks_per_rank: int = 8):
        check(lib().mvs_deform_set_shard_interleaved(self._h, C.c_int(rank), C.c_int(nranks),
                                                     C.c_int(chunks_per_rank)))

    def SetShardParts(self, parts: int):
        check(lib().mvs_deform_set_shard_parts(self._h, C.c_int(parts)))

    def SetShard(self, v_begin: int, v_end: int):
        check(lib().mvs_deform_set_shard(self._h, C.c_int64(v_begin), C.c_int64(v_end)))

    # ---- results
    def Vertices(self):
        out = np.empty((self.n, 3), np.float64)
        check(lib().mvs_deform_get_vertices(self._h, _ptr(out, C.c_double)))
        return out

    def Normals(self):
        out = np.empty((self.n, 3), np.float64)
        check(lib().mvs_deform_get_normals(self._h, _ptr(out, C.c_double)))
        return out

    def Facets(self):
        out = np.empty((self.nf, 3), np.int32)
        check(lib().mvs_deform_get_facets(self._h, _ptr(out, C.c_int32), None))
        return out


def exportOBJ(filename: str, deform: Deformation):
    """[ref: Deformation/Deformation.h:174-221]"""
    check(lib().mvs_deform_export_obj(deform._h, filename.encode()))


# ============================================================================ point sets / alignment
def apply_srt(pts, normals, scale, R, t, inverse=False, ctx: Context | None = None):
    """v = s R p + t, n = R n [ref: Alignment/Alignment.cpp:31-34, Processor.cpp:1025-1026];
    inverse: p_k = (1/s) R^T (p - t), n_k = R^T n [ref: Processor/Processor.cpp:1177-1189]."""
    p = _f64(pts).reshape(-1, 3).copy()
    nr = _f64(normals).reshape(-1, 3).copy() if normals is not None else None
    R, t = _f64(R).reshape(9), _f64(t).reshape(3)
    fn = lib().mvs_apply_srt_inverse if inverse else lib().mvs_apply_srt
    check(fn(_ctxh(ctx), _ptr(p, C.c_double), _ptr(nr, C.c_double), C.c_int64(len(p)),
             C.c_double(scale), _ptr(R, C.c_double), _ptr(t, C.c_double)))
    return p, nr


def apply_srt_dev(in_ptr: int, out_ptr: int, n: int, scale, R, t, ctx: Context | None = None):
    R, t = _f64(R).reshape(9), _f64(t).reshape(3)
    check(lib().mvs_apply_srt_dev_f32(_ctxh(ctx), _dev(in_ptr), _dev(out_ptr), C.c_int64(n),
                                      C.c_double(scale), _ptr(R, C.c_double), _ptr(t, C.c_double)))


class PointSetUtils:
    """[ref: SetUtils/PointSetUtils.h / .cpp:3-62]"""

    def __init__(self, ctx: Context | None = None):
        self._ctx = ctx

    def SetInput(self, points):
        self.points = _f64(points).reshape(-1, 3)
        b, mn, mx, pv, ev = (np.empty(3), np.empty(3), np.empty(3), np.empty(9), np.empty(3))
        check(lib().mvs_pointset_stats(_ctxh(self._ctx), _ptr(self.points, C.c_double),
                                       C.c_int64(len(self.points)), _ptr(b, C.c_double),
                                       _ptr(mn, C.c_double), _ptr(mx, C.c_double),
                                       _ptr(pv, C.c_double), _ptr(ev, C.c_double)))
        self._b, self._mn, self._mx, self._pv, self._ev = b, mn, mx, pv.reshape(3, 3), ev

    def GetBarycenter(self):
        return self._b

    def GetBoundingBox(self):
        return self._mn, self._mx

    def CalcPivots(self, pivot_num: int = 3):
        return self._pv[:, :pivot_num].copy()

    def Eigenvalues(self):
        return self._ev


def project_range(pts, origin, axis, labels=None, label=0, ctx: Context | None = None):
    p = _f64(pts).reshape(-1, 3)
    lb = _i32(labels) if labels is not None else None
    o, a = _f64(origin).reshape(3), _f64(axis).reshape(3)
    tmin, tmax = C.c_double(), C.c_double()
    imin, imax = C.c_int64(), C.c_int64()
    check(lib().mvs_project_range(_ctxh(ctx), _ptr(p, C.c_double), C.c_int64(len(p)),
                                  _ptr(lb, C.c_int32), C.c_int32(label), _ptr(o, C.c_double),
                                  _ptr(a, C.c_double), C.byref(tmin), C.byref(tmax),
                                  C.byref(imin), C.byref(imax)))
    return tmin.value, tmax.value, imin.value, imax.value


def align_params(**kw) -> AlignParams:
    p = AlignParams()
    check(lib().mvs_align_default_params(C.byref(p)))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise TypeError(f"unknown alignment parameter {k}")
        setattr(p, k, v)
    return p


class Alignment:
    """[ref: Alignment/Alignment.h:21-36].  The reference mutates its std::vector arguments in
    place; here the updated arrays are returned."""

    def __init__(self, ctx: Context | None = None, params: AlignParams | None = None):
        self._ctx = ctx
        self.params = params or align_params()

    def Align(self, src, s_normals, s_labels, tgt, t_normals, t_facets, viewRay,
              ground_ray=None):
        """Returns (src, s_normals, tgt, t_normals, t_facets, t_labels, groundRay)."""
        s = _f64(src).reshape(-1, 3).copy()
        sn = _f64(s_normals).reshape(-1, 3).copy() if s_normals is not None else None
        sl = _i32(s_labels)
        t = _f64(tgt).reshape(-1, 3).copy()
        tn = _f64(t_normals).reshape(-1, 3).copy() if t_normals is not None else None
        tf = _i32(t_facets).reshape(-1).copy()
        nt, ntf = C.c_int64(len(t)), C.c_int64(len(tf) // 3)
        vr = _f64(viewRay).reshape(3)
        gr = _f64(ground_ray).reshape(3).copy() if ground_ray is not None else np.zeros(3)
        tl = np.empty(len(t), np.int32)
        check(lib().mvs_align(_ctxh(self._ctx), _ptr(s, C.c_double), _ptr(sn, C.c_double),
                              C.c_int64(len(s)), _ptr(sl, C.c_int32), _ptr(t, C.c_double),
                              _ptr(tn, C.c_double), C.byref(nt), _ptr(tf, C.c_int32),
                              C.byref(ntf), _ptr(vr, C.c_double), C.byref(self.params),
                              _ptr(tl, C.c_int32), _ptr(gr, C.c_double)))
        n, f = nt.value, ntf.value
        return (s, sn, t[:n].copy(), tn[:n].copy() if tn is not None else None,
                tf[:3 * f].reshape(-1, 3).copy(), tl[:n].copy(), gr)

    def RemoveGround(self, points, normals, facets):
        """[ref: Alignment/Alignment.cpp:79-233] -> (points, normals, facets, groundRay)"""
        p = _f64(points).reshape(-1, 3).copy()
        nr = _f64(normals).reshape(-1, 3).copy() if normals is not None else None
        f = _i32(facets).reshape(-1).copy()
        n, nf = C.c_int64(len(p)), C.c_int64(len(f) // 3)
        gr = np.zeros(3)
        check(lib().mvs_remove_ground(_ctxh(self._ctx), _ptr(p, C.c_double), _ptr(nr, C.c_double),
                                      C.byref(n), _ptr(f, C.c_int32), C.byref(nf),
                                      C.byref(self.params), _ptr(gr, C.c_double)))
        return (p[:n.value].copy(), nr[:n.value].copy() if nr is not None else None,
                f[:3 * nf.value].reshape(-1, 3).copy(), gr)

    def RetainConnectRegion(self, points, normals, facets):
        """[ref: Alignment/Alignment.cpp:618-654]"""
        p = _f64(points).reshape(-1, 3).copy()
        nr = _f64(normals).reshape(-1, 3).copy() if normals is not None else None
        f = _i32(facets).reshape(-1).copy()
        n, nf = C.c_int64(len(p)), C.c_int64(len(f) // 3)
        check(lib().mvs_retain_connected(_ptr(p, C.c_double), _ptr(nr, C.c_double), C.byref(n),
                                         _ptr(f, C.c_int32), C.byref(nf)))
        return (p[:n.value].copy(), nr[:n.value].copy() if nr is not None else None,
                f[:3 * nf.value].reshape(-1, 3).copy())


# ============================================================================ Model2Depth
class Model2Depth:
    """[ref: Model2Depth/Model2Depth.h:17-54] SetInput(points, normals, facets, cameras, path) then
    Run(): renders mesh[seq] into every camera of sequence seq and writes
    <path[seq]>DATA/Render/_depth<frame>.raw (float32 inverse depth).  Run() returns the images."""

    znear = 0.01   # Model2Depth.cpp:19
    zfar = 2000.0  # Model2Depth.cpp:20

    def __init__(self, ctx: Context | None = None):
        self._ctx = ctx
        self.points, self.facets, self.cameras, self.path = [], [], [], []

    def SetInput(self, points, normals, facets, cameras, path=None):
        self.points = [_f32(p).reshape(-1, 3) for p in points]   # Model2Depth.h:35-39 float cast
        self.facets = [_i32(f).reshape(-1) for f in facets]
        self.cameras = [list(c) for c in cameras]
        self.path = list(path) if path is not None else [None] * len(self.points)

    def render(self, seq: int):
        cams = self.cameras[seq]
        arr = (Camera * len(cams))(*cams)
        w, h = cams[0].w, cams[0].h
        out = np.empty((len(cams), h, w), np.float32)
        p, f = self.points[seq], self.facets[seq]
        check(lib().mvs_model2depth(_ctxh(self._ctx), _ptr(p, C.c_float), C.c_int64(len(p)),
                                    _ptr(f, C.c_int32), C.c_int64(len(f) // 3), arr,
                                    C.c_int(len(cams)), C.c_float(self.znear), C.c_float(self.zfar),
                                    _ptr(out, C.c_float)))
        return out

    def Run(self):
        import os
        images = []
        for seq in range(len(self.points)):
            out = self.render(seq)
            images.append(out)
            if self.path[seq] is not None:
                d = os.path.join(self.path[seq], "DATA", "Render")
                os.makedirs(d, exist_ok=True)
                for frm in range(out.shape[0]):
                    save_depth_raw(os.path.join(d, f"_depth{frm}.raw"), out[frm])
        return images


def model2depth_dev(pts_ptr: int, n: int, facets_ptr: int, nf: int, cams, out_ptr: int,
                    znear=0.0, zfar=0.0, ctx: Context | None = None):
    arr = (Camera * len(cams))(*cams)
    check(lib().mvs_model2depth_dev(_ctxh(ctx), _dev(pts_ptr), C.c_int64(n), _dev(facets_ptr),
                                    C.c_int64(nf), arr, C.c_int(len(cams)), C.c_float(znear),
                                    C.c_float(zfar), _dev(out_ptr)))


def save_depth_raw(filename: str, inv_depth):
    """[ref: Common/Utils.h:177-185 SaveDepth]"""
    a = _f32(inv_depth)
    h, w = a.shape
    check(lib().mvs_save_depth_raw(filename.encode(), _ptr(a, C.c_float), C.c_int(w), C.c_int(h)))


def load_depth_raw(filename: str, w: int, h: int):
    """[ref: Common/Utils.h:166-176 LoadDepth]"""
    a = np.empty((h, w), np.float32)
    check(lib().mvs_load_depth_raw(filename.encode(), _ptr(a, C.c_float), C.c_int(w), C.c_int(h)))
    return a


def depth2model(inv_depth, cam: Camera, min_dsp: float, max_dsp: float, smooth_threshold: float,
                ctx: Context | None = None):
    """[ref: Depth2Model/Depth2Model.cpp:7-82 SaveModel] -> (point3d, texIndex, facets)"""
    a = _f32(inv_depth).reshape(cam.h, cam.w)
    n, nf = C.c_int64(), C.c_int64()
    args = (_ctxh(ctx), _ptr(a, C.c_float), C.byref(cam), C.c_double(min_dsp), C.c_double(max_dsp),
            C.c_double(smooth_threshold))
    check(lib().mvs_depth2model(*args, None, None, None, C.byref(n), C.byref(nf)))
    xyz = np.empty((n.value, 3), np.float64)
    tex = np.empty(n.value, np.int32)
    fac = np.empty((nf.value, 3), np.int32)
    check(lib().mvs_depth2model(*args, _ptr(xyz, C.c_double), _ptr(tex, C.c_int32),
                                _ptr(fac, C.c_int32), C.byref(n), C.byref(nf)))
    return xyz, tex, fac


def check_consistency(depth, cam: Camera, ref_depths, ref_cams, min_dsp, max_dsp, reproj_err,
                      ctx: Context | None = None):
    """[ref: Processor/Processor.cpp:66-115 CheckConsistencyCore] -> filtered inverse-depth map"""
    d = _f32(depth).reshape(cam.h, cam.w).copy()
    rd = _f32(ref_depths).reshape(len(ref_cams), cam.h, cam.w)
    arr = (Camera * max(1, len(ref_cams)))(*ref_cams)
    check(lib().mvs_check_consistency(_ctxh(ctx), _ptr(d, C.c_float), C.byref(cam),
                                      _ptr(rd, C.c_float), arr, C.c_int(len(ref_cams)),
                                      C.c_double(min_dsp), C.c_double(max_dsp), C.c_double(reproj_err)))
    return d


def unproject_pixels(inv_depth, cam: Camera, uv, min_dsp: float, max_dsp: float,
                     ctx: Context | None = None):
    """[ref: Camera/Camera.cpp:40-44,61-67] world points of the pixels uv (n x 2: column, row) ->
    (points n x 3, valid n); bit-identical to the matching rows of unproject_depth."""
    a = _f32(inv_depth).reshape(cam.h, cam.w)
    px = np.ascontiguousarray(uv, np.int32).reshape(-1, 2)
    out = np.empty((len(px), 3), np.float64)
    valid = np.empty(len(px), np.int32)
    check(lib().mvs_unproject_pixels(_ctxh(ctx), _ptr(a, C.c_float), C.byref(cam),
                                     C.c_double(min_dsp), C.c_double(max_dsp),
                                     _ptr(px, C.c_int32), C.c_int64(len(px)),
                                     _ptr(out, C.c_double), _ptr(valid, C.c_int32)))
    return out, valid


def unproject_depth(inv_depth, cam: Camera, min_dsp: float, max_dsp: float,
                    ctx: Context | None = None):
    """[ref: Image3D/Image3D.cpp:92-106 SolveUnProjectionD] -> (points h*w x 3, valid h*w)"""
    a = _f32(inv_depth).reshape(cam.h, cam.w)
    out = np.empty((cam.h * cam.w, 3), np.float64)
    valid = np.empty(cam.h * cam.w, np.int32)
    check(lib().mvs_unproject_depth(_ctxh(ctx), _ptr(a, C.c_float), C.byref(cam),
                                    C.c_double(min_dsp), C.c_double(max_dsp),
                                    _ptr(out, C.c_double), _ptr(valid, C.c_int32)))
    return out, valid


# ============================================================================ view pairs
def pair_params(**kw) -> PairParams:
    """mvs_pair_params with the library defaults (config.txt values), overridden by keyword;
    `gn` and `outliers` take the nested structures, gn_iters / pcg_iters / iter_num / pixel_err /
    adapt_ratio are shortcuts into them."""
    p = PairParams()
    check(lib().mvs_pair_default_params(C.byref(p)))
    for k, v in kw.items():
        if k in ("gn_iters", "pcg_iters"):
            setattr(p.gn, k, v)
        elif k in ("iter_num", "pixel_err", "adapt_ratio", "rounds"):
            setattr(p.outliers, k, v)
        else:
            setattr(p, k, v)
    return p


class PairBatch:
    """A list of independent view pairs for mvs_stitch_pairs [ref: Processor/Processor.cpp:629-826].
    add() keeps the host arrays alive; run() is ONE library call for all pairs."""

    def __init__(self, ctx: Context | None = None, want_outputs: bool = True):
        self._ctx = ctx
        self._jobs = []
        self._keep = []
        self._out = []
        self.want_outputs = want_outputs

    def add(self, depth_a, cam_a: Camera, depth_b, cam_b: Camera, matches, seed: int = 0):
        da = _f32(depth_a).reshape(-1)
        db = _f32(depth_b).reshape(-1)
        mt = _i32(matches).reshape(-1, 4)
        npx = cam_a.w * cam_a.h
        j = PairJob()
        j.depth_a, j.cam_a = _ptr(da, C.c_float), C.pointer(cam_a)
        j.depth_b, j.cam_b = _ptr(db, C.c_float), C.pointer(cam_b)
        j.matches, j.n_matches, j.seed = _ptr(mt, C.c_int32), len(mt), seed
        o = None
        if self.want_outputs:
            o = dict(kept=np.empty(len(mt), np.int32), template=np.empty((npx, 3), np.float64),
                     facets=np.empty((2 * (cam_a.w - 1) * (cam_a.h - 1), 3), np.int32),
                     deformed=np.empty((npx, 3), np.float64), render=np.empty(npx, np.float32))
            j.kept_idx = _ptr(o["kept"], C.c_int32)
            j.template_xyz = _ptr(o["template"], C.c_double)
            j.facets = _ptr(o["facets"], C.c_int32)
            j.deformed_xyz = _ptr(o["deformed"], C.c_double)
            j.render = _ptr(o["render"], C.c_float)
        self._jobs.append(j)
        self._keep.append((da, db, mt, cam_a, cam_b))
        self._out.append(o)

    def __len__(self):
        return len(self._jobs)

    def run(self, params: PairParams | None = None):
        n = len(self._jobs)
        jobs = (PairJob * n)(*self._jobs)
        res = (PairResult * n)()
        prm = params if params is not None else pair_params()
        check(lib().mvs_stitch_pairs(_ctxh(self._ctx), jobs, C.c_int64(n), C.byref(prm), res))
        outs = []
        for r, o, k in zip(res, self._out, self._keep):
            d = dict(scale=r.scale, R=np.array(r.R[:]).reshape(3, 3), t=np.array(r.t[:]), err=r.err,
                     residual=r.residual, n_kept=r.n_kept, n_template=r.n_template,
                     n_facets=r.n_facets, n_target=r.n_target)
            if o is not None:
                cam = k[3]
                d.update(kept=o["kept"][:r.n_kept], template=o["template"][:r.n_template],
                         facets=o["facets"][:r.n_facets], deformed=o["deformed"][:r.n_template],
                         render=o["render"].reshape(cam.h, cam.w))
            outs.append(d)
        return outs
