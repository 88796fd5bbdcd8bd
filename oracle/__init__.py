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


_LIB_MT_PATH = _LIB_PATH.parent / "liboracle_mt.so"


def build_mt() -> Path | None:
    """The OpenMP courtesy build (same sources, -fopenmp).  Optional: None when the toolchain has no
    libgomp.  Only bench.py's courtesy multi-thread timing uses it; tests and parity never do."""
    build()
    srcs = list(_HERE.glob("*.cpp")) + list(_HERE.glob("*.h")) + [_HERE / "Makefile"]
    stale = (not _LIB_MT_PATH.exists()) or any(s.stat().st_mtime > _LIB_MT_PATH.stat().st_mtime for s in srcs)
    if stale:
        r = subprocess.run(["make", "-C", str(_HERE), "-j8", "mt"], capture_output=True, text=True)
        if r.returncode != 0:
            return None
    return _LIB_MT_PATH


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
        for name in ("orc_uniform_sampling", "orc_sample_fps", "orc_cot_weights"):
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


# ----------------------------------------------------------------------------- deformation graph
def uniform_sampling(pts, suppress_k=16):
    pts = _f32(pts).reshape(-1, 3)
    out = np.empty(len(pts), np.int32)
    m = lib().orc_uniform_sampling(_p(pts, C.c_float), C.c_int64(len(pts)), C.c_int(suppress_k),
                                   _p(out, C.c_int32))
    return out[:m].copy()


def sample_fps(pts, n_nodes, first=0):
    pts = _f32(pts).reshape(-1, 3)
    out = np.empty(n_nodes, np.int32)
    m = lib().orc_sample_fps(_p(pts, C.c_float), C.c_int64(len(pts)), C.c_int64(n_nodes),
                             C.c_int64(first), _p(out, C.c_int32))
    return out[:m].copy()


def node_graph(nodes, K=8):
    nodes = _f32(nodes).reshape(-1, 3)
    out = np.empty((len(nodes), K + 1), np.int32)
    lib().orc_node_graph(_p(nodes, C.c_float), C.c_int64(len(nodes)), C.c_int(K), _p(out, C.c_int32))
    return out


def node_weights(pts, nodes, k=4):
    pts, nodes = _f32(pts).reshape(-1, 3), _f32(nodes).reshape(-1, 3)
    idx = np.empty((len(pts), k), np.int32)
    w = np.empty((len(pts), k), np.float32)
    lib().orc_node_weights(_p(pts, C.c_float), C.c_int64(len(pts)), _p(nodes, C.c_float),
                           C.c_int64(len(nodes)), C.c_int(k), _p(idx, C.c_int32), _p(w, C.c_float))
    return idx, w


def node_targets(nodes, node_normals, tpts, tnormals, graph, proj_len_err=100.0,
                 proj_dist_err=100.0, smooth_sweeps=2):
    nodes, node_normals = _f32(nodes).reshape(-1, 3), _f32(node_normals).reshape(-1, 3)
    tpts, tnormals = _f32(tpts).reshape(-1, 3), _f32(tnormals).reshape(-1, 3)
    graph = _i32(graph)
    m = len(nodes)
    controls = np.empty((m, 3), np.float64)
    valid = np.empty(m, np.int32)
    lib().orc_node_targets(_p(nodes, C.c_float), _p(node_normals, C.c_float), C.c_int64(m),
                           _p(tpts, C.c_float), _p(tnormals, C.c_float), C.c_int64(len(tpts)),
                           C.c_double(proj_len_err), C.c_double(proj_dist_err),
                           _p(graph, C.c_int32), C.c_int(graph.shape[1]), C.c_int(smooth_sweeps),
                           _p(controls, C.c_double), _p(valid, C.c_int32))
    return controls, valid


def node_targets_threshold(nodes, node_normals, tpts, threshold):
    """Deformation.cpp:155-230: per-node 1-NN target, kept iff d2 <= threshold and cos >= 0.5."""
    nodes, node_normals = _f32(nodes).reshape(-1, 3), _f32(node_normals).reshape(-1, 3)
    tpts = _f32(tpts).reshape(-1, 3)
    m = len(nodes)
    controls = np.empty((m, 3), np.float64)
    valid = np.empty(m, np.int32)
    lib().orc_node_targets_threshold(_p(nodes, C.c_float), _p(node_normals, C.c_float), C.c_int64(m),
                                     _p(tpts, C.c_float), C.c_int64(len(tpts)), C.c_double(threshold),
                                     _p(controls, C.c_double), _p(valid, C.c_int32))
    return controls, valid


def set_literal_comparator(on: bool):
    """Measurement switch of node_targets: the reference's literal, non-transitive priority-queue
    comparator (Deformation.cpp:318-327) instead of the bucketed strict-weak ordering."""
    lib().orc_set_literal_comparator(C.c_int(1 if on else 0))


def gn_assemble_dense(pts, node_ids, widx, w, graph, targets, tweights, prm, x):
    pts = _f32(pts).reshape(-1, 3)
    node_ids, widx, w, graph = _i32(node_ids), _i32(widx), _f32(w), _i32(graph)
    targets = _f64(targets).reshape(-1, 3)
    tweights = _f32(tweights) if tweights is not None else None
    x = _f64(x).reshape(-1)
    m = len(node_ids)
    H = np.empty((12 * m, 12 * m), np.float64)
    g = np.empty(12 * m, np.float64)
    lib().orc_gn_assemble_dense(_p(pts, C.c_float), C.c_int64(len(pts)), _p(node_ids, C.c_int32),
                                C.c_int64(m), _p(widx, C.c_int32), _p(w, C.c_float),
                                _p(graph, C.c_int32), C.c_int(graph.shape[1]),
                                _p(targets, C.c_double), _p(tweights, C.c_float), C.byref(prm),
                                _p(x, C.c_double), _p(H, C.c_double), _p(g, C.c_double))
    return H, g


_lib_mt = None


def lib_mt():
    """liboracle_mt.so or None (see build_mt)."""
    global _lib_mt
    if _lib_mt is None:
        path = build_mt()
        if path is None:
            return None
        _lib_mt = C.CDLL(str(path))
    return _lib_mt


def gn_solve(pts, normals, node_ids, widx, w, graph, prm, x0=None, targets=None, tweights=None,
             tpts=None, tnormals=None, mt=False):
    pts = _f32(pts).reshape(-1, 3)
    normals = _f32(normals).reshape(-1, 3) if normals is not None else None
    node_ids, widx, w, graph = _i32(node_ids), _i32(widx), _f32(w), _i32(graph)
    m, n = len(node_ids), len(pts)
    if x0 is None:
        x = np.tile(np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0], np.float64), (m, 1))
    else:
        x = _f64(x0).reshape(m, 12).copy()
    targets = _f64(targets).reshape(-1, 3) if targets is not None else None
    tweights = _f32(tweights) if tweights is not None else None
    tpts = _f32(tpts).reshape(-1, 3) if tpts is not None else None
    tnormals = _f32(tnormals).reshape(-1, 3) if tnormals is not None else None
    out = np.empty((n, 3), np.float64)
    energy = np.zeros(3, np.float64)
    L = lib()
    if mt:
        L = lib_mt()
        if L is None:
            raise RuntimeError("oracle: the OpenMP courtesy build is not available (no libgomp)")
    L.orc_gn_solve(_p(pts, C.c_float), _p(normals, C.c_float), C.c_int64(n),
                   _p(node_ids, C.c_int32), C.c_int64(m), _p(widx, C.c_int32),
                   _p(w, C.c_float), _p(graph, C.c_int32), C.c_int(graph.shape[1]),
                   _p(targets, C.c_double), _p(tweights, C.c_float), _p(tpts, C.c_float),
                   _p(tnormals, C.c_float), C.c_int64(len(tpts) if tpts is not None else 0),
                   C.byref(prm), _p(x, C.c_double), _p(out, C.c_double),
                   _p(energy, C.c_double))
    return x, out, energy


def gn_correspondences(pts, node_ids, widx, w, x, tpts, want_nn=True):
    """Query points (float32, the device's evaluation order) and exact 1-NN correspondences of one
    Gauss-Newton iteration at node state x (m x 12)."""
    pts, tpts = _f32(pts).reshape(-1, 3), _f32(tpts).reshape(-1, 3)
    node_ids, widx, w = _i32(node_ids), _i32(widx), _f32(w)
    x = _f64(x).reshape(-1)
    n = len(pts)
    q = np.empty((n, 3), np.float32)
    idx = np.empty(n, np.int32) if want_nn else None
    d2 = np.empty(n, np.float32) if want_nn else None
    lib().orc_gn_correspondences(_p(pts, C.c_float), C.c_int64(n), _p(node_ids, C.c_int32),
                                 C.c_int64(len(node_ids)), _p(widx, C.c_int32), _p(w, C.c_float),
                                 _p(x, C.c_double), _p(tpts, C.c_float), C.c_int64(len(tpts)),
                                 _p(q, C.c_float), _p(idx, C.c_int32), _p(d2, C.c_float))
    return q, idx, d2


def deform_vertices(pts, nodes, widx, w, x):
    pts, nodes = _f32(pts).reshape(-1, 3), _f32(nodes).reshape(-1, 3)
    widx, w, x = _i32(widx), _f32(w), _f64(x).reshape(-1)
    out = np.empty((len(pts), 3), np.float64)
    lib().orc_deform_vertices(_p(pts, C.c_float), C.c_int64(len(pts)), _p(nodes, C.c_float),
                              C.c_int64(len(nodes)), _p(widx, C.c_int32), _p(w, C.c_float),
                              _p(x, C.c_double), _p(out, C.c_double))
    return out


def vertex_normals(pts, facets):
    pts = _f64(pts).reshape(-1, 3)
    facets = _i32(facets).reshape(-1, 3)
    out = np.empty_like(pts)
    lib().orc_vertex_normals(_p(pts, C.c_double), C.c_int64(len(pts)), _p(facets, C.c_int32),
                             C.c_int64(len(facets)), _p(out, C.c_double))
    return out


# ----------------------------------------------------------------------------- point sets / alignment
def apply_srt(pts, normals, scale, R, t, inverse=False):
    p = _f64(pts).reshape(-1, 3).copy()
    nr = _f64(normals).reshape(-1, 3).copy() if normals is not None else None
    R, t = _f64(R).reshape(9), _f64(t).reshape(3)
    fn = lib().orc_apply_srt_inverse if inverse else lib().orc_apply_srt
    fn(_p(p, C.c_double), _p(nr, C.c_double), C.c_int64(len(p)), C.c_double(scale),
       _p(R, C.c_double), _p(t, C.c_double))
    return p, nr


def pointset_stats(pts):
    p = _f64(pts).reshape(-1, 3)
    b, mn, mx, pv, ev = np.empty(3), np.empty(3), np.empty(3), np.empty(9), np.empty(3)
    lib().orc_pointset_stats(_p(p, C.c_double), C.c_int64(len(p)), _p(b, C.c_double),
                             _p(mn, C.c_double), _p(mx, C.c_double), _p(pv, C.c_double),
                             _p(ev, C.c_double))
    return b, mn, mx, pv.reshape(3, 3), ev


def project_range(pts, origin, axis, labels=None, label=0):
    p = _f64(pts).reshape(-1, 3)
    lb = _i32(labels) if labels is not None else None
    o, a = _f64(origin).reshape(3), _f64(axis).reshape(3)
    tmin, tmax = C.c_double(), C.c_double()
    imin, imax = C.c_int64(), C.c_int64()
    lib().orc_project_range(_p(p, C.c_double), C.c_int64(len(p)), _p(lb, C.c_int32),
                            C.c_int32(label), _p(o, C.c_double), _p(a, C.c_double),
                            C.byref(tmin), C.byref(tmax), C.byref(imin), C.byref(imax))
    return tmin.value, tmax.value, imin.value, imax.value


def retain_connected(pts, normals, facets):
    p = _f64(pts).reshape(-1, 3).copy()
    nr = _f64(normals).reshape(-1, 3).copy() if normals is not None else None
    f = _i32(facets).reshape(-1).copy()
    n, nf = C.c_int64(len(p)), C.c_int64(len(f) // 3)
    lib().orc_retain_connected(_p(p, C.c_double), _p(nr, C.c_double), C.byref(n),
                               _p(f, C.c_int32), C.byref(nf))
    return (p[:n.value].copy(), nr[:n.value].copy() if nr is not None else None,
            f[:3 * nf.value].reshape(-1, 3).copy())


def remove_ground(pts, normals, facets, dist_thres=0.8, ground_ratio=0.28):
    p = _f64(pts).reshape(-1, 3).copy()
    nr = _f64(normals).reshape(-1, 3).copy() if normals is not None else None
    f = _i32(facets).reshape(-1).copy()
    n, nf = C.c_int64(len(p)), C.c_int64(len(f) // 3)
    gr = np.zeros(3)
    lib().orc_remove_ground(_p(p, C.c_double), _p(nr, C.c_double), C.byref(n), _p(f, C.c_int32),
                            C.byref(nf), C.c_double(dist_thres), C.c_double(ground_ratio),
                            _p(gr, C.c_double))
    return (p[:n.value].copy(), nr[:n.value].copy() if nr is not None else None,
            f[:3 * nf.value].reshape(-1, 3).copy(), gr)


def init_alignment(src, tgt, ground_ray, view_ray):
    s, t = _f64(src).reshape(-1, 3), _f64(tgt).reshape(-1, 3)
    gr, vr = _f64(ground_ray).reshape(3), _f64(view_ray).reshape(3)
    T = np.empty(16)
    sc = C.c_double()
    lib().orc_init_alignment(_p(s, C.c_double), C.c_int64(len(s)), _p(t, C.c_double),
                             C.c_int64(len(t)), _p(gr, C.c_double), _p(vr, C.c_double),
                             _p(T, C.c_double), C.byref(sc))
    return T.reshape(4, 4), sc.value


def local_alignment_core(src, s_labels, tgt, t_labels, sel_labels, label):
    s, t = _f64(src).reshape(-1, 3), _f64(tgt).reshape(-1, 3)
    sl, tl = _i32(s_labels), _i32(t_labels)
    mask = 0
    for k in sel_labels:
        mask |= 1 << int(k)
    T = np.empty(16)
    sc = C.c_double()
    ok = lib().orc_local_alignment_core(_p(s, C.c_double), _p(sl, C.c_int32), C.c_int64(len(s)),
                                        _p(t, C.c_double), _p(tl, C.c_int32), C.c_int64(len(t)),
                                        C.c_uint(mask), C.c_int(label), _p(T, C.c_double),
                                        C.byref(sc))
    return T.reshape(4, 4), sc.value, bool(ok)


def align(src, s_normals, s_labels, tgt, t_normals, t_facets, view_ray, dist_thres=0.8,
          ground_ratio=0.28, remove_ground=1, local_align=1, neck_label=1, ground_ray=None):
    s = _f64(src).reshape(-1, 3).copy()
    sn = _f64(s_normals).reshape(-1, 3).copy() if s_normals is not None else None
    sl = _i32(s_labels)
    t = _f64(tgt).reshape(-1, 3).copy()
    tn = _f64(t_normals).reshape(-1, 3).copy() if t_normals is not None else None
    tf = _i32(t_facets).reshape(-1).copy()
    nt, ntf = C.c_int64(len(t)), C.c_int64(len(tf) // 3)
    vr = _f64(view_ray).reshape(3)
    gr = _f64(ground_ray).reshape(3).copy() if ground_ray is not None else np.zeros(3)
    tl = np.empty(len(t), np.int32)
    lib().orc_align(_p(s, C.c_double), _p(sn, C.c_double), C.c_int64(len(s)), _p(sl, C.c_int32),
                    _p(t, C.c_double), _p(tn, C.c_double), C.byref(nt), _p(tf, C.c_int32),
                    C.byref(ntf), _p(vr, C.c_double), C.c_double(dist_thres),
                    C.c_double(ground_ratio), C.c_int(remove_ground), C.c_int(local_align),
                    C.c_int(neck_label), _p(tl, C.c_int32), _p(gr, C.c_double))
    n, f = nt.value, ntf.value
    return (s, sn, t[:n].copy(), tn[:n].copy() if tn is not None else None,
            tf[:3 * f].reshape(-1, 3).copy(), tl[:n].copy(), gr)


# ----------------------------------------------------------------------------- Model2Depth
def model2depth(pts, facets, cams, znear=0.0, zfar=0.0):
    p = _f32(pts).reshape(-1, 3)
    f = _i32(facets).reshape(-1)
    arr = (Camera * len(cams))(*cams)
    w, h = cams[0].w, cams[0].h
    out = np.empty((len(cams), h, w), np.float32)
    lib().orc_model2depth(_p(p, C.c_float), C.c_int64(len(p)), _p(f, C.c_int32),
                          C.c_int64(len(f) // 3), arr, C.c_int(len(cams)), C.c_float(znear),
                          C.c_float(zfar), _p(out, C.c_float))
    return out


def unproject_depth(inv_depth, cam, min_dsp, max_dsp):
    a = _f32(inv_depth).reshape(cam.h, cam.w)
    out = np.empty((cam.h * cam.w, 3), np.float64)
    valid = np.empty(cam.h * cam.w, np.int32)
    lib().orc_unproject_depth(_p(a, C.c_float), C.byref(cam), C.c_double(min_dsp),
                              C.c_double(max_dsp), _p(out, C.c_double), _p(valid, C.c_int32))
    return out, valid


# ----------------------------------------------------------------------------- pipeline glue
def srt_inliers(p, q, cam1, cam2, scale, R, t, pixel_thr):
    p, q = _f64(p).reshape(-1, 3), _f64(q).reshape(-1, 3)
    R, t = _f64(R).reshape(9), _f64(t).reshape(3)
    flags = np.empty(len(p), np.uint8)
    fn = lib().orc_srt_inliers
    fn.restype = C.c_double
    e = fn(_p(p, C.c_double), _p(q, C.c_double), C.c_int64(len(p)), C.byref(cam1), C.byref(cam2),
           C.c_double(scale), _p(R, C.c_double), _p(t, C.c_double), C.c_double(pixel_thr),
           _p(flags, C.c_uint8))
    return flags, e


def remove_outliers(p, q, cam1, cam2, rounds=3, iter_num=200, pixel_err=55.0, adapt_ratio=0.6, seed=0):
    p, q = _f64(p).reshape(-1, 3), _f64(q).reshape(-1, 3)
    keep = np.empty(len(p), np.int32)
    nk = C.c_int64()
    err, ir = C.c_double(), C.c_double()
    lib().orc_remove_outliers(_p(p, C.c_double), _p(q, C.c_double), C.c_int64(len(p)),
                              C.byref(cam1), C.byref(cam2), C.c_int(rounds), C.c_int(iter_num),
                              C.c_double(pixel_err), C.c_double(adapt_ratio), C.c_uint64(seed),
                              _p(keep, C.c_int32), C.byref(nk), C.byref(err), C.byref(ir))
    return keep[:nk.value].copy(), err.value, ir.value


def remove_outliers_cb(p, q, cam1, cam2, rounds, iter_num, pixel_err, adapt_ratio, draw):
    """remove_outliers with the hypothesis samples of each round supplied by `draw`, a C function
    pointer or ctypes callback of type orc_triples_fn (size, iter_num, out_triples, user)."""
    p, q = _f64(p).reshape(-1, 3), _f64(q).reshape(-1, 3)
    keep = np.empty(len(p), np.int32)
    nk = C.c_int64()
    err, ir = C.c_double(), C.c_double()
    lib().orc_remove_outliers_cb(_p(p, C.c_double), _p(q, C.c_double), C.c_int64(len(p)),
                                 C.byref(cam1), C.byref(cam2), C.c_int(rounds), C.c_int(iter_num),
                                 C.c_double(pixel_err), C.c_double(adapt_ratio), draw, None,
                                 _p(keep, C.c_int32), C.byref(nk), C.byref(err), C.byref(ir))
    return keep[:nk.value].copy(), err.value, ir.value


def depth2model(inv_depth, cam, min_dsp, max_dsp, smooth_threshold):
    a = _f32(inv_depth).reshape(cam.h, cam.w)
    npx = cam.h * cam.w
    xyz = np.empty((npx, 3), np.float64)
    tex = np.empty(npx, np.int32)
    fac = np.empty((2 * max(cam.h - 1, 0) * max(cam.w - 1, 0), 3), np.int32)
    n, nf = C.c_int64(), C.c_int64()
    lib().orc_depth2model(_p(a, C.c_float), C.byref(cam), C.c_double(min_dsp), C.c_double(max_dsp),
                          C.c_double(smooth_threshold), _p(xyz, C.c_double), _p(tex, C.c_int32),
                          _p(fac, C.c_int32), C.byref(n), C.byref(nf))
    return xyz[:n.value].copy(), tex[:n.value].copy(), fac[:nf.value].copy()


def check_consistency(depth, cam, ref_depths, ref_cams, min_dsp, max_dsp, reproj_err):
    d = _f32(depth).reshape(cam.h, cam.w).copy()
    rd = _f32(ref_depths).reshape(len(ref_cams), cam.h, cam.w)
    arr = (Camera * max(1, len(ref_cams)))(*ref_cams)
    lib().orc_check_consistency(_p(d, C.c_float), C.byref(cam), _p(rd, C.c_float), arr,
                                C.c_int(len(ref_cams)), C.c_double(min_dsp), C.c_double(max_dsp),
                                C.c_double(reproj_err))
    return d


# ----------------------------------------------------------------------------- ARAP
def arap(pts, facets, ctrl_ids, ctrl_xyz, iters=5, tol=1e-4, cg_iters=400, cg_tol=1e-6):
    """oracle/arap.cpp orc_arap: returns (deformed n x 3 float64, energy, iterations run)."""
    pts = _f32(pts).reshape(-1, 3)
    facets = _i32(facets).reshape(-1, 3)
    ctrl_ids = _i32(ctrl_ids)
    ctrl_xyz = _f64(ctrl_xyz).reshape(-1, 3)
    out = np.empty((len(pts), 3), np.float64)
    e = C.c_double(0.0)
    it = C.c_int32(0)
    lib().orc_arap(_p(pts, C.c_float), C.c_int64(len(pts)), _p(facets, C.c_int32),
                   C.c_int64(len(facets)), _p(ctrl_ids, C.c_int32), _p(ctrl_xyz, C.c_double),
                   C.c_int64(len(ctrl_ids)), C.c_int(iters), C.c_double(tol), C.c_int(cg_iters),
                   C.c_double(cg_tol), _p(out, C.c_double), C.byref(e), C.byref(it))
    return out, e.value, it.value


def cot_weights(pts, facets):
    """cotangent Laplacian of oracle/arap.cpp as (row_ptr, col, w)."""
    pts = _f32(pts).reshape(-1, 3)
    facets = _i32(facets).reshape(-1, 3)
    n = len(pts)
    cap = 6 * len(facets) + 1
    rp = np.empty(n + 1, np.int32)
    col = np.empty(cap, np.int32)
    w = np.empty(cap, np.float64)
    nnz = lib().orc_cot_weights(_p(pts, C.c_float), C.c_int64(n), _p(facets, C.c_int32),
                                C.c_int64(len(facets)), _p(rp, C.c_int32), _p(col, C.c_int32),
                                _p(w, C.c_double), C.c_int64(cap))
    return rp, col[:nnz], w[:nnz]
