"""oracle.ref — TEST INFRASTRUCTURE ONLY.

ctypes front-end of oracle/_ref/libmvsref.so: the reference's own first-party sources
(Solver/SRTSolver.cpp, Camera/Camera.cpp, SetUtils/PointSetUtils.cpp, SetUtils/UnionSetUtils.cpp,
Depth2Model/Depth2Model.cpp, Common/Utils.h) compiled UNMODIFIED from /root/reference by
`make -C oracle ref` against the stand-in headers in oracle/compat/ (see oracle/ref_shim.cpp), and
Alignment/Alignment.cpp, PartRecognition/PartRecognition.cpp, PlyObj/PlyObj.cpp,
Parameter/ParamParser.cpp, Deformation/Deformation.cpp, Processor/Processor.cpp compiled from their source text with their
headers passed through sed (MSVC-only forms; see oracle/ref_shim_align.cpp, ref_shim_deform.cpp,
ref_shim_processor.cpp and
oracle/Makefile) — Deformation against a CGAL stand-in that records the constraints handed to the
solver (oracle/compat/CGAL/mvs_cgal_standin.h).
Used to pin the restatement in oracle/*.cpp (tests/test_ref_parity.py), to generate the committed
vectors tests/golden/ref_*.npz (tests/golden/make_ref_golden.py) and as the `kind: "reference"` CPU
leg of bench.py's SRT line.  The product never loads it.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from . import Camera, _f32, _f64, _i32, _p

_HERE = Path(__file__).resolve().parent
REF_ROOT = Path("/root/reference/MultiViewStitch")
_LIB_PATH = _HERE / "_ref" / "libmvsref.so"
_lib = None


def build() -> Path | None:
    """Compile the reference sources where they lie (only possible where /root/reference exists —
    this container; the GPU box uses the prebuilt oracle/_ref/libmvsref.so that travels with the
    snapshot)."""
    if not REF_ROOT.exists():
        return _LIB_PATH if _LIB_PATH.exists() else None
    r = subprocess.run(["make", "-C", str(_HERE), "-j8", "ref"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle/_ref build failed:\n" + r.stdout + r.stderr)
    return _LIB_PATH


def available() -> bool:
    return _LIB_PATH.exists() or REF_ROOT.exists()


def lib():
    global _lib
    if _lib is None:
        if build() is None:
            raise RuntimeError("oracle/_ref/libmvsref.so is not built and /root/reference is absent")
        _lib = C.CDLL(str(_LIB_PATH))
        _lib.ref_srt_residual.restype = C.c_double
        _lib.ref_connected_keep.restype = C.c_int
    return _lib


def _cam(c):
    return C.byref(c) if c is not None else None


def srt_estimate(p, q, cam1: Camera | None = None, cam2: Camera | None = None):
    p, q = _f64(p).reshape(-1, 3), _f64(q).reshape(-1, 3)
    s = C.c_double()
    R, t = np.empty(9), np.empty(3)
    lib().ref_srt_estimate(_p(p, C.c_double), _p(q, C.c_double), C.c_int64(len(p)), _cam(cam1),
                           _cam(cam2), C.byref(s), _p(R, C.c_double), _p(t, C.c_double))
    return s.value, R.reshape(3, 3), t


def srt_residual(p, q, cam1: Camera, cam2: Camera, scale, R, t):
    p, q = _f64(p).reshape(-1, 3), _f64(q).reshape(-1, 3)
    R, t = _f64(R).reshape(9), _f64(t).reshape(3)
    return lib().ref_srt_residual(_p(p, C.c_double), _p(q, C.c_double), C.c_int64(len(p)),
                                  _cam(cam1), _cam(cam2), C.c_double(scale), _p(R, C.c_double),
                                  _p(t, C.c_double))


def srt_triples(n: int, iter_num: int, seed: int):
    tr = np.empty((iter_num, 3), np.int32)
    lib().ref_srt_triples(C.c_int64(n), C.c_int(iter_num), C.c_uint(seed), _p(tr, C.c_int32))
    return tr


def srt_ransac(p, q, cam1: Camera, cam2: Camera, iter_num: int, seed: int, array: bool = False):
    p, q = _f64(p).reshape(-1, 3), _f64(q).reshape(-1, 3)
    s = C.c_double()
    R, t = np.empty(9), np.empty(3)
    lib().ref_srt_ransac(_p(p, C.c_double), _p(q, C.c_double), C.c_int64(len(p)), _cam(cam1),
                         _cam(cam2), C.c_int(iter_num), C.c_uint(seed), C.c_int(int(array)),
                         C.byref(s), _p(R, C.c_double), _p(t, C.c_double))
    return s.value, R.reshape(3, 3), t


def project(cam: Camera, pts):
    pts = _f64(pts).reshape(-1, 3)
    uv = np.empty((len(pts), 2), np.int32)
    lib().ref_project(_cam(cam), _p(pts, C.c_double), C.c_int64(len(pts)), _p(uv, C.c_int32))
    return uv


def unproject(cam: Camera, uv, depth):
    uv = _i32(uv).reshape(-1, 2)
    depth = _f64(depth).reshape(-1)
    out = np.empty((len(uv), 3))
    lib().ref_unproject(_cam(cam), _p(uv, C.c_int32), _p(depth, C.c_double), C.c_int64(len(uv)),
                        _p(out, C.c_double))
    return out


def gl_matrices(cam: Camera, znear: float, zfar: float):
    mv, fr, pr = np.empty(16, np.float32), np.empty(4, np.float32), np.empty(16, np.float32)
    lib().ref_gl_matrices(_cam(cam), C.c_float(znear), C.c_float(zfar), _p(mv, C.c_float),
                          _p(fr, C.c_float), _p(pr, C.c_float))
    return mv.reshape(4, 4), fr, pr.reshape(4, 4)


def pointset(pts):
    pts = _f64(pts).reshape(-1, 3)
    b, c, P = np.empty(3), np.empty(3), np.empty(9)
    lib().ref_pointset(_p(pts, C.c_double), C.c_int64(len(pts)), _p(b, C.c_double),
                       _p(c, C.c_double), _p(P, C.c_double))
    return b, c, P.reshape(3, 3)


def connected_keep(n: int, facets):
    facets = _i32(facets).reshape(-1)
    keep = np.empty(n, np.uint8)
    ncomp = lib().ref_connected_keep(C.c_int64(n), _p(facets, C.c_int32), C.c_int64(len(facets) // 3),
                                     _p(keep, C.c_uint8))
    return keep.astype(bool), int(ncomp)


def depth2model(inv_depth, cam: Camera, min_dsp, max_dsp, smooth):
    d = _f32(inv_depth).reshape(-1)
    n = cam.w * cam.h
    xyz = np.empty((n, 3))
    tex = np.empty(n, np.int32)
    fac = np.empty((2 * (cam.w - 1) * (cam.h - 1), 3), np.int32)
    npts, nf = C.c_int64(), C.c_int64()
    lib().ref_depth2model(_p(d, C.c_float), _cam(cam), C.c_double(min_dsp), C.c_double(max_dsp),
                          C.c_double(smooth), _p(xyz, C.c_double), _p(tex, C.c_int32),
                          _p(fac, C.c_int32), C.byref(npts), C.byref(nf))
    return xyz[:npts.value].copy(), tex[:npts.value].copy(), fac[:nf.value].copy()


def calc_rotation(a, b):
    a, b = _f64(a).reshape(3), _f64(b).reshape(3)
    R = np.empty(9)
    lib().ref_calc_rotation(_p(a, C.c_double), _p(b, C.c_double), _p(R, C.c_double))
    return R.reshape(3, 3)


def axis_angle(axis, angle_deg, pb):
    axis, pb = _f64(axis).reshape(3), _f64(pb).reshape(3)
    out = np.empty(3)
    lib().ref_axis_angle(_p(axis, C.c_double), C.c_double(angle_deg), _p(pb, C.c_double),
                         _p(out, C.c_double))
    return out


def save_depth(filename: str, depth):
    d = _f64(depth).reshape(-1)
    lib().ref_save_depth(str(filename).encode(), _p(d, C.c_double), C.c_int64(len(d)))


def load_depth(filename: str, w: int, h: int):
    d = np.empty(w * h)
    lib().ref_load_depth(str(filename).encode(), _p(d, C.c_double), C.c_int(w), C.c_int(h))
    return d


# ----------------------------------------------------------------------------- Alignment / PlyObj
# (oracle/ref_shim_align.cpp: the reference's Alignment, PartRecognition and Mesh classes)
PART_NAMES = ["Head", "Neck", "LeftUpperArm", "LeftLowerArm", "LeftHand", "RightUpperArm",
              "RightLowerArm", "RightHand", "LeftThigh", "LeftShank", "LeftFoot", "RightThigh",
              "RightShank", "RightFoot", "Truncus", "Hip"]  # PartRecognition.h:13-30


def write_parts_file(workdir, labels):
    """The template's labels in the reference's own file format (`Head=1;5;9`,
    PartRecognition/PartRecognition.cpp:7-48), where Alignment::Align looks for them."""
    d = Path(workdir) / "Template" / "part"
    d.mkdir(parents=True, exist_ok=True)
    labels = np.asarray(labels)
    with open(d / "parts", "w") as f:
        for k, name in enumerate(PART_NAMES):
            ids = np.nonzero(labels == k)[0]
            if len(ids):
                f.write(name + "=" + ";".join(str(int(i)) for i in ids) + "\n")


def retain_connected(pts, normals, facets):
    p = _f64(pts).reshape(-1, 3).copy()
    nr = _f64(normals).reshape(-1, 3).copy()
    f = _i32(facets).reshape(-1).copy()
    n, nf = C.c_int64(len(p)), C.c_int64(len(f) // 3)
    lib().ref_retain_connected(_p(p, C.c_double), _p(nr, C.c_double), C.byref(n), _p(f, C.c_int32),
                               C.byref(nf))
    return p[:n.value].copy(), nr[:n.value].copy(), f[:3 * nf.value].reshape(-1, 3).copy()


def remove_ground(pts, normals, facets, dist_thres, workdir):
    p = _f64(pts).reshape(-1, 3).copy()
    nr = _f64(normals).reshape(-1, 3).copy()
    f = _i32(facets).reshape(-1).copy()
    n, nf = C.c_int64(len(p)), C.c_int64(len(f) // 3)
    gr = np.zeros(3)
    lib().ref_remove_ground(_p(p, C.c_double), _p(nr, C.c_double), C.byref(n), _p(f, C.c_int32),
                            C.byref(nf), C.c_double(dist_thres), str(workdir).encode(),
                            _p(gr, C.c_double))
    return (p[:n.value].copy(), nr[:n.value].copy(), f[:3 * nf.value].reshape(-1, 3).copy(), gr)


def init_alignment(src, tgt, ground_ray, view_ray):
    s, t = _f64(src).reshape(-1, 3), _f64(tgt).reshape(-1, 3)
    gr, vr = _f64(ground_ray).reshape(3), _f64(view_ray).reshape(3)
    T = np.empty(16)
    sc = C.c_double()
    lib().ref_init_alignment(_p(s, C.c_double), C.c_int64(len(s)), _p(t, C.c_double),
                             C.c_int64(len(t)), _p(gr, C.c_double), _p(vr, C.c_double),
                             _p(T, C.c_double), C.byref(sc))
    return T.reshape(4, 4), sc.value


def local_alignment_core(src, s_labels, tgt, t_labels, slabel, tlabel):
    s, t = _f64(src).reshape(-1, 3), _f64(tgt).reshape(-1, 3)
    sl, tl = _i32(s_labels), _i32(t_labels)
    T = np.empty(16)
    sc = C.c_double()
    lib().ref_local_alignment_core(_p(s, C.c_double), _p(sl, C.c_int32), C.c_int64(len(s)),
                                   _p(t, C.c_double), _p(tl, C.c_int32), C.c_int64(len(t)),
                                   C.c_int(slabel), C.c_int(tlabel), _p(T, C.c_double),
                                   C.byref(sc))
    return T.reshape(4, 4), sc.value


def part_recog(tmpl, labels, query):
    tm, q = _f64(tmpl).reshape(-1, 3), _f64(query).reshape(-1, 3)
    lb = _i32(labels)
    out = np.empty(len(q), np.int32)
    lib().ref_part_recog(_p(tm, C.c_double), _p(lb, C.c_int32), C.c_int64(len(tm)),
                         _p(q, C.c_double), C.c_int64(len(q)), _p(out, C.c_int32))
    return out


def align(src, s_normals, s_facets, s_labels, tgt, t_normals, t_facets, view_ray, dist_thres,
          workdir):
    """Alignment::Align as a whole; `workdir` receives Template/part/parts."""
    write_parts_file(workdir, s_labels)
    s, sn = _f64(src).reshape(-1, 3).copy(), _f64(s_normals).reshape(-1, 3).copy()
    sf = _i32(s_facets).reshape(-1)
    t, tn = _f64(tgt).reshape(-1, 3).copy(), _f64(t_normals).reshape(-1, 3).copy()
    tf = _i32(t_facets).reshape(-1).copy()
    nt, ntf = C.c_int64(len(t)), C.c_int64(len(tf) // 3)
    vr = _f64(view_ray).reshape(3)
    lib().ref_align(_p(s, C.c_double), _p(sn, C.c_double), C.c_int64(len(s)), _p(sf, C.c_int32),
                    C.c_int64(len(sf) // 3), _p(t, C.c_double), _p(tn, C.c_double), C.byref(nt),
                    _p(tf, C.c_int32), C.byref(ntf), _p(vr, C.c_double), C.c_double(dist_thres),
                    str(workdir).encode())
    n, f = nt.value, ntf.value
    return s, sn, t[:n].copy(), tn[:n].copy(), tf[:3 * f].reshape(-1, 3).copy()


def mesh_vertex_normals(pts, facets):
    p = _f64(pts).reshape(-1, 3)
    f = _i32(facets).reshape(-1)
    out = np.empty_like(p)
    lib().ref_mesh_vertex_normals(_p(p, C.c_double), C.c_int64(len(p)), _p(f, C.c_int32),
                                  C.c_int64(len(f) // 3), _p(out, C.c_double))
    return out


def write_obj(filename, pts, normals, facets):
    p = _f64(pts).reshape(-1, 3)
    nr = _f64(normals).reshape(-1, 3) if normals is not None else None
    f = _i32(facets).reshape(-1)
    lib().ref_write_obj(str(filename).encode(), _p(p, C.c_double), _p(nr, C.c_double),
                        C.c_int64(len(p)), _p(f, C.c_int32), C.c_int64(len(f) // 3))


def read_obj(filename, cap_n: int, cap_f: int):
    p, nr = np.empty((cap_n, 3)), np.empty((cap_n, 3))
    f = np.empty((cap_f, 3), np.int32)
    n, nf = C.c_int64(), C.c_int64()
    rc = lib().ref_read_obj(str(filename).encode(), _p(p, C.c_double), _p(nr, C.c_double),
                            C.c_int64(cap_n), _p(f, C.c_int32), C.c_int64(cap_f), C.byref(n),
                            C.byref(nf))
    if rc != 0:
        raise RuntimeError("ref.read_obj: capacity too small or no normals")
    return p[:n.value].copy(), nr[:n.value].copy(), f[:nf.value].copy()


# ----------------------------------------------------------------------------- Deformation (D2-D5, D10)
# (oracle/ref_shim_deform.cpp: the reference's Deformation class against the CGAL stand-in that
# records the constraints handed to Surface_mesh_deformation)
def _mesh(pts, normals, facets):
    p = _f64(pts).reshape(-1, 3)
    nr = _f64(normals).reshape(-1, 3) if normals is not None else np.zeros_like(p)
    f = _i32(facets).reshape(-1)
    return p, nr, f


def deform_uniform_sampling(pts, normals, facets, workdir):
    p, nr, f = _mesh(pts, normals, facets)
    out = np.empty(len(p), np.int32)
    lib().ref_deform_uniform_sampling.restype = C.c_int64
    m = lib().ref_deform_uniform_sampling(_p(p, C.c_double), _p(nr, C.c_double), C.c_int64(len(p)),
                                          _p(f, C.c_int32), C.c_int64(len(f) // 3),
                                          str(workdir).encode(), _p(out, C.c_int32))
    return out[:m].copy()


def deform_node_graph(pts, normals, facets, ids, K=8):
    p, nr, f = _mesh(pts, normals, facets)
    ids = _i32(ids)
    out = np.empty((len(ids), K + 1), np.int32)
    w = np.empty((len(ids), K + 1))
    lib().ref_deform_node_graph(_p(p, C.c_double), _p(nr, C.c_double), C.c_int64(len(p)),
                                _p(f, C.c_int32), C.c_int64(len(f) // 3), _p(ids, C.c_int32),
                                C.c_int64(len(ids)), C.c_int(K), _p(out, C.c_int32), _p(w, C.c_double))
    return out, w


def deform_node_targets(pts, normals, facets, ids, tpts, tnormals, proj_len_err, proj_dist_err,
                        workdir):
    """-> (vertex ids, target positions): what Deform(tpts, tnormals, projLenErr, projDistErr) hands
    to CGAL (Deformation.cpp:383-391), one constraint per node."""
    p, nr, f = _mesh(pts, normals, facets)
    ids = _i32(ids)
    tp, tn = _f64(tpts).reshape(-1, 3), _f64(tnormals).reshape(-1, 3)
    ov = np.empty(len(p), np.int32)
    ot = np.empty((len(p), 3))
    lib().ref_deform_node_targets.restype = C.c_int64
    k = lib().ref_deform_node_targets(_p(p, C.c_double), _p(nr, C.c_double), C.c_int64(len(p)),
                                      _p(f, C.c_int32), C.c_int64(len(f) // 3), _p(ids, C.c_int32),
                                      C.c_int64(len(ids)), _p(tp, C.c_double), _p(tn, C.c_double),
                                      C.c_int64(len(tp)), C.c_double(proj_len_err),
                                      C.c_double(proj_dist_err), str(workdir).encode(),
                                      _p(ov, C.c_int32), _p(ot, C.c_double))
    return ov[:k].copy(), ot[:k].copy()


def deform_node_targets_threshold(pts, normals, facets, ids, tpts, tnormals, threshold):
    p, nr, f = _mesh(pts, normals, facets)
    ids = _i32(ids)
    tp, tn = _f64(tpts).reshape(-1, 3), _f64(tnormals).reshape(-1, 3)
    ov = np.empty(len(p), np.int32)
    ot = np.empty((len(p), 3))
    lib().ref_deform_node_targets_threshold.restype = C.c_int64
    k = lib().ref_deform_node_targets_threshold(
        _p(p, C.c_double), _p(nr, C.c_double), C.c_int64(len(p)), _p(f, C.c_int32),
        C.c_int64(len(f) // 3), _p(ids, C.c_int32), C.c_int64(len(ids)), _p(tp, C.c_double),
        _p(tn, C.c_double), C.c_int64(len(tp)), C.c_double(threshold), _p(ov, C.c_int32),
        _p(ot, C.c_double))
    return ov[:k].copy(), ot[:k].copy()


def deform_export_obj(filename, pts, normals, facets):
    p, nr, f = _mesh(pts, normals, facets)
    lib().ref_deform_export_obj(str(filename).encode(), _p(p, C.c_double), _p(nr, C.c_double),
                                C.c_int64(len(p)), _p(f, C.c_int32), C.c_int64(len(f) // 3))


def deform_mesh_is_valid(pts, facets) -> bool:
    p = _f64(pts).reshape(-1, 3)
    f = _i32(facets).reshape(-1)
    return bool(lib().ref_deform_mesh_is_valid(_p(p, C.c_double), C.c_int64(len(p)), _p(f, C.c_int32),
                                               C.c_int64(len(f) // 3)))


# ----------------------------------------------------------------------------- Processor (§8(f)1)
# (oracle/ref_shim_processor.cpp: Processor::RemoveOutliers and Processor::CheckConsistencyCore)
def remove_outliers(p, q, cam1: Camera, cam2: Camera, pixel_err, adapt_ratio, seed: int):
    """Processor::RemoveOutliers after srand(seed) -> (kept original indices, err, inlier_ratio)."""
    p, q = _f64(p).reshape(-1, 3), _f64(q).reshape(-1, 3)
    keep = np.empty(len(p), np.int32)
    err, ir = C.c_double(), C.c_double()
    lib().ref_remove_outliers.restype = C.c_int64
    k = lib().ref_remove_outliers(_p(p, C.c_double), _p(q, C.c_double), C.c_int64(len(p)), _cam(cam1),
                                  _cam(cam2), C.c_double(pixel_err), C.c_double(adapt_ratio),
                                  C.c_uint(seed), _p(keep, C.c_int32), C.byref(err), C.byref(ir))
    return keep[:k].copy(), err.value, ir.value


def oracle_remove_outliers_with_ref_samples(p, q, cam1: Camera, cam2: Camera, pixel_err, adapt_ratio,
                                            seed: int, rounds=3, iter_num=200):
    """The ORACLE's round loop (oracle/pipeline.cpp) fed with the hypothesis samples the reference
    draws: srand(seed), then per round iter_num calls of the reference's Shuffle on the C library's
    rand() stream (the oracle calls straight into libmvsref.so for them)."""
    import oracle
    p, q = _f64(p).reshape(-1, 3), _f64(q).reshape(-1, 3)
    keep = np.empty(len(p), np.int32)
    nk = C.c_int64()
    err, ir = C.c_double(), C.c_double()
    lib().ref_srand(C.c_uint(seed))
    draw = C.cast(lib().ref_srt_triples_next, C.c_void_p)
    oracle.lib().orc_remove_outliers_cb(_p(p, C.c_double), _p(q, C.c_double), C.c_int64(len(p)),
                                        C.byref(cam1), C.byref(cam2), C.c_int(rounds),
                                        C.c_int(iter_num), C.c_double(pixel_err),
                                        C.c_double(adapt_ratio), draw, None, _p(keep, C.c_int32),
                                        C.byref(nk), C.byref(err), C.byref(ir))
    return keep[:nk.value].copy(), err.value, ir.value


def check_consistency(depth, cam: Camera, ref_depths, ref_cams, min_dsp, max_dsp, reproj_err: int):
    d = _f64(depth).reshape(-1).copy()
    rd = _f64(ref_depths).reshape(len(ref_cams), -1)
    arr = (Camera * max(1, len(ref_cams)))(*ref_cams)
    lib().ref_check_consistency(_p(d, C.c_double), _cam(cam), _p(rd, C.c_double), arr,
                                C.c_int(len(ref_cams)), C.c_double(min_dsp), C.c_double(max_dsp),
                                C.c_int(int(reproj_err)))
    return d.reshape(cam.h, cam.w)
