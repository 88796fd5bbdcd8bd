"""Seeded inputs of the reference-parity cases, shared by tests/golden/make_ref_golden.py (which runs
them through oracle/_ref — the reference's own code compiled unmodified) and by
tests/test_ref_parity.py / tests/test_gpu_ref_parity.py (which run the oracle restatement and the
CUDA path on the same inputs).  `run(mod)` evaluates every case with either back end: `mod` is
`oracle.ref` (the reference) or an adapter with the same function names."""
from __future__ import annotations

import numpy as np

from multiviewstitch_b200 import synth

MIN_DSP, MAX_DSP, SMOOTH = 0.0025, 0.3, 0.12  # config.txt:23-24,38


def cams(make_camera, w=640, h=480):
    (K, R1, t1, _, _), (_, R2, t2, _, _) = synth.camera_pair()
    return make_camera(K, R1, t1, w, h), make_camera(K, R2, t2, w, h)


def srt_inputs():
    """name -> (p, q): planted similarity with noise / with outliers / an exact reflection (forces
    the det(R) = -1 branch, SRTSolver.cpp:115-119) / the three-pair minimum."""
    rng = np.random.default_rng(4242)
    out = {}
    p, q = synth.planted_pairs(2000, seed=5, noise=1e-3)
    out["noisy"] = (p, q)
    p2, q2 = synth.planted_pairs(500, seed=6, noise=2e-3)
    bad = rng.random(500) < 0.1
    q2 = q2.copy()
    q2[bad] += rng.normal(scale=0.3, size=(int(bad.sum()), 3))
    out["outliers"] = (p2, q2)
    p3 = synth.sphere_points(300, 7, center=(0, 0, 5.0))
    M = synth.PLANTED_R @ np.diag([1.0, 1.0, -1.0])      # improper: a mirrored copy
    q3 = synth.PLANTED_S * p3 @ M.T + synth.PLANTED_T
    out["mirror"] = (p3, q3)
    p4, q4 = synth.planted_pairs(3, seed=8, noise=1e-3)
    out["three"] = (p4, q4)
    return out


def project_points():
    """world points around both cameras, including ones that land at negative pixel coordinates
    (the (int)(x + 0.5) truncation toward zero, Camera.cpp:46-47) and behind the camera"""
    rng = np.random.default_rng(99)
    a = rng.normal(size=(4000, 3)) * np.array([3.0, 3.0, 2.0]) + np.array([0, 0, 5.0])
    b = rng.normal(size=(500, 3)) * 6.0
    return np.concatenate([a, b], 0)


def unproject_inputs(w=640, h=480):
    """distinct pixels with an inverse depth ("dsp", float32 as in the .raw files) inside the valid
    window; depth = 1 / dsp as Image3D.cpp:98-103 / Depth2Model.cpp:37 hand it to the camera"""
    rng = np.random.default_rng(100)
    pix = rng.choice(w * h, 3000, replace=False)
    uv = np.stack([pix % w, pix // w], 1).astype(np.int32)
    dsp = rng.uniform(MIN_DSP * 1.01, MAX_DSP * 0.99, 3000).astype(np.float32)
    return uv, dsp, 1.0 / dsp.astype(np.float64)


def pointsets():
    rng = np.random.default_rng(17)
    body, _ = synth.humanoid(300, seed=3)
    ell = rng.normal(size=(5000, 3)) * np.array([3.0, 1.0, 0.3]) @ synth.axis_angle([1, 1, 0], 0.7).T + 2.0
    return {"humanoid": body, "ellipsoid": ell}


def facet_sets():
    """name -> (n, facets): two components of different size; two of EQUAL size (ProminentRepresent
    keeps the first root with the largest rank, UnionSetUtils.cpp:33-43); isolated vertices"""
    v1, f1 = synth.uv_sphere(12, 8)
    v2, f2 = synth.uv_sphere(8, 6)
    big_small = (len(v1) + len(v2), np.concatenate([f1, f2 + len(v1)], 0))
    small_big = (len(v1) + len(v2), np.concatenate([f2, f1 + len(v2)], 0))
    equal = (2 * len(v2) + 5, np.concatenate([f2 + 5, f2 + 5 + len(v2)], 0))
    rng = np.random.default_rng(5)
    perm = rng.permutation(len(v1) + len(v2))
    shuffled = (len(v1) + len(v2), perm[big_small[1]])
    return {"big_small": big_small, "small_big": small_big, "equal": equal, "shuffled": shuffled}


def depth_map(w, h, seed=0):
    """Synthetic inverse-depth image (float32, the `.raw` format): a bumpy sphere in front of a flat
    wall, with a disparity discontinuity at the silhouette (exercises the three-edge smoothness test,
    Depth2Model.cpp:52-75), out-of-range and zero pixels."""
    rng = np.random.default_rng(seed)
    fx = 525.0 * w / 640.0
    u, v = np.meshgrid(np.arange(w), np.arange(h))
    x, y = (u - (w - 1) / 2) / fx, (v - (h - 1) / 2) / fx
    # ray/sphere: centre (0,0,5), radius 1 (+ bumps)
    rad = 1.0 + 0.05 * np.sin(9 * x) * np.cos(7 * y)
    a = x * x + y * y + 1.0
    disc = 25.0 - a * (25.0 - rad * rad)
    z = np.full((h, w), 9.0 + 0.5 * x)                    # wall
    hit = disc > 0
    z[hit] = ((5.0 - np.sqrt(disc[hit])) / a[hit])
    d = (1.0 / z).astype(np.float32)
    d[rng.random((h, w)) < 0.02] = 0.0                   # holes
    d[:3, :] = 0.5                                        # > MaxDsp
    d[-2:, :] = 0.001                                     # < MinDsp
    return d


def run(mod, make_camera, d2m_size=(160, 120)):
    """Every case through `mod`; returns {key: ndarray}."""
    out = {}
    c1, c2 = cams(make_camera)
    for name, (p, q) in srt_inputs().items():
        s, R, t = mod.srt_estimate(p, q, c1, c2)
        out[f"srt_{name}_s"], out[f"srt_{name}_R"], out[f"srt_{name}_t"] = np.float64(s), R, t
        out[f"srt_{name}_res"] = np.float64(mod.srt_residual(p, q, c1, c2, s, R, t))
        out[f"srt_{name}_res_planted"] = np.float64(
            mod.srt_residual(p, q, c1, c2, synth.PLANTED_S, synth.PLANTED_R, synth.PLANTED_T))
    for name, seed, iters in (("noisy", 1, 40), ("outliers", 2, 200)):
        p, q = srt_inputs()[name]
        tr = mod.srt_triples(len(p), iters, seed)
        s, R, t = mod.srt_ransac(p, q, c1, c2, iters, seed, False, tr)
        out[f"ransac_{name}_triples"] = tr
        out[f"ransac_{name}_s"], out[f"ransac_{name}_R"], out[f"ransac_{name}_t"] = np.float64(s), R, t
        s2, R2, t2 = mod.srt_ransac(p, q, c1, c2, iters, seed, True, tr)
        out[f"ransac_{name}_array_R"], out[f"ransac_{name}_array_t"] = R2, t2
    pts = project_points()
    out["project_uv1"] = mod.project(c1, pts)
    out["project_uv2"] = mod.project(c2, pts)
    uv, _, depth = unproject_inputs()
    out["unproject_1"] = mod.unproject(c1, uv, depth)
    out["unproject_2"] = mod.unproject(c2, uv, depth)
    for name, ps in pointsets().items():
        b, c, P = mod.pointset(ps)
        out[f"ps_{name}_bary"], out[f"ps_{name}_bbc"], out[f"ps_{name}_pivots"] = b, c, P
    for name, (n, f) in facet_sets().items():
        keep, ncomp = mod.connected_keep(n, f)
        out[f"cc_{name}_keep"] = np.asarray(keep, bool)
        out[f"cc_{name}_ncomp"] = np.int64(ncomp)
    w, h = d2m_size
    cd = cams(make_camera, w, h)[1]
    xyz, tex, fac = mod.depth2model(depth_map(w, h), cd, MIN_DSP, MAX_DSP, SMOOTH)
    out["d2m_xyz"], out["d2m_tex"], out["d2m_facets"] = xyz, tex, fac
    return out


# ------------------------------------------------------------------ Alignment / PartRecognition / PlyObj
# (oracle/ref_shim_align.cpp: the reference's Alignment.cpp, PartRecognition.cpp, PlyObj.cpp)
DIST_THRES = 0.8  # config.txt: dist_thres
LIMBS = (((2, 3, 4), 4), ((5, 6, 7), 7), ((8, 9), 9), ((11, 12), 12))  # Alignment.cpp:378-419


def align_scene():
    """Template = a small labelled humanoid mesh; target = the same figure in another pose, 1.6x
    larger, turned and moved, standing on a ground patch — one connected mesh, as
    Alignment::Align expects it (RemoveGround + RetainConnectRegion need the facets)."""
    sv, sf, sl = synth.humanoid_mesh(10, 8, arm_angle=0.5, leg_angle=0.12)
    sn = synth.vertex_normals(sv, sf)
    tv0, tf, _ = synth.humanoid_mesh(14, 10, arm_angle=0.75, leg_angle=0.2, ground=(0.7, 36))
    R = synth.axis_angle([0.1, 1.0, 0.2], 0.2)
    tv = 1.6 * tv0 @ R.T + np.array([0.2, 0.1, 3.0])
    tn = synth.vertex_normals(tv, tf)
    return dict(sv=sv, sn=sn, sf=sf, sl=sl, tv=tv, tn=tn, tf=tf, view_ray=R @ [0, 0, 1.0],
                ground_ray=R @ [0, 1.0, 0])


def run_align(mod):
    """Every alignment case through `mod` (oracle.ref = the reference, or an adapter)."""
    out = {}
    # A7 RetainConnectRegion, full outputs
    v1, f1 = synth.uv_sphere(10, 7)
    v2, f2 = synth.uv_sphere(14, 9)
    v = np.concatenate([v1, v2 + 5.0], 0)
    f = np.concatenate([f1, f2 + len(v1)], 0)
    p, n, ff = mod.retain_connected(v, v.copy(), f)
    out["rc_pts"], out["rc_normals"], out["rc_facets"] = p, n, ff
    # A4 RemoveGround
    bv, bn, bf, nb = synth.body_on_ground()
    p, n, ff, gr = mod.remove_ground(bv, bn, bf, DIST_THRES)
    out["rg_pts"], out["rg_normals"], out["rg_facets"], out["rg_ground_ray"] = p, n, ff, gr
    # A2 InitAlignment: a planted similarity, and two different poses
    src, sl = synth.humanoid(200, seed=1)
    R = synth.axis_angle([0.1, 1.0, 0.2], 0.25)
    tgt = 1.8 * src @ R.T + np.array([0.5, -0.3, 4.0])
    out["ia_planted_T"], s = mod.init_alignment(src, tgt, R @ [0, 1.0, 0], R @ [0, 0, 1.0])
    out["ia_planted_s"] = np.float64(s)
    sc = align_scene()
    body = sc["tv"][: -len(synth.grid_disk(0.7, 36, 0.0)[0])]
    T, s = mod.init_alignment(sc["sv"], body, sc["ground_ray"], sc["view_ray"])
    out["ia_pose_T"], out["ia_pose_s"] = T, np.float64(s)
    # A5 PartRecog on the initially aligned template, then A6 LocalAlignmentCore per limb
    s1 = s * sc["sv"] @ T[:3, :3].T + T[:3, 3]
    tl = mod.part_recog(s1, sc["sl"], body)
    out["pr_labels"] = np.asarray(tl, np.int32)
    for sel, lab in LIMBS:
        Tl, sl_ = mod.local_core(s1, sc["sl"], body, tl, sel, lab)
        out[f"lc_{lab}_T"], out[f"lc_{lab}_s"] = Tl, np.float64(sl_)
    keep = np.asarray(tl) != 4  # the scan misses a hand: the label-set bookkeeping of :477-497
    Tl, sl_ = mod.local_core(s1, sc["sl"], body[keep], np.asarray(tl)[keep], (2, 3, 4), 4)
    out["lc_nohand_T"], out["lc_nohand_s"] = Tl, np.float64(sl_)
    # Alignment::Align as a whole
    a = mod.align(sc["sv"], sc["sn"], sc["sf"], sc["sl"], sc["tv"], sc["tn"], sc["tf"],
                  sc["view_ray"], DIST_THRES)
    out["al_src"], out["al_src_normals"], out["al_tgt"], out["al_tgt_normals"], out["al_tgt_facets"] = a
    # PlyObj vertex normals
    mv, mf = synth.uv_sphere(24, 18, bumps=0.08)
    mv = mv.astype(np.float32).astype(np.float64)
    out["vn_normals"] = mod.vertex_normals(mv, mf)
    return out


# ------------------------------------------------------------------ Deformation (rows D1-D5, D10)
# (oracle/ref_shim_deform.cpp: the reference's Deformation.h / Deformation.cpp; CGAL's solver is a
# recorder there, so the outputs are the constraints the reference hands to it)
def deform_scene():
    """Template: a 9.6K-vertex cube-sphere (float32-representable: the reference reads its meshes
    through float OBJ text, PlyObj.cpp:44-55); target: a larger bumpy UV sphere with normals."""
    verts, faces = synth.cube_sphere(40)
    pts = verts.astype(np.float32)
    nrm = synth.vertex_normals(pts.astype(np.float64), faces).astype(np.float32)
    tv, tf = synth.uv_sphere(150, 110, radius=1.08, bumps=0.03)
    tv = tv.astype(np.float32)
    tn = synth.vertex_normals(tv.astype(np.float64), tf).astype(np.float32)
    return dict(pts=pts, nrm=nrm, faces=faces, tv=tv, tn=tn)


DEFORM_THRESHOLDS = ((100.0, 100.0), (0.085, 100.0), (100.0, 0.012))  # (projLenErr, projDistErr)
DEFORM_D4P_THRESHOLDS = (0.004, 0.0075, 1.0)                          # squared distances


def run_deform(mod):
    out = {}
    sc = deform_scene()
    ids = mod.uniform_sampling(sc)                      # D2, Deformation.cpp:63-106
    out["d2_ids"] = np.asarray(ids, np.int32)
    out["d3_graph"] = mod.node_graph(sc, ids, 8)        # D3, :108-153
    for k, (pl, pd) in enumerate(DEFORM_THRESHOLDS):    # D4 + D5, :232-391
        out[f"d4_controls_{k}"] = mod.node_targets(sc, ids, out["d3_graph"], pl, pd)
    for k, thr in enumerate(DEFORM_D4P_THRESHOLDS):     # D4', :155-208
        v, t = mod.node_targets_threshold(sc, ids, thr)
        out[f"d4p_vertices_{k}"], out[f"d4p_targets_{k}"] = np.asarray(v, np.int32), t
    # D1: the reference accepts only meshes CGAL calls valid (:42-45); a facet listed twice is not
    bad = np.concatenate([sc["faces"], sc["faces"][:1]], 0)
    out["d1_valid"] = np.array([mod.mesh_is_valid(sc["pts"], sc["faces"]),
                                mod.mesh_is_valid(sc["pts"], bad)], np.int32)
    return out


# ------------------------------------------------------------------ Processor (§8(f)1)
# (oracle/ref_shim_processor.cpp: Processor::RemoveOutliers, Processor::CheckConsistencyCore)
REMOVE_OUTLIER_CASES = ((1, 20.0, 0.6), (5, 60.0, 0.75), (77, 3.0, 0.5))  # (seed, pixel_err, adapt ratio)


def outlier_inputs():
    p, q = synth.planted_pairs(400, seed=3, noise=5e-4)
    rng = np.random.default_rng(0)
    bad = rng.choice(400, 40, replace=False)
    q = q.copy()
    q[bad] += rng.normal(scale=0.4, size=(40, 3))
    return p, q, bad


def consistency_inputs(make_camera, model2depth):
    """Three 160x120 views of a sphere, the middle one checked against its neighbours; one neighbour
    carries an inconsistent patch.  Depth maps are float32 inverse depths (the `.raw` format)."""
    w, h = 160, 120
    verts, faces = synth.cube_sphere(48)
    verts = (verts + np.array([0, 0, 4.0])).astype(np.float32)
    K = np.array([[150.0, 0, w / 2 - 0.5], [0, 150.0, h / 2 - 0.5], [0, 0, 1.0]])
    poses = [(synth.axis_angle([0, 1, 0], 0.06 * (k - 1)), np.array([0.15 * (k - 1), 0, 0.0])) for k in range(3)]
    cams = [make_camera(K, R, t, w, h) for R, t in poses]
    maps = model2depth(verts, faces, poses, K, w, h)
    maps = np.array(maps, np.float32)
    maps[2][40:60, 70:90] *= 1.3
    return cams, maps


def run_processor(mod, make_camera, model2depth):
    out = {}
    p, q, _ = outlier_inputs()
    c1, c2 = cams(make_camera)
    for seed, pe, ar in REMOVE_OUTLIER_CASES:
        keep, err, rounds = mod.remove_outliers(p, q, c1, c2, pe, ar, seed)
        out[f"ro_{seed}_keep"], out[f"ro_{seed}_err"] = np.asarray(keep, np.int32), np.float64(err)
        for k, tr in enumerate(rounds):
            out[f"ro_{seed}_round{k}_triples"] = np.asarray(tr, np.int32)
    cs, maps = consistency_inputs(make_camera, model2depth)
    out["cc_depth"] = np.asarray(mod.check_consistency(maps[1], cs[1], maps[[0, 2]], [cs[0], cs[2]],
                                                       MIN_DSP, 0.5, 4), np.float32)
    return out
