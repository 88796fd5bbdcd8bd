"""CPU tests pinning the alignment / point-set / Model2Depth oracle (oracle/align.cpp,
oracle/model2depth.cpp) with independent numpy re-derivations and analytic known answers.
The reference has no tests or golden vectors for these paths."""
import numpy as np
import pytest

import oracle
from tests import synth


def test_pointset_stats_vs_numpy():
    rng = np.random.default_rng(0)
    A = rng.normal(size=(3, 3))
    p = rng.normal(size=(5000, 3)) * np.array([3.0, 1.0, 0.3]) @ A.T + np.array([1, -2, 3.0])
    b, mn, mx, pv, ev = oracle.pointset_stats(p)
    assert np.allclose(b, p.mean(0), atol=1e-12)
    assert np.allclose(mn, p.min(0)) and np.allclose(mx, p.max(0))
    C = np.cov(p.T)  # /(n-1), PointSetUtils.cpp:26
    w, Q = np.linalg.eigh(C)
    assert np.allclose(ev, w[::-1], rtol=1e-10)  # descending, PointSetUtils.cpp:34-39
    for c in range(3):
        q = Q[:, 2 - c]
        assert abs(abs(pv[:, c] @ q) - 1) < 1e-10 and abs(np.linalg.norm(pv[:, c]) - 1) < 1e-12
        assert pv[np.argmax(np.abs(pv[:, c])), c] > 0  # documented sign convention


def test_apply_srt_and_inverse():
    rng = np.random.default_rng(1)
    p, n = rng.normal(size=(1000, 3)), rng.normal(size=(1000, 3))
    s, R, t = synth.PLANTED_S, synth.PLANTED_R, synth.PLANTED_T
    q, m = oracle.apply_srt(p, n, s, R, t)
    assert np.allclose(q, s * p @ R.T + t, atol=1e-13) and np.allclose(m, n @ R.T, atol=1e-13)
    p2, n2 = oracle.apply_srt(q, m, s, R, t, inverse=True)  # Processor.cpp:1177-1189
    assert np.allclose(p2, p, atol=1e-12) and np.allclose(n2, n, atol=1e-12)


def test_project_range_first_extremum_wins():
    p = np.array([[0, 0, 0], [2, 0, 0], [2, 0, 0], [-1, 0, 0], [-1, 0, 0.0]])
    tmin, tmax, imin, imax = oracle.project_range(p, [0, 0, 0], [2, 0, 0])
    assert (tmin, tmax, imin, imax) == (-0.5, 1.0, 3, 1)  # t = axis.(p-o)/|axis|^2, strict compares
    lab = np.array([0, 1, 2, 2, 1], np.int32)
    assert oracle.project_range(p, [0, 0, 0], [2, 0, 0], lab, 2)[2:] == (3, 2)


def test_retain_connected_keeps_largest_component():
    v1, f1 = synth.uv_sphere(12, 9)
    v2, f2 = synth.uv_sphere(20, 15)
    v = np.concatenate([v1, v2 + 5.0], 0)
    f = np.concatenate([f1, f2 + len(v1)], 0)
    p, n, ff = oracle.retain_connected(v, v.copy(), f)
    assert len(p) == len(v2) and np.allclose(p, v2 + 5.0) and np.array_equal(ff, f2)
    assert np.allclose(n, p)


def test_remove_ground():
    v, nr, f, nb = synth.body_on_ground()
    p, n, ff, gr = oracle.remove_ground(v, nr, f, 0.8, 0.28)
    assert gr[1] > 0.9  # ground is at +y (feet), groundRay points to it (Alignment.cpp:129-138)
    assert len(p) < nb and len(p) > 0.8 * nb  # ground patch gone, body kept (minus its sole)
    assert p[:, 1].max() < v[:nb, 1].max() - 0.01
    assert ff.max() < len(p) and ff.min() == 0
    # every kept point is an original body point
    from scipy.spatial import cKDTree
    assert cKDTree(v[:nb]).query(p)[0].max() == 0.0


def test_init_alignment_recovers_scale_and_pose():
    src, sl = synth.humanoid(300, seed=1)
    R = synth.axis_angle([0.1, 1.0, 0.2], 0.25)
    s, t = 1.8, np.array([0.5, -0.3, 4.0])
    tgt = s * src @ R.T + t
    T, sc = oracle.init_alignment(src, tgt, R @ [0, 1.0, 0], R @ [0, 0, 1.0])
    assert abs(sc - s) < 1e-9 * s
    out = sc * src @ T[:3, :3].T + T[:3, 3]
    assert np.abs(out - tgt).max() < 1e-8
    assert abs(np.linalg.det(T[:3, :3]) - 1) < 1e-9


def test_align_pipeline_moves_template_onto_target():
    src, sl = synth.humanoid(300, seed=1, arm_angle=0.5, leg_angle=0.12)
    tgt0, tl0 = synth.humanoid(500, seed=2, arm_angle=0.75, leg_angle=0.2)  # different pose
    R = synth.axis_angle([0.1, 1.0, 0.2], 0.2)
    tgt = 1.6 * tgt0 @ R.T + np.array([0.2, 0.1, 3.0])
    out = oracle.align(src, src.copy(), sl, tgt, tgt.copy(), np.zeros((0, 3), np.int32),
                       R @ [0, 0, 1.0], remove_ground=0, local_align=1,
                       ground_ray=R @ [0, 1.0, 0])
    s2, sn2, t2, tn2, tf2, tl, gr = out
    assert (tl == tl0).mean() > 0.9  # 1-NN label transfer (PartRecognition.cpp:50-77)
    from scipy.spatial import cKDTree
    d_local = cKDTree(t2).query(s2)[0]
    out0 = oracle.align(src, src.copy(), sl, tgt, tgt.copy(), np.zeros((0, 3), np.int32),
                        R @ [0, 0, 1.0], remove_ground=0, local_align=0, ground_ray=R @ [0, 1.0, 0])
    d_init = cKDTree(out0[2]).query(out0[0])[0]
    arms = np.isin(sl, [2, 3, 4, 5, 6, 7])
    assert d_local[arms].mean() < 0.6 * d_init[arms].mean()  # per-limb alignment helps the arms
    assert d_local.mean() < 0.1 * 1.6


def _cam(w=64, h=48, f=60.0, R=None, t=None):
    K = np.array([[f, 0, w / 2 - 0.5], [0, f, h / 2 - 0.5], [0, 0, 1.0]])
    return oracle.make_camera(K, np.eye(3) if R is None else R, np.zeros(3) if t is None else t, w, h)


def _plane_mesh(z_of_xy, ext=10.0):
    xy = np.array([[-ext, -ext], [ext, -ext], [ext, ext], [-ext, ext]])
    v = np.array([[x, y, z_of_xy(x, y)] for x, y in xy], np.float32)
    return v, np.array([[0, 1, 2], [0, 2, 3]], np.int32)


def test_model2depth_fronto_parallel_plane():
    cam = _cam()
    v, f = _plane_mesh(lambda x, y: 4.0)
    img = oracle.model2depth(v, f, [cam])[0]
    assert img.shape == (48, 64) and np.all(img == np.float32(0.25))  # inverse depth 1/z everywhere


def test_model2depth_slanted_plane_samples_pixel_centres():
    # plane z = 5 + 0.2 x + 0.1 y; a ray through image position (u,v) hits it at
    # Z = 5 / (1 - 0.2 (u-cx)/f - 0.1 (v-cy)/f); GL samples at (i+0.5, j+0.5) (oracle header)
    w, h, fl = 64, 48, 60.0
    cam = _cam(w, h, fl)
    v, f = _plane_mesh(lambda x, y: 5 + 0.2 * x + 0.1 * y, ext=8.0)
    img = oracle.model2depth(v, f, [cam])[0]
    i, j = np.meshgrid(np.arange(w), np.arange(h))
    cx, cy = w / 2 - 0.5, h / 2 - 0.5
    Z = 5.0 / (1 - 0.2 * (i + 0.5 - cx) / fl - 0.1 * (j + 0.5 - cy) / fl)
    assert np.all(img > 0)
    assert np.allclose(img, 1.0 / Z, rtol=2e-4)  # vertex snapping to 1/256 px
    Zint = 5.0 / (1 - 0.2 * (i - cx) / fl - 0.1 * (j - cy) / fl)
    assert np.abs(img - 1 / Z).max() < 0.2 * np.abs(img - 1 / Zint).max()


def test_model2depth_occlusion_clipping_and_pose():
    cam = _cam()
    v1, f1 = _plane_mesh(lambda x, y: 6.0)
    v2 = np.array([[-0.5, -0.5, 3], [0.5, -0.5, 3], [0, 0.5, 3]], np.float32)  # small near triangle
    v = np.concatenate([v1, v2]); f = np.concatenate([f1, [[4, 5, 6]]]).astype(np.int32)
    img = oracle.model2depth(v, f, [cam])[0]
    near = np.isclose(img, 1 / 3.0)
    assert 20 < near.sum() < 400 and np.all(np.isclose(img[~near], 1 / 6.0))  # LEQUAL: nearest wins
    # far/near clipping (Model2Depth.cpp:134: z_b <= 0 or >= 1 skipped)
    assert np.all(oracle.model2depth(v1, f1, [cam], 0.01, 5.0)[0] == 0)
    # reversed winding renders identically (no face culling in InitGL)
    assert np.array_equal(oracle.model2depth(v, f[:, ::-1].copy(), [cam])[0], img)
    # camera pose: x_cam = R x_world + t (Camera.cpp:68-72)
    R = synth.axis_angle([0, 1, 0], 0.3); t = np.array([0.1, 0.2, 0.5])
    vw = ((v1.astype(np.float64) - t) @ R).astype(np.float32)  # world points that map onto v1
    img2 = oracle.model2depth(vw, f1, [_cam(R=R, t=t)])[0]
    assert np.allclose(img2, 1 / 6.0, rtol=1e-5)


def test_unproject_depth_roundtrip():
    (K, R1, t1, w, h), (_, R2, t2, _, _) = synth.camera_pair()
    cam = oracle.make_camera(K, R2, t2, 64, 48)
    rng = np.random.default_rng(0)
    dsp = rng.uniform(0.001, 0.4, size=(48, 64)).astype(np.float32)
    pts, valid = oracle.unproject_depth(dsp, cam, 0.0025, 0.3)
    assert np.array_equal(valid.reshape(48, 64), ((dsp >= 0.0025) & (dsp <= 0.3)).astype(np.int32))
    for k in np.flatnonzero(valid)[::97]:
        j, i = divmod(k, 64)
        assert oracle.project(cam, pts[k]) == (i, j)  # Image3D.cpp:92-106 inverts Camera.cpp:45-59
        zc = (R2 @ pts[k] + t2)[2]
        assert abs(zc - 1.0 / float(dsp[j, i])) < 1e-9 * zc
