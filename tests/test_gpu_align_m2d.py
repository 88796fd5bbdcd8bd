"""GPU parity tests: point-set utilities, Alignment and Model2Depth through the C ABI vs the oracle.

Bars: transforms / PCA / ranges within 1e-9 relative of the float64 oracle (double accumulation on
both sides, different summation trees); arg-extrema, transferred labels, kept-vertex sets and facets
exact; Model2Depth images bit-exact (coverage from exact integer edge functions, 1/Z from FMA-free
float64 in the same order as the oracle)."""
import numpy as np
import pytest

import oracle
from multiviewstitch_b200 import api
from tests import synth

pytestmark = pytest.mark.gpu
RTOL = 1e-9


def test_apply_srt_matches_oracle(ctx):
    rng = np.random.default_rng(1)
    p, n = rng.normal(size=(10001, 3)), rng.normal(size=(10001, 3))
    s, R, t = synth.PLANTED_S, synth.PLANTED_R, synth.PLANTED_T
    for inv in (False, True):
        gp, gn = api.apply_srt(p, n, s, R, t, inverse=inv, ctx=ctx)
        op, on = oracle.apply_srt(p, n, s, R, t, inverse=inv)
        assert np.allclose(gp, op, rtol=0, atol=1e-13) and np.allclose(gn, on, rtol=0, atol=1e-13)
    gp, _ = api.apply_srt(p, None, s, R, t, ctx=ctx)
    assert np.allclose(gp, s * p @ R.T + t, atol=1e-13)


def test_apply_srt_dev_f32(ctx):
    import torch
    p = torch.randn(100000, 4, device="cuda")
    out = torch.empty_like(p)
    api.apply_srt_dev(p.data_ptr(), out.data_ptr(), len(p), synth.PLANTED_S, synth.PLANTED_R,
                      synth.PLANTED_T, ctx)
    ctx.sync()
    ref = (synth.PLANTED_S * p[:, :3].double().cpu().numpy() @ synth.PLANTED_R.T + synth.PLANTED_T)
    assert np.array_equal(out[:, :3].cpu().numpy(), ref.astype(np.float32))
    assert torch.equal(out[:, 3], p[:, 3])


def test_pointset_stats_and_range(ctx):
    rng = np.random.default_rng(0)
    A = rng.normal(size=(3, 3))
    p = rng.normal(size=(200003, 3)) * np.array([3.0, 1.0, 0.3]) @ A.T + np.array([1, -2, 3.0])
    ps = api.PointSetUtils(ctx)
    ps.SetInput(p)
    b, mn, mx, pv, ev = oracle.pointset_stats(p)
    assert np.allclose(ps.GetBarycenter(), b, rtol=RTOL, atol=1e-12)
    assert np.array_equal(ps.GetBoundingBox()[0], mn) and np.array_equal(ps.GetBoundingBox()[1], mx)
    assert np.allclose(ps.CalcPivots(), pv, atol=1e-9) and np.allclose(ps.Eigenvalues(), ev, rtol=RTOL)
    lab = rng.integers(0, 5, len(p)).astype(np.int32)
    for labels, l in ((None, 0), (lab, 3)):
        g = api.project_range(p, b, pv[:, 0] * 2.5, labels, l, ctx)
        o = oracle.project_range(p, b, pv[:, 0] * 2.5, labels, l)
        assert g[2:] == o[2:] and g[:2] == o[:2]  # same FMA-free float64 expression: exact
    q = np.array([[0, 0, 0], [2, 0, 0], [2, 0, 0], [-1, 0, 0], [-1, 0, 0.0]])
    assert api.project_range(q, [0, 0, 0], [2, 0, 0], ctx=ctx) == (-0.5, 1.0, 3, 1)


def test_retain_connected_and_remove_ground(ctx):
    v1, f1 = synth.uv_sphere(12, 9)
    v2, f2 = synth.uv_sphere(20, 15)
    v = np.concatenate([v1, v2 + 5.0], 0)
    f = np.concatenate([f1, f2 + len(v1)], 0)
    al = api.Alignment(ctx)
    gp, gn, gf = al.RetainConnectRegion(v, v.copy(), f)
    op, on, of = oracle.retain_connected(v, v.copy(), f)
    assert np.array_equal(gp, op) and np.array_equal(gf, of) and np.array_equal(gn, on)
    v, nr, f, nb = synth.body_on_ground()
    gp, gn, gf, gg = al.RemoveGround(v, nr, f)
    op, on, of, og = oracle.remove_ground(v, nr, f, 0.8, 0.28)
    assert np.array_equal(gp, op) and np.array_equal(gn, on) and np.array_equal(gf, of)
    assert np.allclose(gg, og, atol=1e-9) and gg[1] > 0.9
    bad = f.copy(); bad[0, 0] = len(v)
    with pytest.raises(api.MvsError):
        al.RemoveGround(v, nr, bad)


@pytest.mark.parametrize("local", [0, 1])
def test_align_matches_oracle(ctx, local):
    src, sl = synth.humanoid(300, seed=1, arm_angle=0.5, leg_angle=0.12)
    tgt0, tl0 = synth.humanoid(2000, seed=2, arm_angle=0.75, leg_angle=0.2)
    R = synth.axis_angle([0.1, 1.0, 0.2], 0.2)
    tgt = 1.6 * tgt0 @ R.T + np.array([0.2, 0.1, 3.0])
    nf = np.zeros((0, 3), np.int32)
    al = api.Alignment(ctx, api.align_params(remove_ground=0, local_align=local))
    gs, gsn, gt, gtn, gtf, gtl, gg = al.Align(src, src.copy(), sl, tgt, tgt.copy(), nf,
                                              R @ [0, 0, 1.0], ground_ray=R @ [0, 1.0, 0])
    os_, osn, ot, otn, otf, otl, og = oracle.align(src, src.copy(), sl, tgt, tgt.copy(), nf,
                                                   R @ [0, 0, 1.0], remove_ground=0,
                                                   local_align=local, ground_ray=R @ [0, 1.0, 0])
    assert np.array_equal(gtl, otl) and (gtl == tl0).mean() > 0.9
    assert np.allclose(gs, os_, rtol=0, atol=1e-8) and np.allclose(gsn, osn, rtol=0, atol=1e-8)
    assert np.array_equal(gt, tgt)


def test_align_with_ground_removal(ctx):
    # target = connected body mesh standing on a ground patch; template = the same body, smaller
    v, nr, f, nb = synth.body_on_ground()
    src = v[:nb] * 0.5 + np.array([3.0, 0, 0])
    sl = np.full(nb, 14, np.int32)
    sl[v[:nb, 1] < -0.8] = 1  # a "neck" cap so the neck-centroid translation runs
    al = api.Alignment(ctx, api.align_params(local_align=0))
    g = al.Align(src, nr[:nb], sl, v, nr, f, [0, 0, 1.0])
    o = oracle.align(src, nr[:nb], sl, v, nr, f, [0, 0, 1.0], local_align=0)
    assert np.array_equal(g[2], o[2]) and np.array_equal(g[4], o[4]) and np.array_equal(g[5], o[5])
    assert np.allclose(g[0], o[0], atol=1e-8) and np.allclose(g[6], o[6], atol=1e-9)
    assert len(g[2]) < nb  # ground patch and the sole of the body removed


def _cams(n=3, w=160, h=120, f=150.0):
    K = np.array([[f, 0, w / 2 - 0.5], [0, f, h / 2 - 0.5], [0, 0, 1.0]])
    out_a, out_o = [], []
    for k in range(n):
        R = synth.axis_angle([0.1, 1.0, 0.05], 0.35 * k)
        t = np.array([0.05 * k, -0.02 * k, 3.0 + 0.3 * k])
        out_a.append(api.make_camera(K, R, t, w, h))
        out_o.append(oracle.make_camera(K, R, t, w, h))
    return out_a, out_o


def test_model2depth_bit_exact(ctx):
    verts, faces = synth.uv_sphere(200, 150)  # ~30K sub-pixel to few-pixel triangles
    big = np.array([[-3, -3, 1.5], [3, -3, 1.5], [0, 4, 1.2]])  # one large triangle behind
    v = np.concatenate([verts, big]).astype(np.float32)
    f = np.concatenate([faces, [[len(verts), len(verts) + 1, len(verts) + 2]]]).astype(np.int32)
    ca, co = _cams()
    m2d = api.Model2Depth(ctx)
    m2d.SetInput([v], [None], [f], [ca])
    img = m2d.render(0)
    ref = oracle.model2depth(v, f, co)
    assert img.shape == ref.shape == (3, 120, 160)
    assert np.array_equal(img, ref)
    assert (img > 0).mean() > 0.3
    # float64 entry point
    out = np.empty_like(img)
    import ctypes as C
    arr = (api.Camera * 3)(*ca)
    v64 = v.astype(np.float64)
    api.check(api.lib().mvs_model2depth_f64(ctx.handle, v64.ctypes.data_as(C.POINTER(C.c_double)),
                                            C.c_int64(len(v)), f.ctypes.data_as(C.POINTER(C.c_int32)),
                                            C.c_int64(len(f)), arr, 3, C.c_float(0), C.c_float(0),
                                            out.ctypes.data_as(C.POINTER(C.c_float))))
    assert np.array_equal(out, ref)


def test_model2depth_known_answers_and_raw_io(ctx, tmp_path):
    K = np.array([[60.0, 0, 31.5], [0, 60.0, 23.5], [0, 0, 1.0]])
    cam = api.make_camera(K, np.eye(3), np.zeros(3), 64, 48)
    v = np.array([[-10, -10, 4], [10, -10, 4], [10, 10, 4], [-10, 10, 4]], np.float32)
    f = np.array([[0, 1, 2], [0, 2, 3]], np.int32)
    m2d = api.Model2Depth(ctx)
    m2d.SetInput([v], [None], [f], [[cam]], [str(tmp_path) + "/"])
    img = m2d.Run()[0][0]
    assert np.all(img == np.float32(0.25))  # fronto-parallel plane at z: inverse depth 1/z
    raw = tmp_path / "DATA" / "Render" / "_depth0.raw"  # Model2Depth.cpp:149-151
    assert raw.stat().st_size == 64 * 48 * 4
    assert np.array_equal(np.fromfile(raw, np.float32).reshape(48, 64), img)
    assert np.array_equal(api.load_depth_raw(str(raw), 64, 48), img)
    with pytest.raises(api.MvsError):
        api.load_depth_raw(str(raw), 65, 48)
    # empty mesh -> empty image; behind-camera triangles dropped
    m2d.SetInput([np.zeros((0, 3), np.float32)], [None], [np.zeros(0, np.int32)], [[cam]])
    assert np.all(m2d.render(0) == 0)
    vb = v.copy(); vb[:, 2] = -4
    m2d.SetInput([vb], [None], [f], [[cam]])
    assert np.all(m2d.render(0) == 0)


def test_model2depth_dev_large_mesh_matches_oracle(ctx):
    import torch
    verts, faces = synth.uv_sphere(400, 300)
    v = verts.astype(np.float32)
    ca, co = _cams(2, 320, 240, 300.0)
    v4 = torch.zeros((len(v), 4), device="cuda"); v4[:, :3] = torch.from_numpy(v).cuda()
    fd = torch.from_numpy(faces.reshape(-1)).cuda()
    out = torch.empty((2, 240, 320), device="cuda")
    api.model2depth_dev(v4.data_ptr(), len(v), fd.data_ptr(), len(faces), ca, out.data_ptr(), ctx=ctx)
    ctx.sync()
    assert np.array_equal(out.cpu().numpy(), oracle.model2depth(v, faces, co))


def test_unproject_depth_matches_oracle(ctx):
    (K, R1, t1, w, h), (_, R2, t2, _, _) = synth.camera_pair()
    rng = np.random.default_rng(0)
    dsp = rng.uniform(0.001, 0.4, size=(480, 640)).astype(np.float32)
    gp, gv = api.unproject_depth(dsp, api.make_camera(K, R2, t2, 640, 480), 0.0025, 0.3, ctx)
    op, ov = oracle.unproject_depth(dsp, oracle.make_camera(K, R2, t2, 640, 480), 0.0025, 0.3)
    assert np.array_equal(gv, ov) and np.array_equal(gp, op)
