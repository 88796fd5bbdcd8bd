"""GPU parity tests for the callers / data formats either side of the solver (RemoveOutliers,
CheckConsistencyCore, Depth2Model) and the configs[0] case: a two-view 640x480 depth-map pair
through the whole hot path (unproject -> RemoveOutliers -> SRT -> Depth2Model mesh -> apply ->
deformation -> Model2Depth), stage by stage against the oracle."""
import numpy as np
import pytest

import oracle
from multiviewstitch_b200 import _capi, api, pipeline
from tests import synth

pytestmark = pytest.mark.gpu

from pathlib import Path
ROOT_DIR = Path(__file__).resolve().parents[1]


def _cam(mod, w=64, h=48, f=60.0, R=None, t=None):
    K = np.array([[f, 0, w / 2 - 0.5], [0, f, h / 2 - 0.5], [0, 0, 1.0]])
    return mod.make_camera(K, np.eye(3) if R is None else R, np.zeros(3) if t is None else t, w, h)


def test_depth2model_matches_oracle(ctx):
    rng = np.random.default_rng(0)
    w, h = 160, 120
    yy, xx = np.mgrid[0:h, 0:w]
    d = (0.2 + 0.0004 * np.sin(xx / 9.0) + 0.0003 * np.cos(yy / 7.0)).astype(np.float32)
    d[:, 90:] -= 0.05                       # depth jump
    d[rng.random((h, w)) < 0.03] = 0.0      # holes
    d[5:9, 5:30] = 0.4                      # beyond MaxDsp
    R = synth.axis_angle([0.2, 1, 0.1], 0.3); t = np.array([0.1, -0.2, 0.4])
    gx, gt, gf = api.depth2model(d, _cam(api, w, h, 150.0, R, t), 0.0025, 0.3, 0.12, ctx)
    ox, ot, of = oracle.depth2model(d, _cam(oracle, w, h, 150.0, R, t), 0.0025, 0.3, 0.12)
    assert np.array_equal(gt, ot) and np.array_equal(gf, of) and np.array_equal(gx, ox)
    assert len(gf) > 10000
    # empty image
    e = api.depth2model(np.zeros((h, w), np.float32), _cam(api, w, h), 0.0025, 0.3, 0.12, ctx)
    assert len(e[0]) == 0 and len(e[2]) == 0


def test_check_consistency_matches_oracle(ctx):
    rng = np.random.default_rng(1)
    w, h = 160, 120
    cams_a, cams_o, maps = [], [], []
    verts, faces = synth.cube_sphere(48)
    verts = (verts + np.array([0, 0, 4.0])).astype(np.float32)
    K = np.array([[150.0, 0, w / 2 - 0.5], [0, 150.0, h / 2 - 0.5], [0, 0, 1.0]])
    for k in range(3):
        R = synth.axis_angle([0, 1, 0], 0.06 * (k - 1)); t = np.array([0.15 * (k - 1), 0, 0.0])
        cams_a.append(api.make_camera(K, R, t, w, h)); cams_o.append(oracle.make_camera(K, R, t, w, h))
    maps = oracle.model2depth(verts, faces, cams_o)
    maps[2][40:60, 70:90] *= 1.3            # an inconsistent patch in one neighbour
    g = api.check_consistency(maps[1], cams_a[1], maps[[0, 2]], [cams_a[0], cams_a[2]], 0.0025, 0.5, 4.0, ctx)
    o = oracle.check_consistency(maps[1], cams_o[1], maps[[0, 2]], [cams_o[0], cams_o[2]], 0.0025, 0.5, 4.0)
    assert np.array_equal(g, o)
    assert 0.1 < (g > 0).mean() < (maps[1] > 0).mean()


def test_remove_outliers_matches_oracle(ctx):
    p, q = synth.planted_pairs(3000, seed=3, noise=5e-4)
    rng = np.random.default_rng(0)
    bad = rng.choice(3000, 300, replace=False)
    q[bad] += rng.normal(scale=0.4, size=(300, 3))
    (K, R1, t1, w, h), (_, R2, t2, _, _) = synth.camera_pair()
    ca = [api.make_camera(K, R1, t1, w, h), api.make_camera(K, R2, t2, w, h)]
    co = [oracle.make_camera(K, R1, t1, w, h), oracle.make_camera(K, R2, t2, w, h)]
    gk, ge, gi = api.remove_outliers(p, q, ca[0], ca[1], 3, 200, 20.0, 0.6, 5, ctx)
    ok, oe, oi = oracle.remove_outliers(p, q, co[0], co[1], 3, 200, 20.0, 0.6, 5)
    assert np.array_equal(gk, ok) and gi == oi == 0.0 and abs(ge - oe) <= 1e-9 * oe
    assert len(set(gk.tolist()) & set(bad.tolist())) <= 0.1 * len(bad)  # pixel-level test: a few slip through
    # the inlier test alone
    s = api.SRTSolver(10, ctx); s.SetPrintFlag(False); s.SetInput(p, q, ca[0], ca[1])
    sc, R, t = oracle.srt_estimate(p[ok], q[ok])
    gf, ges, gn = s.Inliers(sc, R, t, 3.0)
    of, oes = oracle.srt_inliers(p, q, co[0], co[1], sc, R, t, 3.0)
    assert np.array_equal(gf, of) and abs(ges - oes) <= 1e-9 * oes and gn == of.sum()


def test_two_view_pair_end_to_end(ctx):
    """configs[0]: two 640x480 views of the C1 scene, every stage against the oracle."""
    sc = pipeline.make_scene(side=72)
    w, h = 640, 480
    ca, cb = api.make_camera(*sc["cam_a"]), api.make_camera(*sc["cam_b"])
    oa, ob = oracle.make_camera(*sc["cam_a"]), oracle.make_camera(*sc["cam_b"])
    # depth maps: Model2Depth of the scene in each frame (bit-exact vs the oracle rasteriser)
    m2d = api.Model2Depth(ctx)
    m2d.SetInput([sc["verts_a"], sc["verts_b"]], [None, None], [sc["faces"], sc["faces"]], [[ca], [cb]])
    da, db = m2d.render(0)[0], m2d.render(1)[0]
    assert np.array_equal(da, oracle.model2depth(sc["verts_a"].astype(np.float32), sc["faces"], [oa])[0])
    assert (da > 0).mean() > 0.05 and (db > 0).mean() > 0.05
    assert MINMAX(da) and MINMAX(db)
    pa, va = api.unproject_depth(da, ca, pipeline.MIN_DSP, pipeline.MAX_DSP, ctx)
    pb, vb = api.unproject_depth(db, cb, pipeline.MIN_DSP, pipeline.MAX_DSP, ctx)
    matches = pipeline.lattice_matches(pa, va, sc["cam_b"], vb, sc["planted"], w, h, seed=0)
    assert len(matches) > 60
    out = pipeline.stitch_pair(da, ca, db, cb, matches, ctx, n_nodes=256, gn_iters=3, seed=3,
                               host_vectors=True)
    # ---- the device-resident form of the same pair (matched pixels unprojected on their own, target
    # unprojected + compacted where it is used) gives the same bits at every stage
    dev = pipeline.stitch_pair(da, ca, db, cb, matches, ctx, n_nodes=256, gn_iters=3, seed=3)
    assert dev["scale"] == out["scale"] and np.array_equal(dev["R"], out["R"])
    assert np.array_equal(dev["t"], out["t"]) and np.array_equal(dev["kept"], out["kept"])
    assert dev["err"] == out["err"] and dev["residual"] == out["residual"]
    assert np.array_equal(dev["template"], out["template"])
    assert np.array_equal(dev["deformed"], out["deformed"])
    assert np.array_equal(dev["render"], out["render"])
    uv = np.concatenate([matches[:, 0:2], [[0, 0], [w - 1, h - 1]]]).astype(np.int32)
    pp, pv = api.unproject_pixels(da, ca, uv, pipeline.MIN_DSP, pipeline.MAX_DSP, ctx)
    rows = uv[:, 1] * w + uv[:, 0]
    assert np.array_equal(pp, pa[rows]) and np.array_equal(pv, va[rows])
    with pytest.raises(_capi.MvsError):
        api.unproject_pixels(da, ca, [[w, 0]], pipeline.MIN_DSP, pipeline.MAX_DSP, ctx)
    dd = api.Deformation(out["template"], None, out["facets"], ctx)
    assert dd.SetTargetDepth(db, cb, pipeline.MIN_DSP, pipeline.MAX_DSP) == int(vb.sum())
    with pytest.raises(_capi.MvsError):  # no valid pixel: an error code, not an empty index
        dd.SetTargetDepth(np.zeros_like(db), cb, pipeline.MIN_DSP, pipeline.MAX_DSP)
    dd.close()
    # ---- oracle pipeline, stage by stage
    opa, ova = oracle.unproject_depth(da, oa, pipeline.MIN_DSP, pipeline.MAX_DSP)
    opb, ovb = oracle.unproject_depth(db, ob, pipeline.MIN_DSP, pipeline.MAX_DSP)
    assert np.array_equal(pa, opa) and np.array_equal(pb, opb)
    p = opa[matches[:, 1] * w + matches[:, 0]]
    q = opb[matches[:, 3] * w + matches[:, 2]]
    keep, err, _ = oracle.remove_outliers(p, q, oa, ob, 3, 200, 60.0, 0.75, 3)
    assert np.array_equal(out["kept"], keep) and abs(out["err"] - err) <= 1e-9 * max(err, 1.0)
    s, R, t = oracle.srt_estimate(p[keep], q[keep])
    assert abs(out["scale"] - s) <= 1e-6 * s and np.abs(out["R"] - R).max() <= 1e-6
    assert np.abs(out["t"] - t).max() <= 1e-6 * max(1.0, np.abs(t).max())
    ps, pR, pt = sc["planted"]
    assert abs(s - ps) < 2e-2 * ps and np.abs(R - pR).max() < 2e-2 and np.abs(t - pt).max() < 0.1
    ova_, otex, ofa = oracle.depth2model(da, oa, pipeline.MIN_DSP, pipeline.MAX_DSP, pipeline.SMOOTH_THRES)
    assert np.array_equal(out["facets"], ofa)
    tmpl, _ = oracle.apply_srt(ova_, None, s, R, t)
    assert np.allclose(out["template"], tmpl, rtol=0, atol=1e-5)
    # deformation onto the view-2 points (same template bits as the GPU run: cast like the library)
    pts32 = out["template"].astype(np.float32)
    tgt32 = out["target"].astype(np.float32)
    ids = oracle.sample_fps(pts32, 256, 0)
    widx, wts = oracle.node_weights(pts32, pts32[ids], 4)
    graph = oracle.node_graph(pts32[ids], 8)
    _, odef, _ = oracle.gn_solve(pts32, None, ids, widx, wts, graph,
                                 oracle.gn_params(gn_iters=3, pcg_iters=20, corr_mode=1), tpts=tgt32)
    diag = float(np.linalg.norm(pts32.max(0) - pts32.min(0)))
    assert np.abs(out["deformed"] - odef).max() <= 1e-4 * diag
    # the deformation pulls the transformed view-1 mesh onto the view-2 points
    from scipy.spatial import cKDTree
    tree = cKDTree(out["target"])
    assert tree.query(out["deformed"])[0].mean() < 0.7 * tree.query(out["template"])[0].mean()
    # re-projection (Model2Depth) of the deformed mesh agrees with view 2's own depth map
    r = out["render"]
    both = (r > 0) & (db > 0)
    assert both.sum() > 0.3 * (db > 0).sum()
    assert np.median(np.abs(r[both] - db[both]) / db[both]) < 1e-2
    oref = oracle.model2depth(odef.astype(np.float32), ofa, [ob])[0]
    agree = ((oref > 0) == (r > 0)).mean()
    assert agree > 0.999
    m = (oref > 0) & (r > 0)
    assert np.abs(oref[m] - r[m]).max() < 1e-3 * r[m].max()


def MINMAX(d):
    v = d[d > 0]
    return v.min() >= pipeline.MIN_DSP and v.max() <= pipeline.MAX_DSP


def _pair_inputs(ctx, u, side=48):
    sc = pipeline.make_scene(side=side, view_angle_deg=10.0 + 1.5 * u, seed=0)
    ca, cb = api.make_camera(*sc["cam_a"]), api.make_camera(*sc["cam_b"])
    m2d = api.Model2Depth(ctx)
    m2d.SetInput([sc["verts_a"], sc["verts_b"]], [None, None], [sc["faces"], sc["faces"]], [[ca], [cb]])
    da, db = m2d.render(0)[0], m2d.render(1)[0]
    pa, va = api.unproject_depth(da, ca, pipeline.MIN_DSP, pipeline.MAX_DSP, ctx)
    pb, vb = api.unproject_depth(db, cb, pipeline.MIN_DSP, pipeline.MAX_DSP, ctx)
    mt = pipeline.lattice_matches(pa, va, sc["cam_b"], vb, sc["planted"], 640, 480, seed=u)
    return da, ca, db, cb, mt, sc["planted"]


@pytest.mark.parametrize("workers", [1, 3])
def test_stitch_pairs_one_call_equals_stage_by_stage(ctx, workers):
    """mvs_stitch_pairs (Processor.cpp:629-826): all pairs in ONE library call, worked
    through by library threads with their own streams — every output carries the bits of the
    stage-by-stage sequence (itself checked against the oracle above), for any worker count."""
    pairs = [_pair_inputs(ctx, u) for u in range(4)]
    batch = api.PairBatch(ctx)
    for u, (da, ca, db, cb, mt, _) in enumerate(pairs):
        batch.add(da, ca, db, cb, mt, seed=u)
    outs = batch.run(api.pair_params(n_nodes=200, gn_iters=3, workers=workers))
    for u, ((da, ca, db, cb, mt, planted), o) in enumerate(zip(pairs, outs)):
        ref = pipeline.stitch_pair(da, ca, db, cb, mt, ctx, n_nodes=200, gn_iters=3, seed=u)
        assert o["scale"] == ref["scale"] and np.array_equal(o["R"], ref["R"])
        assert np.array_equal(o["t"], ref["t"]) and np.array_equal(o["kept"], ref["kept"])
        assert o["err"] == ref["err"] and o["residual"] == ref["residual"]
        assert np.array_equal(o["facets"], ref["facets"])
        assert np.array_equal(o["template"], ref["template"])
        assert np.array_equal(o["deformed"], ref["deformed"])
        assert np.array_equal(o["render"], ref["render"])
        assert o["n_target"] == int((db > 0).sum())
        ps, pR, _ = planted
        assert abs(o["scale"] - ps) < 2e-2 * ps and np.abs(o["R"] - pR).max() < 2e-2
    # a malformed job is an error code, not a crash
    bad = api.PairBatch(ctx)
    da, ca, db, cb, mt, _ = pairs[0]
    bad.add(np.zeros_like(da), ca, db, cb, mt)
    with pytest.raises(_capi.MvsError):
        bad.run(api.pair_params(workers=1))


def test_stitch_pairs_arena_fallback_and_growth(ctx):
    """The workers of mvs_stitch_pairs carve a pair's temporaries from a bump arena; a pair that does
    not fit falls back to the stream-ordered pool for the rest and the arena grows for the next pair.
    Same bits either way.  (Run in a subprocess: the arena size is read from the environment.)"""
    import os, subprocess, sys, textwrap
    code = textwrap.dedent("""
        import sys, numpy as np
        sys.path.insert(0, %r)
        from multiviewstitch_b200 import api
        from tests.test_gpu_pipeline import _pair_inputs
        ctx = api.Context(0)
        da, ca, db, cb, mt, _ = _pair_inputs(ctx, 0)
        outs = []
        for rep in range(3):
            b = api.PairBatch(ctx)
            b.add(da, ca, db, cb, mt, seed=1)
            o = b.run(api.pair_params(n_nodes=150, gn_iters=2, workers=1))[0]
            outs.append(o)
        for o in outs[1:]:
            assert np.array_equal(o["deformed"], outs[0]["deformed"]) and np.array_equal(o["render"], outs[0]["render"])
        np.save(sys.argv[1], outs[0]["deformed"])
    """ % str(ROOT_DIR))
    outs = []
    for mb in ("1", "96"):
        path = os.path.join(os.environ.get("TMPDIR", "/tmp"), f"mvs_arena_{mb}.npy")
        env = dict(os.environ, MVS_ARENA_MB=mb)
        r = subprocess.run([sys.executable, "-c", code, path], env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        outs.append(np.load(path))
    assert np.array_equal(outs[0], outs[1])
