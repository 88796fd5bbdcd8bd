"""Golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py): the oracle must still
reproduce them (CPU), and the CUDA path must match them through the C ABI (GPU)."""
from pathlib import Path

import numpy as np
import pytest

import oracle
from tests import synth

G = Path(__file__).resolve().parent / "golden"


def _cams(mod):
    (K, R1, t1, w, h), (_, R2, t2, _, _) = synth.camera_pair()
    return mod.make_camera(K, R1, t1, w, h), mod.make_camera(K, R2, t2, w, h)


def _m2d_cams(mod, g):
    return [mod.make_camera(g["K"], c[:9].reshape(3, 3), c[9:], 96, 72) for c in g["cams"]]


# ------------------------------------------------------------------ CPU: oracle vs golden
def test_oracle_arap_golden():
    g = np.load(G / "arap.npz")
    rp, col, w = oracle.cot_weights(g["pts"], g["faces"])
    assert np.array_equal(rp, g["row_ptr"]) and np.array_equal(col, g["col"]) and np.array_equal(w, g["w"])
    out, energy, it = oracle.arap(g["pts"], g["faces"], g["ids"], g["ctrl"], iters=4, tol=0.0,
                                  cg_iters=300, cg_tol=0.0)
    assert it == 4 and np.array_equal(out, g["out"]) and energy == g["energy"]


def test_oracle_knn_golden():
    g = np.load(G / "knn.npz")
    for method in ("brute", "kdtree"):
        idx, d2 = oracle.knn(g["q"], g["t"], 5, method)
        assert np.array_equal(idx, g["idx"]) and np.array_equal(d2, g["d2"])


def test_oracle_srt_golden():
    g = np.load(G / "srt.npz")
    c1, c2 = _cams(oracle)
    s, R, t = oracle.srt_estimate(g["p"], g["q"])
    assert s == g["s"] and np.array_equal(R, g["R"]) and np.array_equal(t, g["t"])
    assert oracle.srt_residual(g["p"], g["q"], c1, c2, s, R, t) == g["residual"]
    rs, rR, rt, rerr, rbest, errs = oracle.srt_ransac(g["p"], g["q"], c1, c2, g["triples"])
    assert rbest == g["rbest"] and np.array_equal(errs, g["errs"]) and np.array_equal(rR, g["rR"])
    assert np.array_equal(oracle.draw_triples(300, 50, 7), g["triples"])


def test_oracle_deform_golden():
    g = np.load(G / "deform.npz")
    pts = g["pts"]
    assert np.array_equal(oracle.sample_fps(pts, 60, 0), g["fps"])
    assert np.array_equal(oracle.uniform_sampling(pts, 16), g["uniform"])
    widx, w = oracle.node_weights(pts, pts[g["fps"]], 4)
    assert np.array_equal(widx, g["widx"]) and np.array_equal(w, g["w"])
    assert np.array_equal(oracle.node_graph(pts[g["fps"]], 8), g["graph"])
    x, out, en = oracle.gn_solve(pts, None, g["fps"], widx, w, g["graph"],
                                 oracle.gn_params(gn_iters=6, pcg_iters=20), targets=g["targets"])
    assert np.allclose(x, g["x"], atol=1e-12) and np.allclose(out, g["out"], atol=1e-12)


def test_oracle_model2depth_and_align_golden():
    g = np.load(G / "model2depth.npz")
    assert np.array_equal(oracle.model2depth(g["v"], g["f"], _m2d_cams(oracle, g)), g["img"])
    a = np.load(G / "align.npz")
    out = oracle.align(a["src"], a["src"].copy(), a["sl"], a["tgt"], a["tgt"].copy(),
                       np.zeros((0, 3), np.int32), a["view"], remove_ground=0, local_align=1,
                       ground_ray=a["ground"])
    assert np.allclose(out[0], a["out_src"], atol=1e-12) and np.array_equal(out[5], a["out_tl"])


# ------------------------------------------------------------------ GPU: C ABI vs golden
@pytest.mark.gpu
def test_gpu_knn_golden(ctx):
    from multiviewstitch_b200 import api
    g = np.load(G / "knn.npz")
    idx, d2 = api.nn_search(g["q"], g["t"], 5, ctx)
    assert np.array_equal(idx, g["idx"]) and np.array_equal(d2, g["d2"])
    i1, d1 = api.nn_search(g["q"], g["t"], 1, ctx)  # the 1-NN path walks the occupancy pyramid
    assert np.array_equal(i1[:, 0], g["idx"][:, 0]) and np.array_equal(d1[:, 0], g["d2"][:, 0])


@pytest.mark.gpu
def test_gpu_srt_golden(ctx):
    from multiviewstitch_b200 import api
    g = np.load(G / "srt.npz")
    c1, c2 = _cams(api)
    s = api.SRTSolver(50, ctx)
    s.SetInput(g["p"], g["q"], c1, c2)
    sc, R, t = s.EstimateTransform()
    assert abs(sc - g["s"]) <= 1e-6 * g["s"] and np.abs(R - g["R"]).max() <= 1e-6
    assert np.abs(t - g["t"]).max() <= 1e-6 * np.abs(g["t"]).max()
    assert abs(s.ResidualError(g["s"], g["R"], g["t"]) - g["residual"]) <= 1e-9 * g["residual"]
    rs, rR, rt, rerr, rbest = s.EstimateTransformRansac(g["triples"], full=True)
    assert rbest == g["rbest"] and np.abs(rR - g["rR"]).max() <= 1e-6
    assert np.array_equal(api.srt_draw_triples(300, 50, 7), g["triples"])


@pytest.mark.gpu
def test_gpu_deform_golden(ctx):
    from multiviewstitch_b200 import api
    g = np.load(G / "deform.npz")
    pts = g["pts"]
    diag = float(np.linalg.norm(pts.max(0) - pts.min(0)))
    d = api.Deformation(pts, None, g["faces"], ctx)
    assert np.array_equal(d.UniformSampling(16), g["uniform"])
    assert np.array_equal(d.SampleFPS(60, 0), g["fps"])
    assert np.array_equal(d.KNearestNeighbor(8), g["graph"])
    widx, w = d.Weights()
    assert np.array_equal(widx, g["widx"]) and np.array_equal(w, g["w"])
    d.SetTargets(g["targets"])
    d.GNBegin(api.gn_params(gn_iters=6, pcg_iters=20, corr_mode=2))
    for _ in range(6):
        d.GNStep()
    d.GNEnd()
    assert np.abs(d.Vertices() - g["out"]).max() <= 1e-4 * diag
    d.SetTarget(g["tv"])
    d.GNBegin(api.gn_params(gn_iters=4, pcg_iters=20, corr_mode=1))
    for _ in range(4):
        d.GNStep()
    d.GNEnd()
    assert np.abs(d.Vertices() - g["out_nn"]).max() <= 1e-4 * diag


@pytest.mark.gpu
def test_gpu_arap_golden(ctx):
    from multiviewstitch_b200 import api
    g = np.load(G / "arap.npz")
    pts = g["pts"]
    diag = float(np.linalg.norm(pts.max(0) - pts.min(0)))
    d = api.Deformation(pts, None, g["faces"], ctx)
    assert np.array_equal(d.SampleFPS(30, 0), g["ids"])
    e, it = d.ARAP(g["ctrl"], api.arap_params(iters=4, tol=0.0, cg_iters=300, cg_tol=0.0))
    assert it == 4
    assert np.abs(d.Vertices() - g["out"]).max() <= 1e-4 * diag
    assert abs(e - float(g["energy"])) <= 1e-3 * float(g["energy"])


@pytest.mark.gpu
def test_gpu_model2depth_and_align_golden(ctx):
    from multiviewstitch_b200 import api
    g = np.load(G / "model2depth.npz")
    m2d = api.Model2Depth(ctx)
    m2d.SetInput([g["v"]], [None], [g["f"]], [_m2d_cams(api, g)])
    assert np.array_equal(m2d.render(0), g["img"])
    a = np.load(G / "align.npz")
    al = api.Alignment(ctx, api.align_params(remove_ground=0, local_align=1))
    out = al.Align(a["src"], a["src"].copy(), a["sl"], a["tgt"], a["tgt"].copy(),
                   np.zeros((0, 3), np.int32), a["view"], ground_ray=a["ground"])
    assert np.allclose(out[0], a["out_src"], atol=1e-8) and np.array_equal(out[5], a["out_tl"])
