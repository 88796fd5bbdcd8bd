"""GPU parity tests: exact kNN and the SRT solver through the C ABI vs the CPU oracle.

Bars: kNN indices bit-exact with index tie-break (d2 bit-exact too);
SRT parameters within 1e-6 relative of the float64 oracle.
"""
import numpy as np
import pytest

import oracle
from multiviewstitch_b200 import api
from tests import synth

pytestmark = pytest.mark.gpu

SRT_RTOL = 1e-6


def _cams(mod):
    (K, R1, t1, w, h), (_, R2, t2, _, _) = synth.camera_pair()
    return mod.make_camera(K, R1, t1, w, h), mod.make_camera(K, R2, t2, w, h)


@pytest.mark.parametrize("k", [1, 4, 5, 9, 16, 20])
def test_knn_bit_exact_vs_oracle(ctx, k):
    rng = np.random.default_rng(k)
    t = rng.normal(size=(20000, 3)).astype(np.float32)
    q = rng.normal(size=(5000, 3)).astype(np.float32) * 1.3  # some queries outside the bbox
    idx = api.NNIndex(ctx).build(t)
    gi, gd = idx.knn(q, k)
    oi, od = oracle.knn(q, t, k, "kdtree")
    assert np.array_equal(gi, oi)
    assert np.array_equal(gd, od)


def test_knn_ties_duplicates_lattice(ctx):
    g = np.stack(np.meshgrid(*[np.arange(12)] * 3, indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    t = np.concatenate([g, g[:300]], 0)
    q = np.concatenate([g + 0.5, g], 0)
    gi, gd = api.nn_search(q, t, 8, ctx)
    oi, od = oracle.knn(q, t, 8, "brute")
    assert np.array_equal(gi, oi) and np.array_equal(gd, od)


def test_knn_surface_large_binned_queries(ctx):
    # > 16384 queries exercises the coarse query binning; sphere surface exercises the
    # occupancy-driven cell refinement
    t = synth.sphere_points(200000, 1, noise=1e-3).astype(np.float32)
    q = synth.sphere_points(60000, 2, noise=1e-3).astype(np.float32)
    gi, gd = api.nn_search(q, t, 1, ctx)
    oi, od = oracle.knn(q, t, 1, "kdtree")
    assert np.array_equal(gi, oi) and np.array_equal(gd, od)


def test_knn_edge_cases(ctx):
    rng = np.random.default_rng(0)
    t = rng.normal(size=(3, 3)).astype(np.float32)
    gi, gd = api.nn_search(t, t, 5, ctx)  # k > n: padded with -1 / inf
    oi, od = oracle.knn(t, t, 5)
    assert np.array_equal(gi, oi) and np.array_equal(gd, od)
    one = np.zeros((1, 3), np.float32)
    gi, gd = api.nn_search(rng.normal(size=(10, 3)), one, 1, ctx)  # single target
    assert np.all(gi == 0)
    same = np.ones((100, 3), np.float32)  # all coincident
    gi, gd = api.nn_search(same[:7], same, 3, ctx)
    assert np.array_equal(gi, np.tile([0, 1, 2], (7, 1))) and np.all(gd == 0)
    gi, gd = api.nn_search(np.zeros((0, 3), np.float32), t, 1, ctx)  # empty query set
    assert gi.shape == (0, 1)
    with pytest.raises(api.MvsError):
        api.nn_search(t, np.zeros((0, 3), np.float32), 1, ctx)  # empty target: error code, no exit()
    with pytest.raises(api.MvsError):
        api.nn_search(t, t, 33, ctx)


def test_part_recog_matches_oracle(ctx):
    rng = np.random.default_rng(4)
    tmpl = rng.normal(size=(9999, 3))
    labels = rng.integers(0, 16, size=9999).astype(np.int32)
    query = rng.normal(size=(30000, 3))
    pr = api.PartRecognition(ctx)
    pr.SetInput(tmpl)
    pr.SetLabels(labels)
    assert np.array_equal(pr.PartRecog(query), oracle.part_recog(tmpl, labels, query))


def _rel(a, b):
    return np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(1e-300, np.max(np.abs(b)))


@pytest.mark.parametrize("n", [3, 1000, 100003])
def test_srt_closed_form_f64(ctx, n):
    p, q = synth.planted_pairs(n, noise=1e-3)
    solver = api.SRTSolver(200, ctx)
    solver.SetPrintFlag(False)
    solver.SetInput(p, q)
    s, R, t = solver.EstimateTransform()
    so, Ro, to = oracle.srt_estimate(p, q)
    assert abs(s - so) <= SRT_RTOL * so
    assert _rel(R, Ro) <= SRT_RTOL and _rel(t, to) <= SRT_RTOL
    assert abs(np.linalg.det(R) - 1) < 1e-9


def test_srt_closed_form_f32_float4(ctx):
    p, q = synth.planted_pairs(250000, noise=1e-3)
    p4 = np.zeros((len(p), 4), np.float32); p4[:, :3] = p
    q4 = np.zeros((len(q), 4), np.float32); q4[:, :3] = q
    solver = api.SRTSolver(200, ctx)
    solver.SetPrintFlag(False)
    solver.SetInput(p4, q4)
    s, R, t = solver.EstimateTransform()
    # oracle on the SAME float32-rounded inputs
    so, Ro, to = oracle.srt_estimate(p4[:, :3].astype(np.float64), q4[:, :3].astype(np.float64))
    assert abs(s - so) <= SRT_RTOL * so and _rel(R, Ro) <= SRT_RTOL and _rel(t, to) <= SRT_RTOL


@pytest.mark.parametrize("unit", [1e-22, 1e-9, 1.0, 1e9, 1e20])
def test_srt_scale_norm_ratio_ranges(ctx, unit):
    """The mean of |q-c2|/|p-c1| (SRTSolver.cpp:42-45) is evaluated with float32-seeded Newton
    steps inside the float32 range of the squared norms and with plain float64 sqrt/div outside it
    (units of 1e-22 and 1e20 here): both must agree with the oracle to ~1e-12, far inside the bar."""
    p, q = synth.planted_pairs(20000, noise=1e-3)
    p, q = p * unit, q * unit
    solver = api.SRTSolver(10, ctx)
    solver.SetPrintFlag(False)
    solver.SetInput(p, q)
    s, R, t = solver.EstimateTransform()
    so, Ro, to = oracle.srt_estimate(p, q)
    assert abs(s - so) <= 1e-12 * so
    assert _rel(R, Ro) <= SRT_RTOL and _rel(t, to) <= SRT_RTOL


def test_srt_point_at_centroid_behaves_like_reference(ctx):
    """SRTSolver.cpp:44 divides by |p_i - centroid| without a guard: a point that IS the centroid
    makes the scale non-finite in the reference; the device path keeps that (no silent clamp)."""
    p = np.array([[-1, 0, 0], [1, 0, 0], [0, -1, 0], [0, 1, 0], [0, 0, -1], [0, 0, 1], [0, 0, 0]], float)
    q = 2.0 * p + 1.0
    solver = api.SRTSolver(10, ctx)
    solver.SetPrintFlag(False)
    solver.SetInput(p, q)
    s, R, t = solver.EstimateTransform()
    so, _, _ = oracle.srt_estimate(p, q)
    assert not np.isfinite(so) and not np.isfinite(s)


def test_srt_planted_recovered_and_reflection(ctx):
    p, q = synth.planted_pairs(4096, noise=0.0)
    solver = api.SRTSolver(10, ctx)
    solver.SetPrintFlag(False)
    solver.SetInput(p, q)
    s, R, t = solver.EstimateTransform()
    assert abs(s - synth.PLANTED_S) < 1e-11 and np.allclose(R, synth.PLANTED_R, atol=1e-11)
    assert np.allclose(t, synth.PLANTED_T, atol=1e-10)
    solver.SetInput(p, p * np.array([1.0, 1.0, -1.0]))
    s, R, t = solver.EstimateTransform()
    assert abs(np.linalg.det(R) - 1) < 1e-9


def test_srt_residual_matches_oracle(ctx):
    p, q = synth.planted_pairs(50000, noise=2e-3)
    c1, c2 = _cams(api)
    o1, o2 = _cams(oracle)
    solver = api.SRTSolver(1, ctx)
    solver.SetPrintFlag(False)
    solver.SetInput(p, q, c1, c2)
    so, Ro, to = oracle.srt_estimate(p, q)
    e = solver.ResidualError(so, Ro, to)
    eo = oracle.srt_residual(p, q, o1, o2, so, Ro, to)
    assert abs(e - eo) <= 1e-12 * max(1.0, eo)


def test_srt_ransac_same_triples_same_answer(ctx):
    n = 20000
    p, q = synth.planted_pairs(n, noise=1e-3)
    rng = np.random.default_rng(9)
    out = rng.choice(n, n // 20, replace=False)
    q[out] += rng.normal(scale=0.3, size=(len(out), 3))
    c1, c2 = _cams(api)
    o1, o2 = _cams(oracle)
    tr = api.srt_draw_triples(n, 200, seed=42)
    assert np.array_equal(tr, oracle.draw_triples(n, 200, 42))
    tr[17] = tr[5]
    solver = api.SRTSolver(200, ctx)
    solver.SetPrintFlag(False)
    solver.SetInput(p, q, c1, c2)
    s, R, t, err, best = solver.EstimateTransformRansac(tr, full=True)
    so, Ro, to, eo, besto, errs = oracle.srt_ransac(p, q, o1, o2, tr)
    assert best == besto
    assert abs(err - eo) <= 1e-9 * max(1.0, eo)
    assert abs(s - so) <= SRT_RTOL * so and _rel(R, Ro) <= SRT_RTOL and _rel(t, to) <= SRT_RTOL
    # seeded internal sampling = the published triples
    tr42 = api.srt_draw_triples(n, 200, seed=42)
    solver.SetIterationNum(200)
    s2, R2, t2, err2, best2 = solver.EstimateTransformRansac(None, seed=42, full=True)
    s3, R3, t3, err3, best3 = solver.EstimateTransformRansac(tr42, full=True)
    assert best2 == best3 and err2 == err3 and np.array_equal(R2, R3)


def test_srt_error_codes(ctx):
    solver = api.SRTSolver(10, ctx)
    solver.SetPrintFlag(False)
    with pytest.raises(api.MvsError):
        solver.EstimateTransform()  # before SetInput
    p, q = synth.planted_pairs(10, noise=0.0)
    solver.SetInput(p, q)  # no cameras
    with pytest.raises(api.MvsError):
        solver.ResidualError(1.0, np.eye(3), np.zeros(3))
    solver.SetInput(p[:0], q[:0])
    with pytest.raises(api.MvsError):
        solver.EstimateTransform()
