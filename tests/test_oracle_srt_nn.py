"""CPU tests pinning the oracle (oracle/*.cpp) for NN search, 3x3 decompositions, camera and SRT.

The reference has no tests or golden vectors; the oracle is pinned by analytic
known-answer tests and independent numpy re-derivations.
"""
import numpy as np
import pytest

import oracle
from tests import synth


def numpy_knn(q, t, k):
    q = q.astype(np.float32)
    t = t.astype(np.float32)
    d = (q[:, None, :] - t[None, :, :]).astype(np.float32)
    dx, dy, dz = d[..., 0], d[..., 1], d[..., 2]
    d2 = ((dx * dx + dy * dy).astype(np.float32) + (dz * dz).astype(np.float32)).astype(np.float32)
    order = np.lexsort((np.broadcast_to(np.arange(len(t)), d2.shape), d2), axis=1)[:, :k]
    return order.astype(np.int32), np.take_along_axis(d2, order, 1)


@pytest.mark.parametrize("k", [1, 5, 9, 16])
def test_oracle_knn_matches_numpy(k):
    rng = np.random.default_rng(7)
    t = rng.normal(size=(700, 3)).astype(np.float32)
    q = rng.normal(size=(300, 3)).astype(np.float32)
    i_ref, d_ref = numpy_knn(q, t, k)
    i_b, d_b = oracle.knn(q, t, k, "brute")
    i_k, d_k = oracle.knn(q, t, k, "kdtree")
    assert np.array_equal(i_b, i_ref) and np.array_equal(d_b, d_ref)
    assert np.array_equal(i_k, i_ref) and np.array_equal(d_k, d_ref)


def test_oracle_knn_ties_lowest_index():
    # lattice: many exact distance ties; duplicates: identical points
    g = np.stack(np.meshgrid(*[np.arange(6)] * 3, indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    t = np.concatenate([g, g[:50]], 0)
    q = g + 0.5
    i_ref, d_ref = numpy_knn(q, t, 8)
    for m in ("brute", "kdtree"):
        i, d = oracle.knn(q, t, 8, m)
        assert np.array_equal(i, i_ref), m
        assert np.array_equal(d, d_ref), m


def test_oracle_knn_self_is_identity_and_k_gt_n():
    rng = np.random.default_rng(3)
    t = rng.normal(size=(100, 3)).astype(np.float32)
    i, d = oracle.knn(t, t, 1)
    assert np.array_equal(i[:, 0], np.arange(100)) and np.all(d == 0)
    i, d = oracle.knn(t[:2], t[:3], 5, "kdtree")
    assert np.all(i[:, 3:] == -1) and np.all(np.isinf(d[:, 3:]))


def test_oracle_svd_eig_vs_numpy():
    rng = np.random.default_rng(0)
    for trial in range(50):
        S = rng.normal(size=(3, 3))
        if trial % 5 == 0:
            S[:, 2] = S[:, 0] * 0.3  # rank deficient
        U, sig, V = oracle.svd3(S)
        assert np.allclose(U @ np.diag(sig) @ V.T, S, atol=1e-12)
        assert np.allclose(U.T @ U, np.eye(3), atol=1e-12) and np.allclose(V.T @ V, np.eye(3), atol=1e-12)
        assert np.allclose(sig, np.linalg.svd(S, compute_uv=False), atol=1e-12)
        A = S @ S.T
        w, Q = oracle.sym_eig3(A)
        assert np.allclose(w, np.linalg.eigvalsh(A), atol=1e-12)
        assert np.allclose(Q @ np.diag(w) @ Q.T, A, atol=1e-12)


def _cams():
    (K, R1, t1, w, h), (_, R2, t2, _, _) = synth.camera_pair()
    return oracle.make_camera(K, R1, t1, w, h), oracle.make_camera(K, R2, t2, w, h), (K, R1, t1, R2, t2)


def test_oracle_projection_rounding_truncates_toward_zero():
    # Camera.cpp:45-48: (int)(x + 0.5) truncates, so -0.7 -> (int)(-0.2) = 0, not -1
    K = np.array([[100.0, 0, 0.0], [0, 100.0, 0.0], [0, 0, 1]])
    cam = oracle.make_camera(K, np.eye(3), np.zeros(3), 640, 480)
    assert oracle.project(cam, [-0.007, -0.012, 1.0]) == (0, 0)  # -0.7+0.5=-0.2 -> 0 ; -1.2+.5=-0.7 -> 0
    assert oracle.project(cam, [-0.016, 0.0149, 1.0]) == (-1, 1)  # -1.6+.5=-1.1 -> -1 ; 1.49+.5=1.99 -> 1
    assert oracle.project(cam, [0.0151, 0.02, 2.0]) == (1, 1)


def test_oracle_unproject_project_roundtrip():
    c1, c2, _ = _cams()
    for (u, v, d) in [(10, 20, 4.0), (320, 240, 5.5), (639, 479, 6.0)]:
        for cam in (c1, c2):
            p = oracle.unproject(cam, u, v, d)
            assert oracle.project(cam, p) == (u, v)


def test_oracle_srt_recovers_planted_transform_noise_free():
    p, q = synth.planted_pairs(5000, noise=0.0)
    s, R, t = oracle.srt_estimate(p, q)
    assert abs(s - synth.PLANTED_S) < 1e-12
    assert np.allclose(R, synth.PLANTED_R, atol=1e-12)
    assert np.allclose(t, synth.PLANTED_T, atol=1e-11)
    assert abs(np.linalg.det(R) - 1) < 1e-12
    c1, c2, _ = _cams()
    assert oracle.srt_residual(p, q, c1, c2, s, R, t) < 0.02  # only pixel-rounding flips remain


def test_oracle_srt_matches_numpy_rederivation():
    p, q = synth.planted_pairs(3000, noise=1e-3)
    s, R, t = oracle.srt_estimate(p, q)
    c1, c2 = p.mean(0), q.mean(0)
    s_np = np.mean(np.linalg.norm(q - c2, axis=1) / np.linalg.norm(p - c1, axis=1))
    S = (s_np * (p - c1)).T @ (q - c2)
    U, _, Vt = np.linalg.svd(S)
    R_np = Vt.T @ U.T
    if abs(np.linalg.det(R_np) + 1) <= 1e-9:
        R_np = Vt.T @ np.diag([1, 1, -1]) @ U.T
    t_np = c2 - s_np * R_np @ c1
    assert abs(s - s_np) < 1e-12 * s_np
    assert np.allclose(R, R_np, atol=1e-11) and np.allclose(t, t_np, atol=1e-10)


def test_oracle_srt_reflection_fix():
    # mirrored correspondences force det(VU^T) = -1; the fix must return a proper rotation
    p, _ = synth.planted_pairs(500, noise=0.0)
    q = p * np.array([1.0, 1.0, -1.0])
    s, R, t = oracle.srt_estimate(p, q)
    assert abs(np.linalg.det(R) - 1) < 1e-9


def test_oracle_residual_numpy_rederivation():
    p, q = synth.planted_pairs(2000, noise=2e-3)
    c1, c2, (K, R1, t1, R2, t2) = _cams()
    s, R, t = oracle.srt_estimate(p, q)

    def proj(Rc, tc, x):
        pc = x @ Rc.T + tc
        u = np.trunc(K[0, 0] * pc[:, 0] / pc[:, 2] + K[0, 2] + 0.5)
        v = np.trunc(K[1, 1] * pc[:, 1] / pc[:, 2] + K[1, 2] + 0.5)
        return u, v

    tp = s * p @ R.T + t
    tq = (q - t) @ R / s
    u1, v1 = proj(R2, t2, tp); u2, v2 = proj(R2, t2, q)
    a1, b1 = proj(R1, t1, p); a2, b2 = proj(R1, t1, tq)
    e = np.mean(0.5 * (np.hypot(u1 - u2, v1 - v2) + np.hypot(a1 - a2, b1 - b2)))
    assert abs(oracle.srt_residual(p, q, c1, c2, s, R, t) - e) < 1e-9


def test_oracle_ransac_first_wins_and_outlier_robust():
    p, q = synth.planted_pairs(800, noise=0.0)
    rng = np.random.default_rng(5)
    out = rng.choice(800, 40, replace=False)
    q[out] += rng.normal(scale=0.5, size=(40, 3))
    c1, c2, _ = _cams()
    tr = oracle.draw_triples(800, 60, seed=11)
    assert np.all(tr[:, 0] < tr[:, 1]) and np.all(tr[:, 1] < tr[:, 2]) and tr.max() < 800
    tr[7] = tr[3]  # duplicate hypothesis: the earlier one must win ties
    s, R, t, e, best, errs = oracle.srt_ransac(p, q, c1, c2, tr)
    assert best == int(np.argmin(errs)) and e == errs.min()
    assert errs[7] == errs[3] and best != 7
    inl = np.setdiff1d(np.arange(800), out)
    assert oracle.srt_residual(p[inl], q[inl], c1, c2, s, R, t) < 25.0
