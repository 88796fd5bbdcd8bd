"""CPU tests pinning the as-rigid-as-possible oracle (oracle/arap.cpp).

CGAL 4.6 `Surface_mesh_deformation` (Deformation/Deformation.cpp:256-260,383-400) is not vendored
under /root/reference and the reference has no tests at that boundary, so the published algorithm
(Sorkine & Alexa 2007) is pinned by independent numpy/scipy statements: cotangent weights from
triangle angles, exact recovery of a rigid motion (zero energy), and one global step with identity
rotations against scipy's sparse direct solve of the constrained Laplace system.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import oracle
from tests import synth


def mesh(nu=20, nv=14):
    verts, faces = synth.uv_sphere(nu, nv)
    return verts.astype(np.float32), faces.astype(np.int32)


def numpy_cot_laplacian(pts, faces):
    p = pts.astype(np.float64)
    n = len(p)
    W = sp.lil_matrix((n, n))
    for f in faces:
        for c in range(3):
            o, a, b = f[c], f[(c + 1) % 3], f[(c + 2) % 3]
            u, w = p[a] - p[o], p[b] - p[o]
            cr = np.linalg.norm(np.cross(u, w))
            cot = np.dot(u, w) / cr if cr > 0 else 0.0
            W[a, b] += 0.5 * cot
            W[b, a] += 0.5 * cot
    W = W.tocsr()
    W.data = np.maximum(W.data, 0.0)
    return W


def test_cot_weights_match_numpy():
    pts, faces = mesh()
    rp, col, w = oracle.cot_weights(pts, faces)
    W = numpy_cot_laplacian(pts, faces)
    W.sort_indices()
    keep = np.diff(W.indptr)
    assert np.array_equal(rp, W.indptr)
    assert np.array_equal(col, W.indices)
    assert keep.sum() == len(w)
    # float32 storage of the half-cotangents
    assert np.allclose(w, W.data, rtol=2e-6, atol=1e-6)
    # symmetric
    M = sp.csr_matrix((w, col, rp), shape=W.shape)
    assert abs(M - M.T).max() < 1e-12


def test_rigid_motion_is_recovered_with_zero_energy():
    pts, faces = mesh()
    ids = oracle.sample_fps(pts, 12, 0)
    R = synth.axis_angle([1, 2, 3], 0.7)
    t = np.array([0.3, -0.2, 0.5])
    ctrl = pts[ids].astype(np.float64) @ R.T + t
    # alternating local/global minimisation converges linearly: 5 iterations leave 4e-2, 30 leave
    # 2e-4, 100 reach the float32 rounding of the control targets
    out, e, it = oracle.arap(pts, faces, ids, ctrl, iters=100, tol=0.0, cg_iters=400, cg_tol=1e-12)
    want = pts.astype(np.float64) @ R.T + t
    assert np.abs(out - want).max() < 1e-6
    assert e < 1e-9


def test_first_global_step_is_the_constrained_laplace_solve():
    pts, faces = mesh()
    n = len(pts)
    ids = oracle.sample_fps(pts, 16, 0)
    rng = np.random.default_rng(3)
    ctrl = pts[ids].astype(np.float64) + 0.1 * rng.normal(size=(len(ids), 3))
    out, _, it = oracle.arap(pts, faces, ids, ctrl, iters=1, tol=0.0, cg_iters=2000, cg_tol=1e-13)
    assert it == 1
    # numpy/scipy: L x = L p on the free rows, x = ctrl on the fixed rows (rotations = identity)
    W = numpy_cot_laplacian(pts, faces)
    rp, col, w = oracle.cot_weights(pts, faces)
    W = sp.csr_matrix((w, col, rp), shape=(n, n))  # the float32-rounded weights the solver uses
    Lm = sp.diags(np.asarray(W.sum(1)).ravel()) - W
    p = pts.astype(np.float64)
    fixed = np.zeros(n, bool)
    fixed[ids] = True
    xc = np.zeros((n, 3))
    xc[ids] = ctrl.astype(np.float32).astype(np.float64)
    free = ~fixed
    A = Lm[free][:, free].tocsc()
    rhs = (Lm @ p)[free] - Lm[free][:, fixed] @ xc[fixed]
    x = p.copy()
    x[fixed] = xc[fixed]
    x[free] = spla.spsolve(A, rhs)
    assert np.abs(out - x).max() < 1e-7


def test_energy_decreases_and_tolerance_stops_early():
    pts, faces = mesh()
    ids = oracle.sample_fps(pts, 16, 0)
    rng = np.random.default_rng(5)
    ctrl = pts[ids].astype(np.float64) * 1.1 + 0.05 * rng.normal(size=(len(ids), 3))
    energies = [oracle.arap(pts, faces, ids, ctrl, iters=k, tol=0.0, cg_iters=800, cg_tol=1e-12)[1]
                for k in (1, 2, 4, 8)]
    assert all(a >= b - 1e-12 for a, b in zip(energies, energies[1:]))
    _, _, it = oracle.arap(pts, faces, ids, ctrl, iters=50, tol=1e-3, cg_iters=800, cg_tol=1e-12)
    assert 1 < it < 50
