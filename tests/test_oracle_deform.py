"""CPU tests pinning the deformation-graph oracle (oracle/deform.cpp).

No reference lines exist for the Gauss-Newton solve (the reference hands the nodes to CGAL ARAP),
so the formulation of SURVEY.md §8(a) D7-D9 is pinned here by an INDEPENDENT numpy statement of the
energy: finite-difference gradient/Gauss-Newton Hessian, and scipy's sparse direct solve for PCG.
"""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import oracle
from tests import synth


def small_problem(nu=24, nv=16, m=24, seed=0):
    verts, faces = synth.uv_sphere(nu, nv)
    pts = verts.astype(np.float32)
    ids = oracle.sample_fps(pts, m, 0)
    nodes = pts[ids]
    widx, w = oracle.node_weights(pts, nodes, 4)
    graph = oracle.node_graph(nodes, 8)
    return pts, faces, ids, nodes, widx, w, graph


def numpy_residuals(x, pts, nodes, widx, w, graph, targets, tw, prm):
    """Stacked weighted residual vector of the D7 energy, written independently in numpy."""
    m = len(nodes)
    X = x.reshape(m, 12)
    A = np.stack([X[:, 0:3], X[:, 3:6], X[:, 6:9]], axis=2)  # A[j][:, c] = column c
    t = X[:, 9:12]
    p = pts.astype(np.float64)
    g = nodes.astype(np.float64)
    vt = np.zeros_like(p)
    for a in range(4):
        j = widx[:, a]
        d = p - g[j]
        vt += w[:, a:a + 1].astype(np.float64) * (np.einsum("nrc,nc->nr", A[j], d) + g[j] + t[j])
    om = np.ones(len(p)) if tw is None else tw.astype(np.float64)
    r_fit = (np.sqrt(prm.w_fit * om)[:, None] * (vt - targets)).ravel()
    a1, a2, a3 = A[:, :, 0], A[:, :, 1], A[:, :, 2]
    dot = lambda u, v: np.sum(u * v, 1)
    r_rot = np.sqrt(prm.w_rot) * np.stack(
        [dot(a1, a2), dot(a1, a3), dot(a2, a3), dot(a1, a1) - 1, dot(a2, a2) - 1, dot(a3, a3) - 1], 1).ravel()
    r_reg = []
    for j in range(m):
        for l in graph[j]:
            if l < 0 or l == j:
                continue
            e = g[l] - g[j]
            r_reg.append(np.sqrt(prm.w_reg) * (A[j] @ e + g[j] + t[j] - g[l] - t[l]))
    return np.concatenate([r_fit, r_rot, np.concatenate(r_reg)])


def test_fps_and_uniform_sampling_properties():
    pts, *_ = small_problem()
    ids = oracle.sample_fps(pts, 20, 3)
    assert ids[0] == 3 and len(set(ids.tolist())) == 20
    # numpy re-derivation of FPS with float32 distances
    mind = np.full(len(pts), np.inf, np.float32)
    cur, ref = 3, []
    for _ in range(20):
        ref.append(cur)
        d = pts - pts[cur]
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]).astype(np.float32) + (d[:, 2] * d[:, 2]).astype(np.float32)
        mind = np.minimum(mind, d2.astype(np.float32))
        cur = int(np.argmax(mind))
    assert ids.tolist() == ref
    # Deformation.cpp:63-106: kept nodes are never within each other's 16-NN *at selection time*
    s = oracle.uniform_sampling(pts, 16)
    assert s[0] == 0 and np.all(np.diff(s) > 0)
    nn, _ = oracle.knn(pts, pts, 16, "kdtree")
    removed = np.zeros(len(pts), bool)
    ref = []
    for i in range(len(pts)):
        if not removed[i]:
            ref.append(i)
            removed[nn[i][nn[i] != i]] = True
    assert s.tolist() == ref


def test_node_weights_properties():
    pts, faces, ids, nodes, widx, w, graph = small_problem()
    assert np.allclose(w.sum(1), 1.0, atol=1e-6) and np.all(w >= 0)
    i5, d5 = oracle.knn(pts, nodes, 5, "kdtree")
    assert np.array_equal(widx, i5[:, :4])
    d = np.sqrt(d5)
    wr = (1 - d[:, :4] / d[:, 4:5]) ** 2
    wr /= wr.sum(1, keepdims=True)
    assert np.allclose(w, wr, atol=1e-5)
    # a node's own vertex has itself as nearest node
    assert np.array_equal(widx[ids, 0], np.arange(len(ids)))
    assert np.array_equal(graph[:, 0], np.arange(len(ids)))  # Deformation.cpp:108-153 includes self


def test_gn_normal_equations_match_finite_differences():
    pts, faces, ids, nodes, widx, w, graph = small_problem(m=12)
    rng = np.random.default_rng(1)
    m = len(ids)
    prm = oracle.gn_params(w_rot=1.0, w_reg=10.0, w_fit=100.0, lam=0.0)
    x = np.tile(np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0.0]), (m, 1)) + 0.05 * rng.normal(size=(m, 12))
    targets = synth.bend_twist(pts.astype(np.float64), 0.05)
    tw = (rng.random(len(pts)) > 0.2).astype(np.float32)
    H, g = oracle.gn_assemble_dense(pts, ids, widx, w, graph, targets, tw, prm, x)
    f = lambda xx: numpy_residuals(xx, pts, nodes, widx, w, graph, targets, tw, prm)
    r0 = f(x.ravel())
    J = np.empty((len(r0), 12 * m))
    eps = 1e-6
    for k in range(12 * m):
        xp = x.ravel().copy(); xp[k] += eps
        xm = x.ravel().copy(); xm[k] -= eps
        J[:, k] = (f(xp) - f(xm)) / (2 * eps)
    assert np.allclose(H, J.T @ J, rtol=1e-6, atol=1e-6 * np.abs(H).max())
    assert np.allclose(g, J.T @ r0, rtol=1e-6, atol=1e-6 * np.abs(g).max())
    assert np.allclose(H, H.T, atol=1e-12 * np.abs(H).max())


def test_gn_pcg_converges_to_direct_solve_and_energy_decreases():
    pts, faces, ids, nodes, widx, w, graph = small_problem(m=24)
    m = len(ids)
    targets = synth.bend_twist(pts.astype(np.float64), 0.05)
    x0 = np.tile(np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0.0]), (m, 1))
    prm = oracle.gn_params(gn_iters=1, pcg_iters=400, lam=1e-6)
    H, g = oracle.gn_assemble_dense(pts, ids, widx, w, graph, targets, None, prm, x0)
    delta = spla.spsolve(sp.csc_matrix(H), -g)
    x1, out, en = oracle.gn_solve(pts, None, ids, widx, w, graph, prm, x0, targets=targets)
    assert np.allclose(x1.ravel() - x0.ravel(), delta, rtol=1e-6, atol=1e-8)
    # energies over GN iterations decrease monotonically with enough PCG iterations
    energies = []
    for it in range(1, 5):
        prm = oracle.gn_params(gn_iters=it, pcg_iters=200)
        _, _, e = oracle.gn_solve(pts, None, ids, widx, w, graph, prm, x0, targets=targets)
        energies.append(e.sum())
    assert all(b <= a * (1 + 1e-9) for a, b in zip(energies, energies[1:]))
    # identity target => zero update
    prm = oracle.gn_params(gn_iters=2, pcg_iters=20)
    xi, outi, ei = oracle.gn_solve(pts, None, ids, widx, w, graph, prm, x0, targets=pts.astype(np.float64))
    assert np.allclose(xi, x0, atol=1e-6) and np.allclose(outi, pts, atol=1e-6)  # float32 weights sum to 1 within 1e-7


def test_gn_recovers_rigid_motion():
    pts, faces, ids, nodes, widx, w, graph = small_problem(m=24)
    R = synth.axis_angle([0.2, 1.0, 0.1], 0.3)
    tt = np.array([0.1, -0.05, 0.2])
    targets = pts.astype(np.float64) @ R.T + tt
    prm = oracle.gn_params(gn_iters=8, pcg_iters=200)
    x, out, e = oracle.gn_solve(pts, None, ids, widx, w, graph, prm, targets=targets)
    assert np.abs(out - targets).max() < 1e-4
    A = x.reshape(-1, 12)[:, :9].reshape(-1, 3, 3).transpose(0, 2, 1)  # columns -> matrix
    assert np.allclose(A, R[None], atol=1e-3)


def test_node_targets_rules():
    # template = sphere r=1, target = sphere r=1.1 (denser): each node's target lies radially out
    verts, faces = synth.uv_sphere(40, 30, bumps=0.0)
    nodes = verts[::37].astype(np.float32)
    nn = nodes / np.linalg.norm(nodes, axis=1, keepdims=True)
    tv, tf = synth.uv_sphere(120, 90, radius=1.1, bumps=0.0)
    tn = tv / np.linalg.norm(tv, axis=1, keepdims=True)
    graph = oracle.node_graph(nodes, 8)
    ctrl, valid = oracle.node_targets(nodes, nn, tv, tn, graph, 100.0, 100.0, smooth_sweeps=0)
    assert valid.mean() > 0.9
    v = valid.astype(bool)
    assert np.allclose(np.linalg.norm(ctrl[v], axis=1), 1.1, atol=0.02)
    cosang = np.sum(ctrl[v] * nn[v], 1) / np.linalg.norm(ctrl[v], axis=1)
    assert np.all(cosang > 0.99)
    # normal compatibility (Deformation.cpp:306-312): flipped target normals => nothing valid
    ctrl2, valid2 = oracle.node_targets(nodes, nn, tv, -tn, graph, 100.0, 100.0, 0)
    assert valid2.sum() == 0 and np.allclose(ctrl2, nodes)
    # thresholds (:350)
    _, valid3 = oracle.node_targets(nodes, nn, tv, tn, graph, 0.05, 100.0, 0)
    assert valid3.sum() == 0
    # smoothing (:359-381): numpy re-derivation of the 2 Jacobi sweeps
    ctrl0, _ = oracle.node_targets(nodes, nn, tv, tn, graph, 100.0, 100.0, 0)
    ctrl_s, _ = oracle.node_targets(nodes, nn, tv, tn, graph, 100.0, 100.0, 2)
    c = ctrl0.copy(); o = nodes.astype(np.float64)
    for _ in range(2):
        c = o + (c - o)[graph].sum(1) / 9.0
    assert np.allclose(ctrl_s, c, atol=1e-12)


def test_vertex_normals_sphere():
    verts, faces = synth.uv_sphere(48, 32, bumps=0.0)
    n = oracle.vertex_normals(verts, faces)
    radial = verts / np.linalg.norm(verts, axis=1, keepdims=True)
    assert np.all(np.sum(n * radial, 1) > 0.99)
    assert np.allclose(n, synth.vertex_normals(verts, faces), atol=1e-12)


def test_pipelined_pcg_recurrences_track_standard_pcg():
    """Pins the plan of DESIGN.md §8 (one exchange per PCG iteration): the pipelined recurrences of
    Ghysels & Vanroose reproduce the oracle's standard PCG iterates in float64; with float32 vectors
    they drift far more than standard float32 PCG does; with float64 recurrences and a float64 halo
    (float32 matrix and preconditioner) they are at least as close to the oracle as the float32
    standard kernel.  Real normal equations of the embedded-deformation solve, 20 iterations."""
    verts, faces = synth.cube_sphere(24)
    pts = verts.astype(np.float32)
    m = 200
    ids = oracle.sample_fps(pts, m, 0)
    widx, w = oracle.node_weights(pts, pts[ids], 4)
    graph = oracle.node_graph(pts[ids], 8)
    prm = oracle.gn_params(gn_iters=1, pcg_iters=20)
    x0 = np.tile(np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0], np.float64), (m, 1))
    H, g = oracle.gn_assemble_dense(pts, ids, widx, w, graph, synth.bend_twist(verts, 0.05), None, prm, x0)
    H = H + prm.lam * np.eye(len(g)) if hasattr(prm, "lam") else H
    b = (-g).astype(np.float32).astype(np.float64)  # the device holds the gradient in float32
    N = len(b)
    Minv = np.zeros_like(H)
    for j in range(N // 12):
        sl = slice(12 * j, 12 * j + 12)
        Minv[sl, sl] = np.linalg.inv(H[sl, sl])
    d64 = lambda a, c: float(np.dot(a.astype(np.float64), c.astype(np.float64)))

    def standard(dt):
        A, Mi = H.astype(dt), Minv.astype(dt)
        x = np.zeros(N, dt); r = b.astype(dt); z = Mi @ r; p = z.copy(); rz = d64(r, z)
        for _ in range(20):
            q = A @ p
            a = dt(rz / d64(p, q))
            x = x + a * p; r = r - a * q; z = Mi @ r
            rzn = d64(r, z); p = z + dt(rzn / rz) * p; rz = rzn
        return x.astype(np.float64)

    def pipelined(vec, mat, halo):
        A, Mi = H.astype(mat).astype(vec), Minv.astype(mat).astype(vec)
        x = np.zeros(N, vec); r = b.astype(vec); u = Mi @ r
        wv = A @ u.astype(halo).astype(vec)
        z = np.zeros(N, vec); q = z.copy(); s = z.copy(); p = z.copy()
        go = ao = 1.0
        for it in range(20):
            gam, dl = d64(r, u), d64(wv, u)
            mm = Mi @ wv
            nn = A @ mm.astype(halo).astype(vec)
            be = gam / go if it else 0.0
            al = gam / (dl - be * gam / ao) if it else gam / dl
            be_, al_ = vec(be), vec(al)
            z = nn + be_ * z; q = mm + be_ * q; s = wv + be_ * s; p = u + be_ * p
            x = x + al_ * p; r = r - al_ * s; u = u - al_ * q; wv = wv - al_ * z
            go, ao = gam, al
        return x.astype(np.float64)

    ref = standard(np.float64)
    rel = lambda x: np.linalg.norm(x - ref) / np.linalg.norm(ref)
    e_std32 = rel(standard(np.float32))
    e_p64 = rel(pipelined(np.float64, np.float64, np.float64))
    e_p32 = rel(pipelined(np.float32, np.float32, np.float32))
    e_mixed = rel(pipelined(np.float64, np.float32, np.float64))
    assert e_p64 < 1e-10                       # same iterates in exact arithmetic
    assert e_p32 > 5 * e_std32                 # float32 recurrences drift
    assert e_mixed < 5e-6 and e_mixed < 5 * max(e_std32, 1e-7)   # float64 recurrences + halo do not


def test_openmp_courtesy_build_matches_single_thread():
    """bench.py reports a courtesy multi-thread CPU number from the same oracle sources compiled with
    -fopenmp; its results must agree with the single-thread oracle (summation order only)."""
    if oracle.lib_mt() is None:
        pytest.skip("no libgomp in this toolchain")
    verts, faces = synth.cube_sphere(20)
    pts = verts.astype(np.float32)
    ids = oracle.sample_fps(pts, 60, 0)
    widx, w = oracle.node_weights(pts, pts[ids], 4)
    graph = oracle.node_graph(pts[ids], 8)
    tv = synth.bend_twist(synth.cube_sphere(22, rot=synth.axis_angle([1, 2, 3], 0.7))[0], 0.04).astype(np.float32)
    prm = oracle.gn_params(gn_iters=3, pcg_iters=15, corr_mode=1)
    x1, o1, e1 = oracle.gn_solve(pts, None, ids, widx, w, graph, prm, tpts=tv)
    x2, o2, e2 = oracle.gn_solve(pts, None, ids, widx, w, graph, prm, tpts=tv, mt=True)
    assert np.abs(o1 - o2).max() < 1e-10 and np.abs(x1 - x2).max() < 1e-10


def test_query_point_float32_order_tracks_float64():
    """The correspondence query point (float32, fixed rounding order shared with the device) is the
    float64 deformed vertex to float32 accuracy, for identity and perturbed node states."""
    verts, faces = synth.uv_sphere(48, 36)
    pts = verts.astype(np.float32)
    ids = oracle.sample_fps(pts, 60, 0)
    widx, w = oracle.node_weights(pts, pts[ids], 4)
    rng = np.random.default_rng(3)
    ident = np.tile(np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0.0]), (60, 1))
    for x in (ident, ident + 0.05 * rng.normal(size=ident.shape)):
        xf = x.astype(np.float32).astype(np.float64)  # the device state is float32
        q, idx, d2 = oracle.gn_correspondences(pts, ids, widx, w, xf, pts)
        want = oracle.deform_vertices(pts, pts[ids], widx, w, xf)
        assert np.abs(q - want).max() < 2e-6
        # exact 1-NN of the query in the target, lowest index on ties (brute force)
        bi, bd = oracle.knn(q, pts, 1, "brute")
        assert np.array_equal(idx, bi[:, 0]) and np.array_equal(d2, bd[:, 0])
    # identity state: every vertex is its own query up to rounding of w-weighted sums
    q, idx, _ = oracle.gn_correspondences(pts, ids, widx, w, ident, pts)
    assert (idx == np.arange(len(pts))).mean() > 0.99


def _literal_vs_bucketed(nodes, nn, tv, tn, graph):
    try:
        c0, v0 = oracle.node_targets(nodes, nn, tv, tn, graph, smooth_sweeps=0)
        oracle.set_literal_comparator(True)
        c1, v1 = oracle.node_targets(nodes, nn, tv, tn, graph, smooth_sweeps=0)
    finally:
        oracle.set_literal_comparator(False)
    return c0, v0, c1, v1


def test_literal_reference_comparator_measured_against_bucketed_rule():
    """the node-correspondence candidates are ordered by floor(projDist / 1e-6) buckets,
    the reference by a pairwise |projDist - other| <= 1e-6 priority-queue comparator that is not a
    strict weak ordering (Deformation.cpp:318-327; its pop order depends on the heap algorithm).
    Measured: on a generic scan (asymmetric tessellation, warped) the two pick the same 8 neighbours
    for EVERY node; on symmetric data with rings of equal projDist (concentric UV spheres, a lattice
    plane under a node) a few nodes pick different members of a tied ring — valid flags unchanged,
    controls moved by less than the ring's point spacing."""
    verts, faces = synth.uv_sphere(40, 30)
    pts = verts.astype(np.float32)
    nrm = synth.vertex_normals(verts, faces).astype(np.float32)
    ids = oracle.sample_fps(pts, 120, 0)
    graph = oracle.node_graph(pts[ids], 8)
    # generic: rotated cube-sphere tessellation, scaled and warped
    tv, tf = synth.cube_sphere(60, rot=synth.axis_angle([1, 2, 3], 0.7))
    tv = synth.bend_twist(tv * 1.03, 0.02)
    tn = synth.vertex_normals(tv, tf).astype(np.float32)
    c0, v0, c1, v1 = _literal_vs_bucketed(pts[ids], nrm[ids], tv.astype(np.float32), tn, graph)
    assert np.array_equal(v0, v1) and v0.sum() > 60
    assert np.array_equal(c0, c1)
    # symmetric: concentric UV sphere -> exact projDist ties
    tv, tf = synth.uv_sphere(110, 90, radius=1.03)
    tn = synth.vertex_normals(tv, tf).astype(np.float32)
    c0, v0, c1, v1 = _literal_vs_bucketed(pts[ids], nrm[ids], tv.astype(np.float32), tn, graph)
    assert np.array_equal(v0, v1)
    moved = np.abs(c0 - c1).max(1)
    assert (moved > 0).sum() <= 6 and moved.max() < 0.06  # target point spacing ~ 2 pi 1.03 / 110
    # adversarial: nodes above a regular lattice plane
    gx, gy = np.meshgrid(np.arange(-20, 21), np.arange(-20, 21))
    plane = np.stack([gx.ravel() * 0.01, gy.ravel() * 0.01, np.zeros(gx.size)], 1).astype(np.float32)
    pn = np.tile(np.array([0, 0, 1], np.float32), (len(plane), 1))
    nodes = np.array([[0.0, 0.0, 0.05], [0.005, 0.005, 0.03], [0.0031, -0.0012, 0.04]], np.float32)
    nn = np.tile(np.array([0, 0, 1], np.float32), (3, 1))
    g3 = np.array([[0, 1, 2]] * 3, np.int32)
    a0, w0, a1, w1 = _literal_vs_bucketed(nodes, nn, plane, pn, g3)
    assert np.array_equal(w0, w1)
    assert np.abs(a0 - a1).max() <= 0.01 + 1e-9  # at most one lattice spacing
