"""GPU parity: as-rigid-as-possible mode of Deform() (csrc/arap.cu) against oracle/arap.cpp.

Bar: deformed vertices within 1e-4 of the bounding-box diagonal (north_star) after the same fixed
number of local/global iterations; the CUDA path stores vectors in float32, the oracle in float64.
"""
import numpy as np
import pytest

import oracle
from tests import synth

pytestmark = pytest.mark.gpu


def mesh(nu=48, nv=32):
    verts, faces = synth.uv_sphere(nu, nv)
    return verts.astype(np.float32), faces.astype(np.int32)


def diag(p):
    return float(np.linalg.norm(p.max(0) - p.min(0)))


def vertex_normals(pts, faces):
    return pts / np.linalg.norm(pts, axis=1, keepdims=True)


def make(api, pts, faces, m):
    d = api.Deformation(pts, vertex_normals(pts, faces), faces)
    d.SampleFPS(m, 0)
    return d


def test_arap_matches_oracle_fixed_iterations():
    from multiviewstitch_b200 import api
    pts, faces = mesh()
    d = make(api, pts, faces, 40)
    ids = d.Nodes()
    assert np.array_equal(ids, oracle.sample_fps(pts, 40, 0))
    rng = np.random.default_rng(11)
    R = synth.axis_angle([0, 0, 1], 0.3)
    ctrl = pts[ids].astype(np.float64) @ R.T * 1.05 + 0.03 * rng.normal(size=(len(ids), 3))
    prm = api.arap_params(iters=4, tol=0.0, cg_iters=600, cg_tol=0.0)
    e, it = d.ARAP(ctrl, prm)
    got = d.Vertices()
    want, e_ref, it_ref = oracle.arap(pts, faces, ids, ctrl, iters=4, tol=0.0, cg_iters=600, cg_tol=0.0)
    assert it == it_ref == 4
    err = np.abs(got - want).max() / diag(pts)
    assert err < 1e-4, err
    assert abs(e - e_ref) <= 1e-3 * max(e_ref, 1e-12)
    # the control vertices sit exactly on their (float32-rounded) targets
    assert np.array_equal(got[ids].astype(np.float32), ctrl.astype(np.float32))
    # normals are recomputed from the deformed mesh
    nrm = d.Normals()
    assert np.allclose(np.linalg.norm(nrm, axis=1), 1.0, atol=1e-5)


def test_arap_recovers_rigid_motion():
    from multiviewstitch_b200 import api
    pts, faces = mesh(32, 20)
    d = make(api, pts, faces, 16)
    ids = d.Nodes()
    R = synth.axis_angle([1, 2, 3], 0.6)
    t = np.array([0.2, 0.1, -0.3])
    ctrl = pts[ids].astype(np.float64) @ R.T + t
    e, it = d.ARAP(ctrl, api.arap_params(iters=80, tol=0.0, cg_iters=400, cg_tol=1e-7))
    want = pts.astype(np.float64) @ R.T + t
    assert np.abs(d.Vertices() - want).max() / diag(pts) < 1e-4
    assert e < 1e-6


def test_deform_arap_entry_point_matches_oracle_pipeline():
    """Deform(tpts, tnormals, projLenErr, projDistErr) in ARAP mode: node correspondences and
    smoothing (Deformation.cpp:266-381) feed the ARAP solve, then the result becomes the rest mesh."""
    from multiviewstitch_b200 import api
    pts, faces = mesh(40, 28)
    nrm = vertex_normals(pts, faces)
    tgt = (pts.astype(np.float64) * np.array([1.06, 0.97, 1.02])).astype(np.float32)
    d = make(api, pts, faces, 32)
    ids = d.Nodes()
    d.BuildGraph(4, 8)
    d.SetTarget(tgt, nrm)
    ctrl, valid = d.NodeTargets(100.0, 100.0, 2)
    prm = api.arap_params(iters=3, tol=0.0, cg_iters=500, cg_tol=0.0)
    d.DeformARAP(None, None, 100.0, 100.0, params=prm)
    got = d.Vertices()
    want, _, _ = oracle.arap(pts, faces, ids, ctrl, iters=3, tol=0.0, cg_iters=500, cg_tol=0.0)
    assert np.abs(got - want).max() / diag(pts) < 1e-4


def test_arap_errors():
    from multiviewstitch_b200 import api
    pts, faces = mesh(16, 10)
    d = api.Deformation(pts, vertex_normals(pts, faces), faces)
    with pytest.raises(api.MvsError):
        d.ARAP(np.zeros((0, 3)))  # no nodes sampled
    d.SampleFPS(8, 0)
    with pytest.raises(api.MvsError):
        d.ARAP(pts[d.Nodes()], api.arap_params(iters=0))
