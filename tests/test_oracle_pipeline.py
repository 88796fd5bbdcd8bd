"""CPU tests pinning oracle/pipeline.cpp (RemoveOutliers, CheckConsistencyCore, Depth2Model) with
analytic known answers; the reference has no tests for them."""
import numpy as np

import oracle
from tests import synth


def _cam(w=64, h=48, f=60.0, R=None, t=None):
    K = np.array([[f, 0, w / 2 - 0.5], [0, f, h / 2 - 0.5], [0, 0, 1.0]])
    return oracle.make_camera(K, np.eye(3) if R is None else R, np.zeros(3) if t is None else t, w, h)


def test_depth2model_plane_and_depth_jump():
    cam = _cam()
    d = np.full((48, 64), 0.2, np.float32)
    d[:, 40:] = 0.1          # a depth discontinuity between columns 39 and 40
    d[10:14, 5:9] = 0.0      # a hole
    d[0, 0] = 0.45           # beyond MaxDsp
    xyz, tex, fac = oracle.depth2model(d, cam, 0.0025, 0.3, 0.12)
    valid = (d > 0) & (d >= 0.0025) & (d <= 0.3)
    assert len(xyz) == valid.sum() and np.array_equal(tex, np.flatnonzero(valid.ravel()))
    # vertex k is the unprojection of its pixel at depth 1/d (Depth2Model.cpp:34)
    k = 1234
    v, u = divmod(int(tex[k]), 64)
    assert np.allclose(xyz[k], oracle.unproject(cam, u, v, 1.0 / float(d[v, u])))
    # threshold = 0.12 * (0.3 - 0.0025) / 100 = 3.57e-4: no triangle spans the 0.2 -> 0.1 jump
    px = tex[fac]                       # pixel ids of every facet corner
    cols = px % 64
    assert not np.any((cols.min(1) <= 39) & (cols.max(1) >= 40))
    # a full quad away from holes gives the two triangles (a, below, diag) and (a, diag, right)
    vid = -np.ones(48 * 64, np.int64); vid[tex] = np.arange(len(tex))
    a, b, c, dd = vid[20 * 64 + 20], vid[21 * 64 + 20], vid[21 * 64 + 21], vid[20 * 64 + 21]
    rows = {tuple(r) for r in fac.tolist()}
    assert (a, b, c) in rows and (a, c, dd) in rows
    # every facet uses valid vertices; quads touching the hole lose their triangles
    assert fac.min() >= 0 and fac.max() < len(xyz)
    hole = vid[10 * 64 + 5]
    assert hole == -1


def test_check_consistency_keeps_consistent_and_drops_inconsistent():
    # fronto-parallel plane at z = 4 seen by two cameras 1.0 apart (15 px of parallax): consistent
    # everywhere they overlap
    w, h = 64, 48
    c0 = _cam()
    c1 = _cam(t=np.array([-1.0, 0, 0]))
    d0 = np.full((h, w), 0.25, np.float32)
    d1 = np.full((h, w), 0.25, np.float32)
    out = oracle.check_consistency(d0, c0, d1[None], [c1], 0.0025, 0.3, 4.0)
    keep = out > 0
    assert keep[:, 16:].all() and not keep[:, :14].any()  # left margin projects outside camera 1
    assert np.all(out[keep] == np.float32(0.25))
    # a wrong neighbour depth breaks the round trip (Processor.cpp:93-97)
    d1b = d1.copy(); d1b[:, 30:40] = 0.125
    out2 = oracle.check_consistency(d0, c0, d1b[None], [c1], 0.0025, 0.3, 4.0)
    assert (out2[:, 46:54] == 0).all() and (out2[:, 20:40] > 0).all()  # 7.5 px > ReprojError 4
    # out-of-range own depth is zeroed (Processor.cpp:110-112)
    d0c = d0.copy(); d0c[3, 3] = 0.4
    assert oracle.check_consistency(d0c, c0, d1[None], [c1], 0.0025, 0.3, 4.0)[3, 3] == 0
    assert np.array_equal(oracle.check_consistency(d0, c0, np.zeros((0, h, w)), [], 0.0025, 0.3, 4.0), d0)


def test_remove_outliers_drops_planted_outliers():
    p, q = synth.planted_pairs(400, seed=3, noise=5e-4)
    rng = np.random.default_rng(0)
    bad = rng.choice(400, 40, replace=False)
    q[bad] += rng.normal(scale=0.4, size=(40, 3))
    (K, R1, t1, w, h), (_, R2, t2, _, _) = synth.camera_pair()
    c1, c2 = oracle.make_camera(K, R1, t1, w, h), oracle.make_camera(K, R2, t2, w, h)
    keep, err, ir = oracle.remove_outliers(p, q, c1, c2, 3, 200, 20.0, 0.6, seed=5)
    assert ir == 0.0  # never updated in the reference (Processor.cpp:243-253 commented out)
    assert len(set(keep.tolist()) & set(bad.tolist())) <= 0.1 * len(bad)
    assert len(keep) >= 340 and np.all(np.diff(keep) > 0)
    # the inlier test itself: both symmetric pixel errors under the threshold
    s, R, t = oracle.srt_estimate(p[keep], q[keep])
    flags, eall = oracle.srt_inliers(p, q, c1, c2, s, R, t, 3.0)
    assert flags[bad].sum() <= 2 and flags[keep].mean() > 0.95
    assert abs(eall / 400 - oracle.srt_residual(p, q, c1, c2, s, R, t)) < 1e-9
