"""Host-side driver of the hot path for one view pair, in the reference's call order.

`Processor` itself (file shuffling, SIFT, Poisson, GUI) is out of scope; this module only strings
the hot-path calls together the way `Processor::AlignmentSeq` / `Deform` / `Render` do for one pair of
views (BASELINE.json configs[0] "two-view 640x480 depth-map pair" and, sharded over ranks,
configs[3] "8-view turntable, 28 pairs"):

    depth maps --SolveUnProjectionD--> points                     [Image3D/Image3D.cpp:92-106]
    pixel matches --RemoveOutliers (3 x RANSAC + inlier test)--> inlier matches
                                                                  [Processor/Processor.cpp:177-269]
    --SRTSolver::EstimateTransform--> (s, R, t)                   [Processor/Processor.cpp:814-818]
    view-1 mesh (Depth2Model) --apply SRT--> frame of view 2      [Processor/Processor.cpp:1025-1026]
    --Deformation::Deform onto the view-2 points--> deformed mesh [Processor/Processor.cpp:1135-1137]
    --Model2Depth--> re-projected inverse-depth map of view 2     [Processor/Processor.cpp:1190-1191]

Everything numeric happens in libmvs_b200.so through `api`; this file is marshalling only.
"""
from __future__ import annotations

import numpy as np

from . import api, synth

MIN_DSP, MAX_DSP = 0.0025, 0.3  # config.txt:23-24
SMOOTH_THRES = 0.12             # config.txt:38
SAMPLE_INTERVAL = 20            # config.txt:9


def make_scene(side=96, view_angle_deg=10.0, w=640, h=480, seed=0):
    """C1 scene: bumpy unit sphere 5 units in front of camera 1; view 2 is rotated
    about y and lives in its own world frame x_B = s R x_A + t (the separate-SfM-frame situation the
    SRT solver resolves)."""
    verts, faces = synth.cube_sphere(side, bumps=0.05)
    verts = verts + np.array([0.0, 0.0, 5.0])
    K = synth.c1_intrinsics()
    s, R, t = synth.PLANTED_S, synth.PLANTED_R, synth.PLANTED_T
    verts_b = s * verts @ R.T + t
    R1, t1 = np.eye(3), np.zeros(3)
    c2 = s * R @ np.array([0, 0, 5.0]) + t
    Rv = synth.axis_angle([0, 1, 0], np.deg2rad(view_angle_deg + 3.0 * seed))
    R2 = Rv @ R.T                       # camera 2 sees frame B roughly as camera 1 sees frame A
    t2 = np.array([0, 0, 5.0 * s]) - R2 @ c2
    return dict(verts_a=verts, verts_b=verts_b, faces=faces, K=K, cam_a=(K, R1, t1, w, h),
                cam_b=(K, R2, t2, w, h), planted=(s, R, t))


def lattice_matches(points_a, valid_a, cam_b_tuple, valid_b, planted, w, h, step=SAMPLE_INTERVAL,
                    outlier_frac=0.05, seed=0):
    """Ground-truth pixel matches on a `step`-pixel lattice of view 1 (SampleIterval, config.txt:9)
    plus a fraction of planted outliers; stands in for the SIFT matches (out of scope)."""
    s, R, t = planted
    K, R2, t2, _, _ = cam_b_tuple
    rng = np.random.default_rng(seed)
    us, vs = np.meshgrid(np.arange(step // 2, w, step), np.arange(step // 2, h, step))
    us, vs = us.ravel(), vs.ravel()
    m = []
    for u, v in zip(us, vs):
        k = v * w + u
        if not valid_a[k]:
            continue
        pb = s * R @ points_a[k] + t
        pc = R2 @ pb + t2
        u2 = int(K[0, 0] * pc[0] / pc[2] + K[0, 2] + 0.5)
        v2 = int(K[1, 1] * pc[1] / pc[2] + K[1, 2] + 0.5)
        if 0 <= u2 < w and 0 <= v2 < h and valid_b[v2 * w + u2]:
            m.append((u, v, u2, v2))
    m = np.array(m, np.int32).reshape(-1, 4)
    nout = int(outlier_frac * len(m))
    vb = np.flatnonzero(valid_b)
    for i in rng.choice(len(m), nout, replace=False):
        k = int(rng.choice(vb))
        m[i, 2], m[i, 3] = k % w, k // w
    return m


def stitch_pair(depth_a, cam_a, depth_b, cam_b, matches, ctx=None, n_nodes=512, gn_iters=4,
                pcg_iters=20, seed=0, iter_num=200, pixel_err=60.0, adapt_ratio=0.75,
                host_vectors=False):
    """One view pair through the hot path.  cam_a / cam_b: api.Camera.  matches: (n, 4) int pixel
    pairs (u1, v1, u2, v2).  Returns a dict with every intermediate the tests compare.

    host_vectors=True follows the reference call for call — both views are unprojected into full
    h*w host vectors (Image3D::SolveUnProjectionD), the matches and the valid target points are
    picked out of them on the host and the target goes back to the device through SetTarget.  The
    default keeps those intermediates on the device: only the matched pixels are unprojected
    (unproject_pixels) and the target is unprojected + compacted where it is used
    (Deformation.SetTargetDepth).  Both give bit-identical results (tests/test_gpu_pipeline.py);
    the host-vector form moves 17 MB per pair through pageable memory."""
    w, h = cam_a.w, cam_a.h
    if host_vectors:
        pts_a, val_a = api.unproject_depth(depth_a, cam_a, MIN_DSP, MAX_DSP, ctx)
        pts_b, val_b = api.unproject_depth(depth_b, cam_b, MIN_DSP, MAX_DSP, ctx)
        ka = matches[:, 1] * w + matches[:, 0]
        kb = matches[:, 3] * w + matches[:, 2]
        p, q = pts_a[ka], pts_b[kb]
    else:
        p, _ = api.unproject_pixels(depth_a, cam_a, matches[:, 0:2], MIN_DSP, MAX_DSP, ctx)
        q, _ = api.unproject_pixels(depth_b, cam_b, matches[:, 2:4], MIN_DSP, MAX_DSP, ctx)
    keep, err, _ = api.remove_outliers(p, q, cam_a, cam_b, 3, iter_num, pixel_err, adapt_ratio, seed, ctx)
    solver = api.SRTSolver(iter_num, ctx)
    solver.SetPrintFlag(False)
    solver.SetInput(p[keep], q[keep], cam_a, cam_b)
    scale, R, t = solver.EstimateTransform()
    resid = solver.ResidualError(scale, R, t)
    solver.close()
    # template = view-1 mesh in the frame of view 2
    va, tex_a, fa = api.depth2model(depth_a, cam_a, MIN_DSP, MAX_DSP, SMOOTH_THRES, ctx)
    va_b, _ = api.apply_srt(va, None, scale, R, t, ctx=ctx)
    d = api.Deformation(va_b, None, fa, ctx)
    d.SampleFPS(min(n_nodes, len(va_b)), 0)
    if host_vectors:
        tgt = pts_b[np.flatnonzero(val_b)]  # (integer take: 3x faster than a boolean mask here)
        d.SetTarget(tgt)
    else:
        tgt = None
        d.SetTargetDepth(depth_b, cam_b, MIN_DSP, MAX_DSP)
    d.GNBegin(api.gn_params(gn_iters=gn_iters, pcg_iters=pcg_iters, corr_mode=1))
    for _ in range(gn_iters):
        d.GNStep()
    d.GNEnd()
    deformed = d.Vertices()
    d.close()
    m2d = api.Model2Depth(ctx)
    m2d.SetInput([deformed], [None], [fa], [[cam_b]])
    render = m2d.render(0)[0]
    return dict(scale=scale, R=R, t=t, kept=keep, err=err, residual=resid, template=va_b, facets=fa,
                deformed=deformed, render=render, target=tgt)
