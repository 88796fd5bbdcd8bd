"""Generates tests/golden/*.npz with the CPU oracle (oracle/) from seeded synthetic inputs.

The reference has no golden vectors of its own and cannot be built here, so these are
known-answer vectors OF THE ORACLE: they freeze its behaviour (a regression in oracle or CUDA path
shows up against them) and travel to the GPU box, where /root/reference does not exist.
Run:  python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import oracle  # noqa: E402
from multiviewstitch_b200 import synth  # noqa: E402

OUT = Path(__file__).resolve().parent


def cams(mod):
    (K, R1, t1, w, h), (_, R2, t2, _, _) = synth.camera_pair()
    return mod.make_camera(K, R1, t1, w, h), mod.make_camera(K, R2, t2, w, h)


def arap_golden():
    """ARAP mode of Deform() (oracle/arap.cpp): nodes of a sphere mesh as control vertices moved by
    a rotation, an anisotropic scale and noise; 4 local/global iterations x 300 CG iterations."""
    rng = np.random.default_rng(77)
    verts, faces = synth.uv_sphere(36, 26)
    pts = verts.astype(np.float32)
    ids = oracle.sample_fps(pts, 30, 0)
    R = synth.axis_angle([0.2, 1.0, 0.3], 0.35)
    ctrl = (pts[ids].astype(np.float64) * np.array([1.08, 0.95, 1.02])) @ R.T + 0.02 * rng.normal(size=(30, 3))
    out, energy, it = oracle.arap(pts, faces, ids, ctrl, iters=4, tol=0.0, cg_iters=300, cg_tol=0.0)
    rp, col, w = oracle.cot_weights(pts, faces)
    np.savez(OUT / "arap.npz", pts=pts, faces=faces.astype(np.int32), ids=ids, ctrl=ctrl, out=out,
             energy=energy, row_ptr=rp, col=col, w=w)


def main():
    arap_golden()
    rng = np.random.default_rng(2024)
    # ---- exact kNN
    t = rng.normal(size=(400, 3)).astype(np.float32)
    q = (rng.normal(size=(64, 3)) * 1.2).astype(np.float32)
    idx, d2 = oracle.knn(q, t, 5, "brute")
    np.savez(OUT / "knn.npz", t=t, q=q, idx=idx, d2=d2)
    # ---- SRT closed form, residual, RANSAC
    p, qq = synth.planted_pairs(300, seed=11, noise=2e-3)
    out = rng.random(300) < 0.1
    qq[out] += rng.normal(scale=0.3, size=(out.sum(), 3))
    c1, c2 = cams(oracle)
    s, R, tt = oracle.srt_estimate(p, qq)
    res = oracle.srt_residual(p, qq, c1, c2, s, R, tt)
    triples = oracle.draw_triples(300, 50, 7)
    rs, rR, rt, rerr, rbest, errs = oracle.srt_ransac(p, qq, c1, c2, triples)
    np.savez(OUT / "srt.npz", p=p, q=qq, s=s, R=R, t=tt, residual=res, triples=triples, rs=rs, rR=rR,
             rt=rt, rerr=rerr, rbest=rbest, errs=errs)
    # ---- deformation graph + Gauss-Newton
    verts, faces = synth.uv_sphere(40, 30)
    pts = verts.astype(np.float32)
    ids = oracle.sample_fps(pts, 60, 0)
    us = oracle.uniform_sampling(pts, 16)
    widx, w = oracle.node_weights(pts, pts[ids], 4)
    graph = oracle.node_graph(pts[ids], 8)
    targets = synth.bend_twist(verts, 0.05)
    x, outp, en = oracle.gn_solve(pts, None, ids, widx, w, graph,
                                  oracle.gn_params(gn_iters=6, pcg_iters=20), targets=targets)
    tv = synth.bend_twist(synth.uv_sphere(90, 70)[0], 0.04).astype(np.float32)
    x1, out1, en1 = oracle.gn_solve(pts, None, ids, widx, w, graph,
                                    oracle.gn_params(gn_iters=4, pcg_iters=20, corr_mode=1), tpts=tv)
    np.savez(OUT / "deform.npz", pts=pts, faces=faces, fps=ids, uniform=us, widx=widx, w=w, graph=graph,
             targets=targets, x=x, out=outp, energy=en, tv=tv, x_nn=x1, out_nn=out1)
    # ---- Model2Depth
    K = np.array([[90.0, 0, 47.5], [0, 90.0, 35.5], [0, 0, 1.0]])
    cam_list, cam_par = [], []
    for k in range(2):
        Rk = synth.axis_angle([0.1, 1.0, 0.05], 0.4 * k)
        tk = np.array([0.05 * k, -0.02 * k, 3.0 + 0.3 * k])
        cam_list.append(oracle.make_camera(K, Rk, tk, 96, 72))
        cam_par.append(np.concatenate([Rk.ravel(), tk]))
    mv, mf = synth.uv_sphere(60, 45)
    img = oracle.model2depth(mv.astype(np.float32), mf, cam_list)
    np.savez(OUT / "model2depth.npz", v=mv.astype(np.float32), f=mf, K=K, cams=np.array(cam_par), img=img)
    # ---- alignment
    src, sl = synth.humanoid(120, seed=1, arm_angle=0.5, leg_angle=0.12)
    tgt0, _ = synth.humanoid(300, seed=2, arm_angle=0.75, leg_angle=0.2)
    Rg = synth.axis_angle([0.1, 1.0, 0.2], 0.2)
    tgt = 1.6 * tgt0 @ Rg.T + np.array([0.2, 0.1, 3.0])
    a = oracle.align(src, src.copy(), sl, tgt, tgt.copy(), np.zeros((0, 3), np.int32), Rg @ [0, 0, 1.0],
                     remove_ground=0, local_align=1, ground_ray=Rg @ [0, 1.0, 0])
    np.savez(OUT / "align.npz", src=src, sl=sl, tgt=tgt, view=Rg @ [0, 0, 1.0], ground=Rg @ [0, 1.0, 0],
             out_src=a[0], out_sn=a[1], out_tl=a[5])
    for f in sorted(OUT.glob("*.npz")):
        print(f.name, f.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
