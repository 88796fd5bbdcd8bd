"""Stage times of pipeline.stitch_pair measured in context (outputs retained as bench.py does)."""
import sys, time, numpy as np
sys.path.insert(0, '.')
from multiviewstitch_b200 import api, pipeline
import multiviewstitch_b200.pipeline as P
ctx = api.Context(0)
sc = pipeline.make_scene(side=72)
ca, cb = api.make_camera(*sc["cam_a"]), api.make_camera(*sc["cam_b"])
m2d = api.Model2Depth(ctx)
m2d.SetInput([sc["verts_a"], sc["verts_b"]], [None, None], [sc["faces"], sc["faces"]], [[ca], [cb]])
da, db = m2d.render(0)[0], m2d.render(1)[0]
pa, va = api.unproject_depth(da, ca, pipeline.MIN_DSP, pipeline.MAX_DSP, ctx)
pb, vb = api.unproject_depth(db, cb, pipeline.MIN_DSP, pipeline.MAX_DSP, ctx)
mt = pipeline.lattice_matches(pa, va, sc["cam_b"], vb, sc["planted"], 640, 480)
acc = {}
def tick(name, t0):
    ctx.sync(); t = time.perf_counter(); acc[name] = acc.get(name, 0.0) + (t - t0); return t
def pair():
    w = ca.w
    t = time.perf_counter()
    pts_a, val_a = api.unproject_depth(da, ca, P.MIN_DSP, P.MAX_DSP, ctx); t = tick("unproject a", t)
    pts_b, val_b = api.unproject_depth(db, cb, P.MIN_DSP, P.MAX_DSP, ctx); t = tick("unproject b", t)
    ka = mt[:, 1] * w + mt[:, 0]; kb = mt[:, 3] * w + mt[:, 2]
    p, q = pts_a[ka], pts_b[kb]; t = tick("gather matches", t)
    keep, err, _ = api.remove_outliers(p, q, ca, cb, 3, 200, 60.0, 0.75, 0, ctx); t = tick("remove_outliers", t)
    s = api.SRTSolver(200, ctx); s.SetPrintFlag(False); s.SetInput(p[keep], q[keep], ca, cb)
    scale, R, tt = s.EstimateTransform(); s.ResidualError(scale, R, tt); s.close(); t = tick("srt", t)
    va_, tex, fa = api.depth2model(da, ca, P.MIN_DSP, P.MAX_DSP, P.SMOOTH_THRES, ctx); t = tick("depth2model", t)
    vab, _ = api.apply_srt(va_, None, scale, R, tt, ctx=ctx); t = tick("apply_srt", t)
    tgt = pts_b[val_b.astype(bool)]; t = tick("compact target (numpy)", t)
    d = api.Deformation(vab, None, fa, ctx); t = tick("Deformation()", t)
    d.SampleFPS(512, 0); t = tick("SampleFPS", t)
    d.SetTarget(tgt); t = tick("SetTarget", t)
    d.GNBegin(api.gn_params(gn_iters=4, pcg_iters=20, corr_mode=1)); t = tick("GNBegin (graph)", t)
    for _ in range(4): d.GNStep()
    t = tick("4 GNStep", t)
    d.GNEnd(); t = tick("GNEnd", t)
    out = d.Vertices(); d.close(); t = tick("Vertices+close", t)
    m = api.Model2Depth(ctx); m.SetInput([out], [None], [fa], [[cb]]); r = m.render(0)[0]; t = tick("Model2Depth", t)
    return dict(a=pts_a, b=pts_b, tgt=tgt, out=out, r=r, vab=vab)
keepall = []
for _ in range(3): keepall.append(pair())
acc.clear()
N = 10
for _ in range(N): keepall.append(pair())
tot = 0
for k, v in acc.items():
    print(f"{k:26s} {1e3 * v / N:7.3f} ms"); tot += v
print(f"{'total':26s} {1e3 * tot / N:7.3f} ms")
