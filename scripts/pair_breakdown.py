"""Wall-clock breakdown of one C1 view pair through pipeline.stitch_pair's stages."""
import sys, time, numpy as np
sys.path.insert(0, '.')
from multiviewstitch_b200 import api, pipeline
ctx = api.Context(0)
sc = pipeline.make_scene(side=72)
ca, cb = api.make_camera(*sc["cam_a"]), api.make_camera(*sc["cam_b"])
m2d = api.Model2Depth(ctx)
m2d.SetInput([sc["verts_a"], sc["verts_b"]], [None, None], [sc["faces"], sc["faces"]], [[ca], [cb]])
da, db = m2d.render(0)[0], m2d.render(1)[0]
pa, va = api.unproject_depth(da, ca, pipeline.MIN_DSP, pipeline.MAX_DSP, ctx)
pb, vb = api.unproject_depth(db, cb, pipeline.MIN_DSP, pipeline.MAX_DSP, ctx)
mt = pipeline.lattice_matches(pa, va, sc["cam_b"], vb, sc["planted"], 640, 480)
w = 640
def T(name, f, reps=5):
    f(); ctx.sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = f()
    ctx.sync()
    print(f"{name:28s} {1e3*(time.perf_counter()-t0)/reps:8.2f} ms")
    return r
T("unproject x2", lambda: (api.unproject_depth(da, ca, pipeline.MIN_DSP, pipeline.MAX_DSP, ctx), api.unproject_depth(db, cb, pipeline.MIN_DSP, pipeline.MAX_DSP, ctx)))
p, q = pa[mt[:, 1] * w + mt[:, 0]], pb[mt[:, 3] * w + mt[:, 2]]
keep, err, _ = T("remove_outliers", lambda: api.remove_outliers(p, q, ca, cb, 3, 200, 60.0, 0.75, 0, ctx))
def srt():
    s = api.SRTSolver(200, ctx); s.SetPrintFlag(False); s.SetInput(p[keep], q[keep], ca, cb)
    r = s.EstimateTransform(); s.ResidualError(*r); s.close(); return r
scale, R, t = T("SRT closed form + residual", srt)
va_, tex, fa = T("depth2model", lambda: api.depth2model(da, ca, pipeline.MIN_DSP, pipeline.MAX_DSP, pipeline.SMOOTH_THRES, ctx))
vab, _ = T("apply_srt", lambda: api.apply_srt(va_, None, scale, R, t, ctx=ctx))
tgt = pb[vb.astype(bool)]
print("template verts", len(vab), "facets", len(fa), "target pts", len(tgt), "matches", len(mt))
def deform():
    d = api.Deformation(vab, None, fa, ctx)
    t0 = time.perf_counter(); d.SampleFPS(512, 0); ctx.sync(); t1 = time.perf_counter()
    d.SetTarget(tgt); ctx.sync(); t2 = time.perf_counter()
    d.BuildGraph(4, 8); ctx.sync(); t3 = time.perf_counter()
    d.GNBegin(api.gn_params(gn_iters=4, pcg_iters=20, corr_mode=1))
    for _ in range(4): d.GNStep()
    d.GNEnd(); ctx.sync(); t4 = time.perf_counter()
    out = d.Vertices(); d.close(); t5 = time.perf_counter()
    return out, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)
out, tt = T("Deformation total", deform)
print("   fps %.2f  set_target %.2f  build_graph %.2f  4 GN iters %.2f  get+close %.2f ms" % tuple(1e3 * x for x in tt))
def render():
    m = api.Model2Depth(ctx); m.SetInput([out], [None], [fa], [[cb]]); return m.render(0)
T("Model2Depth", render)
