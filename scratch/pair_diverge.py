import sys, numpy as np
sys.path.insert(0, '.')
from multiviewstitch_b200 import api, pipeline
ctx = api.Context(0)
sc = pipeline.make_scene(side=72, view_angle_deg=10.0, seed=0)
ca, cb = api.make_camera(*sc["cam_a"]), api.make_camera(*sc["cam_b"])
m2d = api.Model2Depth(ctx)
m2d.SetInput([sc["verts_a"], sc["verts_b"]], [None, None], [sc["faces"], sc["faces"]], [[ca], [cb]])
da, db = m2d.render(0)[0], m2d.render(1)[0]
pa, va = api.unproject_depth(da, ca, pipeline.MIN_DSP, pipeline.MAX_DSP, ctx)
pb, vb = api.unproject_depth(db, cb, pipeline.MIN_DSP, pipeline.MAX_DSP, ctx)
mt = pipeline.lattice_matches(pa, va, sc["cam_b"], vb, sc["planted"], 640, 480, seed=0)
out = pipeline.stitch_pair(da, ca, db, cb, mt, ctx, n_nodes=512, gn_iters=4, seed=0)
tmpl, fa = out["template"], out["facets"]
print("template", tmpl.shape, np.isfinite(tmpl).all(), "deformed finite", np.isfinite(out["deformed"]).all(),
      "max move", np.abs(out["deformed"] - tmpl).max())
for nn in (128, 256, 512):
    d = api.Deformation(tmpl, None, fa, ctx)
    d.SampleFPS(nn, 0)
    d.SetTargetDepth(db, cb, pipeline.MIN_DSP, pipeline.MAX_DSP)
    d.GNBegin(api.gn_params(gn_iters=4, pcg_iters=20, corr_mode=1))
    for it in range(4):
        d.GNStep()
        x = d.NodeState()
        rp, col, B, rot, rhs = d.System()
        print(nn, it, "x finite", np.isfinite(x).all(), "|x| max", np.abs(x).max(), "B max", np.abs(B).max(), "rhs max", np.abs(rhs).max())
    try:
        d.GNEnd()
        print("GNEnd ok")
    except Exception as e:
        print("GNEnd:", e)
    d.close()
print("---- Solve() in one call on the same template")
for nn in (128, 512):
    d = api.Deformation(tmpl, None, fa, ctx)
    d.SampleFPS(nn, 0)
    d.SetTargetDepth(db, cb, pipeline.MIN_DSP, pipeline.MAX_DSP)
    try:
        d.Solve(api.gn_params(gn_iters=4, pcg_iters=20, corr_mode=1))
        print(nn, "Solve ok")
    except Exception as e:
        print(nn, "Solve:", str(e)[-100:])
    d.close()
print("---- float32 template")
d = api.Deformation(tmpl.astype(np.float32), None, fa, ctx)
d.SampleFPS(128, 0)
d.SetTargetDepth(db, cb, pipeline.MIN_DSP, pipeline.MAX_DSP)
try:
    d.Solve(api.gn_params(gn_iters=1, pcg_iters=20, corr_mode=1)); print("f32 ok")
except Exception as e:
    print("f32:", str(e)[-100:])
