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
for nn in (128, 256, 512):
    for gi in (1, 4):
        b = api.PairBatch(ctx)
        b.add(da, ca, db, cb, mt, seed=0)
        try:
            o = b.run(api.pair_params(n_nodes=nn, gn_iters=gi, workers=1))[0]
            print(nn, gi, "ok", np.isfinite(o["deformed"]).all(), np.abs(o["deformed"] - o["template"]).max())
        except Exception as e:
            print(nn, gi, "FAILED", str(e))
