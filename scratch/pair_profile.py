"""cProfile of pipeline.stitch_pair (C4 unit of work) to separate host glue from device time."""
import sys, time, cProfile, pstats, numpy as np
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
for _ in range(3):
    pipeline.stitch_pair(da, ca, db, cb, mt, ctx, n_nodes=512, gn_iters=4)
ctx.sync()
t0 = time.perf_counter()
for _ in range(10):
    pipeline.stitch_pair(da, ca, db, cb, mt, ctx, n_nodes=512, gn_iters=4)
ctx.sync()
print("stitch_pair %.2f ms" % (1e2 * (time.perf_counter() - t0)))
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    pipeline.stitch_pair(da, ca, db, cb, mt, ctx, n_nodes=512, gn_iters=4)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
