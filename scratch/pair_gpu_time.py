"""GPU-busy time of one pair: sum of the per-family CUDA-event times of pipeline.stitch_pair."""
import sys, numpy as np
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
ctx.prof_enable(True); ctx.prof_reset()
N = 5
for _ in range(N):
    pipeline.stitch_pair(da, ca, db, cb, mt, ctx, n_nodes=512, gn_iters=4)
prof = ctx.prof(); ctx.prof_enable(False)
tot = 0
for k, (ms, n) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:28s} {1e3 * ms / N:9.1f} us  {n / N:6.1f} launches"); tot += ms
print("sum", 1e3 * tot / N, "us per pair")
