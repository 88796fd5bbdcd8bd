"""How much of the unseeded 1-NN cost is SIMT divergence?  Same queries, (a) as they are (32
different walks per warp), (b) every query replicated over the 32 lanes of its warp."""
import sys, numpy as np
sys.path.insert(0, '.')
from multiviewstitch_b200 import api, synth
verts, faces = synth.cube_sphere(408)
pts = verts.astype(np.float32)
tv, _ = synth.cube_sphere(419, rot=synth.axis_angle([1, 2, 3], 0.7))
tgt = synth.bend_twist(tv, 0.05).astype(np.float32)
ctx = api.Context(0)
idx = api.NNIndex(ctx).build(tgt)
def run(name, q):
    idx.knn(q[:1000], 1)
    ctx.prof_enable(True); ctx.prof_reset()
    i, d2 = idx.knn(q, 1)
    p = ctx.prof(); ctx.prof_enable(False)
    print(f"{name:34s} n={len(q):8d}  nn1 {1e3*p['nn1_pyramid'][0]:9.1f} us  mean d {np.sqrt(d2).mean():.4f}")
run("all, distinct per lane", pts)
sub = pts[::32]
run("1/32 of them, distinct", sub)
run("1/32 of them, each x32 (uniform)", np.repeat(sub, 32, axis=0))
idx.close(); ctx.close()
