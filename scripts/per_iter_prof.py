"""Per-GN-iteration kernel-family times of the C3 solve (CUDA events inside the library)."""
import sys, numpy as np
sys.path.insert(0, '.')
from multiviewstitch_b200 import api, synth
nu = int(sys.argv[1]) if len(sys.argv) > 1 else 408
m = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
verts, faces = synth.cube_sphere(nu)
pts = verts.astype(np.float32)
tv, _ = synth.cube_sphere(int(round(nu * 1.027)), rot=synth.axis_angle([1, 2, 3], 0.7))
tgt = synth.bend_twist(tv, 0.05).astype(np.float32)
ctx = api.Context(0)
d = api.Deformation(pts, None, None, ctx)
d.SampleFPS(m, 0)
d.BuildGraph(4, 8)
d.SetTarget(tgt)
print(d.Stats())
prm = api.gn_params(gn_iters=10, pcg_iters=20, corr_mode=1)
for rep in range(3):
    d.GNBegin(prm)
    rows = []
    for it in range(10):
        ctx.prof_enable(True); ctx.prof_reset()
        d.GNStep()
        ctx.sync()
        rows.append(ctx.prof())
    d.GNEnd()
ctx.prof_enable(False)
names = sorted(rows[0], key=lambda k: -rows[0][k][0])
print("iter  " + "  ".join(f"{n:>16s}" for n in names))
for it, r in enumerate(rows):
    print(f"{it:4d}  " + "  ".join(f"{1e3 * r.get(n, (0, 0))[0]:16.1f}" for n in names), " us")
