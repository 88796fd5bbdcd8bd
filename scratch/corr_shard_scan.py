"""Is the seeded correspondence pass throughput bound or tail bound?  Times gn_deform_corr for
vertex sub-ranges of the C3 mesh (mvs_deform_set_shard) at GN iteration >= 2."""
import sys, numpy as np
sys.path.insert(0, '.')
from multiviewstitch_b200 import api, synth
nu, m = 408, 4096
verts, faces = synth.cube_sphere(nu)
pts = verts.astype(np.float32)
tv, _ = synth.cube_sphere(int(round(nu * 1.027)), rot=synth.axis_angle([1, 2, 3], 0.7))
tgt = synth.bend_twist(tv, 0.05).astype(np.float32)
ctx = api.Context(0)
d = api.Deformation(pts, None, None, ctx)
d.SampleFPS(m, 0)
d.BuildGraph(4, 8)
d.SetTarget(tgt)
n = len(pts)
prm = api.gn_params(gn_iters=10, pcg_iters=20, corr_mode=1)
for frac_lo, frac_hi in [(0, 1), (0, 0.5), (0.5, 1), (0, 0.25), (0.25, 0.5), (0.5, 0.75), (0.75, 1), (0, 0.05), (0.4, 0.45)]:
    d.SetShard(0, n)
    d.GNBegin(prm)
    for it in range(3):
        d.GNStep()
    d.SetShard(int(frac_lo * n), int(frac_hi * n))
    ts = []
    for it in range(3):
        ctx.prof_enable(True); ctx.prof_reset()
        d.GNStep()
        ctx.sync()
        ts.append(1e3 * ctx.prof()['gn_deform_corr'][0])
    d.GNEnd()
    print(f"range {frac_lo:.2f}-{frac_hi:.2f}: corr us {['%.1f' % t for t in ts]}")
ctx.prof_enable(False)
