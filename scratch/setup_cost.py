"""C3 setup cost on the device: construction, sampling, graph build, target index."""
import sys, time, numpy as np
sys.path.insert(0, '.')
from multiviewstitch_b200 import api, synth
nu, m = 408, 4096
verts, faces = synth.cube_sphere(nu)
pts = verts.astype(np.float32)
tv, _ = synth.cube_sphere(int(round(nu * 1.027)), rot=synth.axis_angle([1, 2, 3], 0.7))
tgt = synth.bend_twist(tv, 0.05).astype(np.float32)
ctx = api.Context(0)
for rep in range(2):
    t0 = time.perf_counter(); d = api.Deformation(pts, None, faces, ctx); ctx.sync(); t1 = time.perf_counter()
    d.SampleFPS(m, 0); ctx.sync(); t2 = time.perf_counter()
    d.BuildGraph(4, 8); ctx.sync(); t3 = time.perf_counter()
    d.SetTarget(tgt); ctx.sync(); t4 = time.perf_counter()
    ctx.prof_enable(True); ctx.prof_reset()
    d.GNBegin(api.gn_params(gn_iters=1, pcg_iters=20, corr_mode=1)); d.GNStep(); d.GNEnd(); ctx.sync(); t5 = time.perf_counter()
    ctx.prof_enable(False)
    print("create %.1f  fps %.1f  build_graph %.1f  set_target %.1f  first GN iter %.1f ms" % tuple(1e3 * x for x in (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)))
    d.close()
ctx.prof_enable(True); ctx.prof_reset()
d = api.Deformation(pts, None, faces, ctx); d.SampleFPS(m, 0); d.BuildGraph(4, 8); ctx.sync()
print({k: round(v[0], 2) for k, v in sorted(ctx.prof().items(), key=lambda kv: -kv[1][0])[:12]}, "ms")
