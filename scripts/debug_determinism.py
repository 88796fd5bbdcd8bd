"""Run-to-run spread of the C3 solve (float32 atomics make block sums order-dependent)."""
import sys, numpy as np
sys.path.insert(0, '.')
from multiviewstitch_b200 import api, synth
nu = int(sys.argv[1]) if len(sys.argv) > 1 else 408
m = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
verts, faces = synth.cube_sphere(nu)
pts = verts.astype(np.float32)
tv, _ = synth.cube_sphere(int(round(nu * 1.027)), rot=synth.axis_angle([1, 2, 3], 0.7))
tgt = synth.bend_twist(tv, 0.05).astype(np.float32)
diag = float(np.linalg.norm(pts.max(0) - pts.min(0)))
ctx = api.Context(0)
d = api.Deformation(pts, None, None, ctx)
d.SampleFPS(m, 0)
d.BuildGraph(4, 8)
d.SetTarget(tgt)
for mode, iters in ((1, 10), (2, 10)):
    outs, states = [], []
    if mode == 2:
        d.SetTargets(synth.bend_twist(verts, 0.05))
    for rep in range(3):
        d.GNBegin(api.gn_params(gn_iters=iters, pcg_iters=20, corr_mode=mode))
        st = []
        for it in range(iters):
            d.GNStep()
            st.append(d.NodeState().copy())
        d.GNEnd()
        outs.append(d.Vertices())
        states.append(st)
    print("corr_mode", mode, "vertex spread / diag:", [float(np.abs(outs[0] - o).max() / diag) for o in outs[1:]])
    print("  node-state spread per GN iter:", ["%.1e" % float(np.abs(a - b).max()) for a, b in zip(states[0], states[1])])
