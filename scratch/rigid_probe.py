import numpy as np, sys
sys.path.insert(0, '.')
from multiviewstitch_b200 import api
from tests import synth
ctx = api.Context(0)
verts, faces = synth.uv_sphere(64, 48); pts = verts.astype(np.float32)
R = synth.axis_angle([0.2, 1.0, 0.1], 0.3); tt = np.array([0.1, -0.05, 0.2])
targets = pts.astype(np.float64) @ R.T + tt
for pcg in (20, 50, 200):
    d = api.Deformation(pts, None, faces, ctx)
    d.SampleFPS(60, 0)
    d.SetTargets(targets)
    d.GNBegin(api.gn_params(gn_iters=12, pcg_iters=pcg, corr_mode=2))
    errs = []
    for it in range(12):
        d.GNStep()
        d.GNEnd(); errs.append(float(np.abs(d.Vertices() - targets).max()))
        d._gn_reopen = True
        # GNEnd closes the solve; reopen without resetting the state is not offered -> recompute via energy
        break
    d.close()
    # full run, error after all iterations for several counts
    row = []
    for its in (1, 2, 4, 8, 12):
        d = api.Deformation(pts, None, faces, ctx)
        d.SampleFPS(60, 0)
        d.SetTargets(targets)
        d.Solve(api.gn_params(gn_iters=its, pcg_iters=pcg, corr_mode=2))
        row.append(float(np.abs(d.Vertices() - targets).max()))
        d.close()
    print(pcg, ["%.2e" % e for e in row])
