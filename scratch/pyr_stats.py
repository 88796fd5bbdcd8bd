"""Walk statistics of the C3 first pass (rest vertices against the target) — needs the stats variant
library (scripts/build_variant.sh stats "-DMVS_PYR_STATS" nn_pyramid.cu) and MVS_PYR_DEBUG=1."""
import sys, numpy as np
sys.path.insert(0, '.')
from multiviewstitch_b200 import api, synth
nu = 409
verts, faces = synth.cube_sphere(nu)
pts = verts.astype(np.float32)
tv, _ = synth.cube_sphere(int(round(nu * 1.027)), rot=synth.axis_angle([1, 2, 3], 0.7))
tgt = synth.bend_twist(tv, 0.05).astype(np.float32)
ctx = api.Context(0)
ix = api.NNIndex(ctx)
import torch
t4 = torch.zeros((len(tgt), 4), device="cuda"); t4[:, :3] = torch.from_numpy(tgt).cuda()
q4 = torch.zeros((len(pts), 4), device="cuda"); q4[:, :3] = torch.from_numpy(pts).cuda()
oi = torch.empty(len(pts), dtype=torch.int32, device="cuda"); od = torch.empty(len(pts), dtype=torch.float32, device="cuda")
torch.cuda.synchronize()
ix.build_dev(t4.data_ptr(), len(tgt))
ix.knn_dev(q4.data_ptr(), len(pts), 1, oi.data_ptr(), od.data_ptr())
ctx.sync()
print("points tested per query:", oi.float().mean().item())
