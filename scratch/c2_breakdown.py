"""C2 (n float4 points, exact 1-NN of p' = sRp+t against shuffled q): family breakdown of one call."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from multiviewstitch_b200 import api, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
g = torch.Generator(device="cuda").manual_seed(1234)
p = torch.randn((n, 3), device="cuda", generator=g)
p = p / p.norm(dim=1, keepdim=True) + 1e-3 * torch.randn((n, 3), device="cuda", generator=g)
p[:, 2] += 5.0
R = torch.tensor(synth.PLANTED_R, dtype=torch.float32, device="cuda")
t = torch.tensor(synth.PLANTED_T, dtype=torch.float32, device="cuda")
q = synth.PLANTED_S * p @ R.T + t + 1e-3 * torch.randn((n, 3), device="cuda", generator=g)
q4 = torch.zeros((n, 4), device="cuda"); q4[:, :3] = q
shuf = torch.randperm(n, device="cuda", generator=g)
q4s = q4[shuf].contiguous()
query = torch.zeros((n, 4), device="cuda")
query[:, :3] = synth.PLANTED_S * p @ R.T + t
idx_out = torch.empty(n, dtype=torch.int32, device="cuda")
d2_out = torch.empty(n, dtype=torch.float32, device="cuda")
torch.cuda.synchronize()
ctx = api.Context(0)
index = api.NNIndex(ctx)
index.build_dev(q4s.data_ptr(), n)
index.knn_dev(query.data_ptr(), n, 1, idx_out.data_ptr(), d2_out.data_ptr())
ctx.sync()
for rep in range(2):
    ctx.prof_enable(True); ctx.prof_reset()
    index.knn_dev(query.data_ptr(), n, 1, idx_out.data_ptr(), d2_out.data_ptr())
    ctx.sync()
    pr = ctx.prof()
print(n, {k: round(1e3 * v[0], 1) for k, v in sorted(pr.items(), key=lambda kv: -kv[1][0])}, "us; total", round(1e3 * sum(v[0] for v in pr.values()), 1))
print("checksum", int(idx_out.long().sum().item()), float(d2_out.double().sum().item()))
