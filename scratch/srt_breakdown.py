"""C2 SRT closed form on n float4 pairs: per-kernel times (CUDA events inside the library)."""
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
p4 = torch.zeros((n, 4), device="cuda"); p4[:, :3] = p
q4 = torch.zeros((n, 4), device="cuda"); q4[:, :3] = q
torch.cuda.synchronize()
ctx = api.Context(0)
srt = api.SRTSolver(10, ctx)
srt.SetPrintFlag(False)
srt.SetInputDev(p4.data_ptr(), q4.data_ptr(), n)
srt.EstimateTransform()
for rep in range(3):
    ctx.prof_enable(True); ctx.prof_reset()
    sc, Rr, tt = srt.EstimateTransform()
    ctx.sync()
    pr = ctx.prof()
print(n, {k: round(1e3 * v[0], 1) for k, v in pr.items()}, "us")
print("s = %.15g" % sc, "R00 = %.15g" % Rr[0, 0], "t0 = %.15g" % tt[0])
