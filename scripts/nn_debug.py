import os, sys, numpy as np
os.environ["MVS_PYR_DEBUG"] = "1"
sys.path.insert(0, '.')
from multiviewstitch_b200 import api, synth
nu = 1000
verts, faces = synth.uv_sphere(nu, nu + 1)
pts = verts.astype(np.float32)
tv, _ = synth.uv_sphere(int(round(nu * 1.103)), int(round((nu + 1) * 0.906)))
tgt = synth.bend_twist(tv, 0.05).astype(np.float32)
ctx = api.Context(0)
idx = api.NNIndex(ctx).build(tgt)
q = pts[::16]
cnt, clk = idx.knn(q, 1)
cnt = cnt.ravel(); clk = clk.ravel()
print("points evaluated: mean %.1f median %.0f p99 %.0f max %d" % (cnt.mean(), np.median(cnt), np.percentile(cnt, 99), cnt.max()))
print("clocks: mean %.0f median %.0f p99 %.0f max %.0f" % (clk.mean(), np.median(clk), np.percentile(clk, 99), clk.max()))
o = np.argsort(-clk)[:8]
for i in o:
    print("  q", q[i], "pts", cnt[i], "clk", clk[i])
del idx
ctx.close()
