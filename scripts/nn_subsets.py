import sys, numpy as np
sys.path.insert(0, '.')
from multiviewstitch_b200 import api, synth
nu = 1000
verts, faces = synth.uv_sphere(nu, nu + 1)
pts = verts.astype(np.float32)
tv, _ = synth.uv_sphere(int(round(nu * 1.103)), int(round((nu + 1) * 0.906)))
tgt = synth.bend_twist(tv, 0.05).astype(np.float32)
ctx = api.Context(0)
idx = api.NNIndex(ctx).build(tgt)
def run(name, q):
    idx.knn(q[:1000], 1)
    ctx.prof_enable(True); ctx.prof_reset()
    i, d2 = idx.knn(q, 1)
    p = ctx.prof(); ctx.prof_enable(False)
    print(f"{name:28s} n={len(q):8d}  nn1 {1e3*p['nn1_pyramid'][0]:9.1f} us   mean d/h... max d {np.sqrt(d2.max()):.4f} mean d {np.sqrt(d2).mean():.4f}")
run("all", pts)
run("every 16th", pts[::16])
z = pts[:, 2]
run("polar |z|>0.97", pts[np.abs(z) > 0.97])
run("mid 0.3<|z|<0.7", pts[(np.abs(z) > 0.3) & (np.abs(z) < 0.7)])
run("equator |z|<0.1", pts[np.abs(z) < 0.1])
run("on target (d=0)", tgt[::4])
rng = np.random.default_rng(0)
run("random in bbox", rng.uniform(tgt.min(0), tgt.max(0), size=(100000, 3)).astype(np.float32))
