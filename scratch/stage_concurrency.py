"""Which stage of a pair fails to overlap across host threads?  Each stage alone on N threads
(one context each): ms per round of N calls at N = 1 and N = 8 (perfect overlap: equal)."""
import sys, time, threading, numpy as np
sys.path.insert(0, '.')
from multiviewstitch_b200 import api, pipeline
ctx0 = api.Context(0)
sc = pipeline.make_scene(side=72)
ca, cb = api.make_camera(*sc["cam_a"]), api.make_camera(*sc["cam_b"])
m2d = api.Model2Depth(ctx0)
m2d.SetInput([sc["verts_a"], sc["verts_b"]], [None, None], [sc["faces"], sc["faces"]], [[ca], [cb]])
da, db = m2d.render(0)[0], m2d.render(1)[0]
pa, va = api.unproject_depth(da, ca, pipeline.MIN_DSP, pipeline.MAX_DSP, ctx0)
pb, vb = api.unproject_depth(db, cb, pipeline.MIN_DSP, pipeline.MAX_DSP, ctx0)
mt = pipeline.lattice_matches(pa, va, sc["cam_b"], vb, sc["planted"], 640, 480)
p, _ = api.unproject_pixels(da, ca, mt[:, 0:2], pipeline.MIN_DSP, pipeline.MAX_DSP, ctx0)
q, _ = api.unproject_pixels(db, cb, mt[:, 2:4], pipeline.MIN_DSP, pipeline.MAX_DSP, ctx0)
vtx, tex, fa = api.depth2model(da, ca, pipeline.MIN_DSP, pipeline.MAX_DSP, pipeline.SMOOTH_THRES, ctx0)
tgt = pb[np.flatnonzero(vb)]
prm = api.gn_params(gn_iters=4, pcg_iters=20, corr_mode=1)

def make(N):
    ctxs = [api.Context(0) for _ in range(N)]
    hs = []
    for c in ctxs:
        d = api.Deformation(vtx, None, fa, c); ids = d.SampleFPS(512, 0); d.BuildGraph(4, 8); d.SetTarget(tgt); d.Solve(prm)
        hs.append((d, ids))
    return ctxs, hs

stages = {
    "remove_outliers": lambda c, h: api.remove_outliers(p, q, ca, cb, 3, 200, 60.0, 0.75, 0, c),
    "depth2model": lambda c, h: api.depth2model(da, ca, pipeline.MIN_DSP, pipeline.MAX_DSP, pipeline.SMOOTH_THRES, c),
    "sample_fps": lambda c, h: h[0].SampleFPS(512, 0),
    "build_graph": lambda c, h: (h[0].SetNodes(h[1]), h[0].BuildGraph(4, 8)),
    "set_target_depth": lambda c, h: h[0].SetTargetDepth(db, cb, pipeline.MIN_DSP, pipeline.MAX_DSP),
    "solve": lambda c, h: h[0].Solve(prm),
    "vertices": lambda c, h: h[0].Vertices(),
}
res = {}
for N in (1, 8):
    ctxs, hs = make(N)
    for name, fn in stages.items():
        def work(k, reps):
            for _ in range(reps):
                fn(ctxs[k], hs[k])
        for k in range(N):
            work(k, 2)
        th = [threading.Thread(target=work, args=(k, 10)) for k in range(N)]
        t0 = time.perf_counter()
        for t in th: t.start()
        for t in th: t.join()
        res[(name, N)] = 1e3 * (time.perf_counter() - t0) / 10
    for d, _ in hs: d.close()
for name in stages:
    a, b = res[(name, 1)], res[(name, 8)]
    print(f"{name:18s} 1 thread {a:6.2f} ms   8 threads {b:6.2f} ms per round  (x{b / a:.1f})")
