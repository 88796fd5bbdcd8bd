import sys, time, threading, numpy as np
sys.path.insert(0, '.')
from multiviewstitch_b200 import api, synth
verts, faces = synth.cube_sphere(76)
pts = verts.astype(np.float32)
for N in (1, 8):
    ctxs = [api.Context(0) for _ in range(N)]
    hs = [api.Deformation(pts, None, None, c) for c in ctxs]
    def work(k, reps):
        for _ in range(reps):
            hs[k].SampleFPS(512, 0)
    for k in range(N):
        work(k, 2)
    for c in ctxs:
        c.prof_enable(True); c.prof_reset()
    th = [threading.Thread(target=work, args=(k, 10)) for k in range(N)]
    t0 = time.perf_counter()
    for t in th: t.start()
    for t in th: t.join()
    dt = time.perf_counter() - t0
    pr = ctxs[0].prof()
    print(N, "concurrent:", f"{1e3*dt/10:.2f} ms per round;", {k: round(1e3 * v[0] / 10, 1) for k, v in sorted(pr.items(), key=lambda kv: -kv[1][0])[:6]})
    for h in hs: h.close()
