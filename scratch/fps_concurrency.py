"""Do FPS cluster launches from different streams overlap?  N threads, one context + handle each."""
import sys, time, threading, numpy as np
sys.path.insert(0, '.')
from multiviewstitch_b200 import api, synth
verts, faces = synth.cube_sphere(76)  # ~34.7K points
pts = verts.astype(np.float32)
print(len(pts), "points")
for N in (1, 2, 4, 8, 12):
    ctxs = [api.Context(0) for _ in range(N)]
    hs = [api.Deformation(pts, None, None, c) for c in ctxs]
    def work(k, reps):
        for _ in range(reps):
            hs[k].SampleFPS(512, 0)
    for k in range(N):
        work(k, 2)
    th = [threading.Thread(target=work, args=(k, 10)) for k in range(N)]
    t0 = time.perf_counter()
    for t in th: t.start()
    for t in th: t.join()
    dt = time.perf_counter() - t0
    print(f"{N:2d} concurrent: {1e3 * dt / 10:.2f} ms per round of {N} samplings ({1e3 * dt / 10 / N:.2f} ms each)")
    for h in hs: h.close()
