"""mvs_stitch_pairs throughput on one GPU for several worker counts (28 pairs, C4 workload)."""
import sys, time, numpy as np
sys.path.insert(0, '.')
from multiviewstitch_b200 import api, pipeline
ctx = api.Context(0)
work = []
m2d = api.Model2Depth(ctx)
for u in range(int(sys.argv[1]) if len(sys.argv) > 1 else 28):
    sc = pipeline.make_scene(side=72, view_angle_deg=10.0 + 1.5 * u, seed=0)
    ca, cb = api.make_camera(*sc["cam_a"]), api.make_camera(*sc["cam_b"])
    m2d.SetInput([sc["verts_a"], sc["verts_b"]], [None, None], [sc["faces"], sc["faces"]], [[ca], [cb]])
    da, db = m2d.render(0)[0], m2d.render(1)[0]
    pa, va = api.unproject_depth(da, ca, pipeline.MIN_DSP, pipeline.MAX_DSP, ctx)
    pb, vb = api.unproject_depth(db, cb, pipeline.MIN_DSP, pipeline.MAX_DSP, ctx)
    mt = pipeline.lattice_matches(pa, va, sc["cam_b"], vb, sc["planted"], 640, 480, seed=u)
    work.append((da, ca, db, cb, mt, u))
for outputs in (True, False):
    batch = api.PairBatch(ctx, want_outputs=outputs)
    for da, ca, db, cb, mt, u in work:
        batch.add(da, ca, db, cb, mt, seed=u)
    for W in (1, 2, 4, 8, 12, 16):
        prm = api.pair_params(n_nodes=512, gn_iters=4, workers=W)
        for _ in range(2):
            batch.run(prm)
        t0 = time.perf_counter()
        R = 3
        for _ in range(R):
            outs = batch.run(prm)
        dt = (time.perf_counter() - t0) / R
        print(f"outputs={outputs} workers={W:2d}: {len(work) / dt:7.1f} pairs/s  ({1e3 * dt:.1f} ms per {len(work)} pairs)", flush=True)
