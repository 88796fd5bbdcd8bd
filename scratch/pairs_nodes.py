import sys, time, numpy as np
sys.path.insert(0, '.')
from multiviewstitch_b200 import api, pipeline
ctx = api.Context(0)
work = []
m2d = api.Model2Depth(ctx)
for u in range(28):
    sc = pipeline.make_scene(side=72, view_angle_deg=10.0 + 1.5 * u, seed=0)
    ca, cb = api.make_camera(*sc["cam_a"]), api.make_camera(*sc["cam_b"])
    m2d.SetInput([sc["verts_a"], sc["verts_b"]], [None, None], [sc["faces"], sc["faces"]], [[ca], [cb]])
    da, db = m2d.render(0)[0], m2d.render(1)[0]
    pa, va = api.unproject_depth(da, ca, pipeline.MIN_DSP, pipeline.MAX_DSP, ctx)
    pb, vb = api.unproject_depth(db, cb, pipeline.MIN_DSP, pipeline.MAX_DSP, ctx)
    mt = pipeline.lattice_matches(pa, va, sc["cam_b"], vb, sc["planted"], 640, 480, seed=u)
    work.append((da, ca, db, cb, mt, u))
batch = api.PairBatch(ctx, want_outputs=True)
for da, ca, db, cb, mt, u in work:
    batch.add(da, ca, db, cb, mt, seed=u)
for nodes, gn in ((512, 4), (128, 4), (32, 4), (32, 1)):
    for W in (1, 12):
        prm = api.pair_params(n_nodes=nodes, gn_iters=gn, workers=W)
        for _ in range(2):
            batch.run(prm)
        t0 = time.perf_counter()
        for _ in range(3):
            batch.run(prm)
        dt = (time.perf_counter() - t0) / 3
        print(f"nodes={nodes} gn={gn} workers={W:2d}: {28 / dt:7.1f} pairs/s", flush=True)
