"""Where do the 2.2 ms of one 640x480 unproject_depth call go?  (host-buffer C ABI)"""
import sys, time, ctypes as C, numpy as np
sys.path.insert(0, '.')
from multiviewstitch_b200 import api, pipeline, _capi
from multiviewstitch_b200.api import _ptr, _ctxh
ctx = api.Context(0)
sc = pipeline.make_scene(side=72)
ca = api.make_camera(*sc["cam_a"])
m2d = api.Model2Depth(ctx)
m2d.SetInput([sc["verts_a"]], [None], [sc["faces"]], [[ca]])
da = m2d.render(0)[0]
def T(name, f, reps=20):
    f(); ctx.sync()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    ctx.sync()
    print(f"{name:44s} {1e3*(time.perf_counter()-t0)/reps:7.3f} ms")
T("api.unproject_depth (fresh numpy outputs)", lambda: api.unproject_depth(da, ca, pipeline.MIN_DSP, pipeline.MAX_DSP, ctx))
a = np.ascontiguousarray(da, np.float32)
out = np.empty((ca.h * ca.w, 3), np.float64); valid = np.empty(ca.h * ca.w, np.int32)
out[:] = 0; valid[:] = 0
L = _capi.lib()
T("C call, preallocated + touched outputs", lambda: L.mvs_unproject_depth(_ctxh(ctx), _ptr(a, C.c_float), C.byref(ca), C.c_double(pipeline.MIN_DSP), C.c_double(pipeline.MAX_DSP), _ptr(out, C.c_double), _ptr(valid, C.c_int32)))
T("np.empty x2 + first touch", lambda: (np.empty((ca.h * ca.w, 3), np.float64).fill(0), np.empty(ca.h * ca.w, np.int32).fill(0)))
ctx.prof_enable(True); ctx.prof_reset()
L.mvs_unproject_depth(_ctxh(ctx), _ptr(a, C.c_float), C.byref(ca), C.c_double(pipeline.MIN_DSP), C.c_double(pipeline.MAX_DSP), _ptr(out, C.c_double), _ptr(valid, C.c_int32))
print({k: round(1e3 * v[0], 1) for k, v in ctx.prof().items()}, "us (kernel)")
