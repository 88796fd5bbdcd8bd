"""Numerical check: standard vs pipelined PCG with float32 vectors / float64 scalars on a real
normal-equation system of the embedded-deformation solve."""
import sys, numpy as np
sys.path.insert(0, '/root/repo')
import oracle
from tests import synth
verts, faces = synth.cube_sphere(40)      # 9602 verts
pts = verts.astype(np.float32)
m = int(sys.argv[1]) if len(sys.argv) > 1 else 300
ids = oracle.sample_fps(pts, m, 0)
nodes = pts[ids]
widx, w = oracle.node_weights(pts, nodes, 4)
graph = oracle.node_graph(nodes, 8)
targets = synth.bend_twist(verts, 0.05)
prm = oracle.gn_params(gn_iters=1, pcg_iters=20)
x0 = np.tile(np.array([1,0,0,0,1,0,0,0,1,0,0,0], np.float64), (m, 1))
H, g = oracle.gn_assemble_dense(pts, ids, widx, w, graph, targets, None, prm, x0)[:2]
H = np.asarray(H, np.float64); g = np.asarray(g, np.float64).ravel()
N = H.shape[0]
H = H + getattr(prm, 'lam', 0.0) * np.eye(N)
b = -g
# block-Jacobi 12x12
Minv = np.zeros_like(H)
for j in range(N // 12):
    sl = slice(12 * j, 12 * j + 12)
    Minv[sl, sl] = np.linalg.inv(H[sl, sl])
def std_pcg(dt, iters=20):
    A = H.astype(dt); Mi = Minv.astype(dt)
    x = np.zeros(N, dt); r = b.astype(dt); z = Mi @ r; p = z.copy()
    rz = float(np.dot(r.astype(np.float64), z.astype(np.float64)))
    for it in range(iters):
        q = A @ p
        pq = float(np.dot(p.astype(np.float64), q.astype(np.float64)))
        a = dt(rz / pq)
        x = x + a * p; r = r - a * q
        z = Mi @ r
        rzn = float(np.dot(r.astype(np.float64), z.astype(np.float64)))
        be = dt(rzn / rz); p = z + be * p; rz = rzn
    return x.astype(np.float64)
def pipe_pcg(dt, iters=20):
    A = H.astype(dt); Mi = Minv.astype(dt)
    x = np.zeros(N, dt); r = b.astype(dt); u = Mi @ r; w_ = A @ u
    z = np.zeros(N, dt); q = np.zeros(N, dt); s = np.zeros(N, dt); p = np.zeros(N, dt)
    gam_old = 1.0; al_old = 1.0
    for it in range(iters):
        gam = float(np.dot(r.astype(np.float64), u.astype(np.float64)))
        dl = float(np.dot(w_.astype(np.float64), u.astype(np.float64)))
        mm = Mi @ w_; nn = A @ mm
        if it > 0:
            be = gam / gam_old; al = gam / (dl - be * gam / al_old)
        else:
            be = 0.0; al = gam / dl
        be_, al_ = dt(be), dt(al)
        z = nn + be_ * z; q = mm + be_ * q; s = w_ + be_ * s; p = u + be_ * p
        x = x + al_ * p; r = r - al_ * s; u = u - al_ * q; w_ = w_ - al_ * z
        gam_old, al_old = gam, al
    return x.astype(np.float64)
ref = std_pcg(np.float64)
exact = np.linalg.solve(H, b)
for name, f, dt in [("std f32", std_pcg, np.float32), ("pipe f64", pipe_pcg, np.float64), ("pipe f32", pipe_pcg, np.float32)]:
    x = f(dt)
    print(f"{name:9s} |x - std64| / |std64| = {np.linalg.norm(x - ref) / np.linalg.norm(ref):.3e}   max abs {np.abs(x - ref).max():.3e}")
print("std64 vs exact solve:", np.linalg.norm(ref - exact) / np.linalg.norm(exact), " |delta| max", np.abs(ref).max(), " cond est", np.linalg.cond(Minv @ H))
def pipe_mixed(iters=20, round_m=True, matf32=True):
    A = H.astype(np.float32).astype(np.float64) if matf32 else H
    Mi = Minv.astype(np.float32).astype(np.float64) if matf32 else Minv
    dt = np.float64
    x = np.zeros(N); r = b.astype(np.float32).astype(dt); u = Mi @ r
    w_ = A @ (u.astype(np.float32).astype(dt) if round_m else u)
    z = np.zeros(N); q = np.zeros(N); s = np.zeros(N); p = np.zeros(N)
    gam_old = 1.0; al_old = 1.0
    for it in range(iters):
        gam = float(np.dot(r, u)); dl = float(np.dot(w_, u))
        mm = Mi @ w_
        mg = mm.astype(np.float32).astype(dt) if round_m else mm   # what neighbours read
        nn = A @ mg
        if it > 0:
            be = gam / gam_old; al = gam / (dl - be * gam / al_old)
        else:
            be = 0.0; al = gam / dl
        z = nn + be * z; q = mm + be * q; s = w_ + be * s; p = u + be * p
        x = x + al * p; r = r - al * s; u = u - al * q; w_ = w_ - al * z
        gam_old, al_old = gam, al
    return x
for name, x in [("pipe f64 vectors, f32 matrix + f32 halo", pipe_mixed(20, True, True)), ("pipe f64 vectors, f32 matrix, f64 halo", pipe_mixed(20, False, True))]:
    print(f"{name:42s} |x - std64| / |std64| = {np.linalg.norm(x - ref) / np.linalg.norm(ref):.3e}   max abs {np.abs(x - ref).max():.3e}")
