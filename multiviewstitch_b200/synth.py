"""Seeded synthetic inputs shared by the tests, bench.py and the golden-vector generator.

Shapes/constants follow SURVEY.md §8(d): C1 cameras 640x480, fx=fy=525, cx=319.5, cy=239.5; planted
similarity s=1.37, R=axis-angle(0.4 rad, (1,2,3)/sqrt(14)), t=(0.3,-0.2,0.5).
"""
from __future__ import annotations

import numpy as np

PLANTED_S = 1.37
PLANTED_T = np.array([0.3, -0.2, 0.5])


def axis_angle(axis, angle):
    a = np.asarray(axis, np.float64)
    a = a / np.linalg.norm(a)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * (K @ K)


PLANTED_R = axis_angle([1, 2, 3], 0.4)


def c1_intrinsics():
    return np.array([[525.0, 0, 319.5], [0, 525.0, 239.5], [0, 0, 1.0]])


def sphere_points(n, seed, noise=0.0, radius=1.0, center=(0, 0, 0)):
    rng = np.random.default_rng(seed)
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    v = v * radius + np.asarray(center)
    if noise:
        v = v + rng.normal(scale=noise, size=v.shape)
    return v


def planted_pairs(n, seed=1234, noise=1e-3, center=(0.0, 0.0, 5.0)):
    """C2-style pairs: p on a sphere (in front of camera 1), q = s R p + t (+ noise)."""
    rng = np.random.default_rng(seed + 1)
    p = sphere_points(n, seed, noise=noise, center=center)
    q = PLANTED_S * p @ PLANTED_R.T + PLANTED_T
    if noise:
        q = q + rng.normal(scale=noise, size=q.shape)
    return p, q


def camera_pair():
    """cam1 at the origin looking down +z; cam2 sees the transformed frame: x2 = s R x1 + t, so a
    camera with pose [R2|t2] = [R^T/.. ] keeps the scene in front.  We simply look at the mapped
    sphere centre from 5 units away along -z of the second frame."""
    K = c1_intrinsics()
    R1, t1 = np.eye(3), np.zeros(3)
    c2 = PLANTED_S * PLANTED_R @ np.array([0, 0, 5.0]) + PLANTED_T
    R2 = axis_angle([0, 1, 0], np.deg2rad(10.0))
    # place camera 2 so that the mapped centre sits at depth 5*s on its optical axis
    t2 = np.array([0, 0, 5.0 * PLANTED_S]) - R2 @ c2
    return (K, R1, t1, 640, 480), (K, R2, t2, 640, 480)


def uv_sphere(nu, nv, radius=1.0, bumps=0.05, seed=0):
    """Closed genus-0 triangle mesh (UV sphere with low-frequency bumps): (nu*(nv-1)+2) verts."""
    th = np.linspace(0, np.pi, nv + 1)[1:-1]
    ph = np.linspace(0, 2 * np.pi, nu, endpoint=False)
    T, P = np.meshgrid(th, ph, indexing="ij")
    r = radius * (1 + bumps * (np.sin(3 * T) * np.cos(2 * P) + 0.5 * np.cos(5 * T + 1.0) * np.sin(3 * P)))
    x = r * np.sin(T) * np.cos(P)
    y = r * np.sin(T) * np.sin(P)
    z = r * np.cos(T)
    ring = np.stack([x, y, z], -1).reshape(-1, 3)
    verts = np.concatenate([[[0, 0, radius]], ring, [[0, 0, -radius]]], 0)
    nr = nv - 1
    faces = []
    idx = lambda i, j: 1 + i * nu + (j % nu)
    j = np.arange(nu)
    # caps
    faces.append(np.stack([np.zeros(nu, int), idx(0, j), idx(0, j + 1)], 1))
    south = 1 + nr * nu
    faces.append(np.stack([np.full(nu, south), idx(nr - 1, j + 1), idx(nr - 1, j)], 1))
    if nr > 1:
        i = np.arange(nr - 1)[:, None]
        jj = j[None, :]
        a, b, c, d = idx(i, jj), idx(i, jj + 1), idx(i + 1, jj), idx(i + 1, jj + 1)
        quad = np.stack([np.stack([a, c, b], -1), np.stack([b, c, d], -1)], 2)  # nr-1, nu, 2, 3
        faces.append(quad.reshape(-1, 3))
    faces = np.concatenate(faces, 0).astype(np.int32)
    return verts.astype(np.float64), faces


def vertex_normals(verts, faces):
    v = verts
    n = np.cross(v[faces[:, 1]] - v[faces[:, 0]], v[faces[:, 2]] - v[faces[:, 0]])
    ln = np.linalg.norm(n, axis=1, keepdims=True)
    n = n / np.where(ln > 0, ln, 1)
    out = np.zeros_like(v)
    for k in range(3):
        for c in range(3):
            out[:, c] += np.bincount(faces[:, k], weights=n[:, c], minlength=len(v))
    ln = np.linalg.norm(out, axis=1, keepdims=True)
    return out / np.where(ln > 0, ln, 1)


def bend_twist(verts, amp=0.05):
    """Smooth bend+twist warp, amplitude as a fraction of the bbox diagonal (C3 target)."""
    v = np.asarray(verts, np.float64)
    diag = np.linalg.norm(v.max(0) - v.min(0))
    z = v[:, 2] / (np.abs(v[:, 2]).max() + 1e-12)
    ang = 0.6 * amp * 10 * z  # twist about z
    c, s = np.cos(ang), np.sin(ang)
    out = np.stack([c * v[:, 0] - s * v[:, 1], s * v[:, 0] + c * v[:, 1], v[:, 2]], 1)
    out[:, 0] += amp * diag * z * z  # bend
    return out


def grid_disk(radius, n, y, seed=0):
    """Flat triangulated n x n grid patch in the plane y = const (a 'ground'), clipped to a disk."""
    xs = np.linspace(-radius, radius, n)
    X, Z = np.meshgrid(xs, xs, indexing="ij")
    verts = np.stack([X.ravel(), np.full(X.size, y), Z.ravel()], 1)
    idx = np.arange(n * n).reshape(n, n)
    a, b, c, d = idx[:-1, :-1].ravel(), idx[1:, :-1].ravel(), idx[:-1, 1:].ravel(), idx[1:, 1:].ravel()
    faces = np.concatenate([np.stack([a, b, c], 1), np.stack([b, d, c], 1)], 0)
    keep = (verts[:, 0] ** 2 + verts[:, 2] ** 2) <= radius * radius
    remap = -np.ones(len(verts), np.int64)
    remap[keep] = np.arange(keep.sum())
    fk = keep[faces].all(1)
    return verts[keep], remap[faces[fk]].astype(np.int32)


def body_on_ground(seed=0):
    """Tall ellipsoid 'body' (y down, feet at +y) standing on a flat ground patch: the input shape
    of Alignment::RemoveGround.  Returns (verts, normals, faces, n_body)."""
    bv, bf = uv_sphere(48, 64, bumps=0.03, seed=seed)
    bv = bv[:, [0, 2, 1]] * np.array([0.35, 1.0, 0.22])   # long axis -> y
    bf = bf[:, [0, 2, 1]]
    gv, gf = grid_disk(0.7, 60, y=bv[:, 1].max() - 0.005)
    verts = np.concatenate([bv, gv], 0)
    faces = np.concatenate([bf, gf + len(bv)], 0).astype(np.int32)
    return verts, vertex_normals(verts, faces), faces, len(bv)


# PART enum of PartRecognition/PartRecognition.h:13-30
PARTS = dict(HEAD=0, NECK=1, LEFT_UPPER_ARM=2, LEFT_LOWER_ARM=3, LEFT_HAND=4, RIGHT_UPPER_ARM=5,
             RIGHT_LOWER_ARM=6, RIGHT_HAND=7, LEFT_THIGH=8, LEFT_SHANK=9, LEFT_FOOT=10,
             RIGHT_THIGH=11, RIGHT_SHANK=12, RIGHT_FOOT=13, TRUNCUS=14, HIP=15)


def humanoid(n_per_part=400, seed=0, arm_angle=0.5, leg_angle=0.12, scale=1.0):
    """Labelled point-cloud 'mean body' (y down, facing +z): 16 ellipsoidal parts with the labels of
    Template/part/parts.  arm_angle / leg_angle: how far the limbs are spread from the body axis."""
    rng = np.random.default_rng(seed)

    def blob(center, radii, label, axis=None):
        v = rng.normal(size=(n_per_part, 3))
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        v = v * np.asarray(radii)
        if axis is not None:  # rotate the local +y axis onto `axis`
            a = np.asarray(axis, np.float64)
            a = a / np.linalg.norm(a)
            yv = np.array([0, 1.0, 0])
            c = np.cross(yv, a)
            if np.linalg.norm(c) > 1e-12:
                v = v @ axis_angle(c, np.arccos(np.clip(yv @ a, -1, 1))).T
        return v + np.asarray(center), np.full(n_per_part, label, np.int32)

    P, L = [], []

    def add(*a, **k):
        p, l = blob(*a, **k)
        P.append(p); L.append(l)

    add([0, -0.78, 0], [0.10, 0.12, 0.10], 0)
    add([0, -0.62, 0], [0.05, 0.05, 0.05], 1)
    add([0, -0.35, 0], [0.20, 0.25, 0.11], 14)
    add([0, -0.05, 0], [0.19, 0.10, 0.11], 15)
    for side, (ua, la, ha) in ((+1, (2, 3, 4)), (-1, (5, 6, 7))):
        d = np.array([side * np.sin(arm_angle), np.cos(arm_angle), 0.0])
        sh = np.array([side * 0.24, -0.52, 0.0])
        add(sh + 0.14 * d, [0.05, 0.15, 0.05], ua, axis=d)
        add(sh + 0.42 * d, [0.04, 0.14, 0.04], la, axis=d)
        add(sh + 0.62 * d, [0.04, 0.06, 0.03], ha, axis=d)
    for side, (th, shk, ft) in ((+1, (8, 9, 10)), (-1, (11, 12, 13))):
        d = np.array([side * np.sin(leg_angle), np.cos(leg_angle), 0.0])
        hp = np.array([side * 0.10, 0.05, 0.0])
        add(hp + 0.22 * d, [0.08, 0.22, 0.08], th, axis=d)
        add(hp + 0.64 * d, [0.06, 0.21, 0.06], shk, axis=d)
        add(hp + 0.90 * d + np.array([0, 0, 0.05]), [0.05, 0.04, 0.10], ft)
    pts = np.concatenate(P, 0) * scale
    return pts, np.concatenate(L, 0)


def humanoid_mesh(nu=16, nv=12, arm_angle=0.5, leg_angle=0.12, ground=None):
    """The layout of `humanoid` as ONE connected triangle mesh: every part an ellipsoid mesh, joined to
    the previous part by a bridging facet (connectivity is all Alignment::RetainConnectRegion looks
    at).  ground = (radius, n): a flat grid patch under the feet, joined to a foot the same way — the
    input shape of Alignment::Align's target.  Returns (verts, faces, labels); ground label = -1."""
    sv, sf = uv_sphere(nu, nv, bumps=0.0)
    V, F, L = [], [], []
    off = 0

    def add(center, radii, label, axis=None):
        nonlocal off
        v = sv[:, [0, 2, 1]] * np.asarray(radii)  # the sphere's pole axis -> y
        f = sf[:, [0, 2, 1]]
        if axis is not None:
            a = np.asarray(axis, np.float64)
            a = a / np.linalg.norm(a)
            yv = np.array([0, 1.0, 0])
            c = np.cross(yv, a)
            if np.linalg.norm(c) > 1e-12:
                v = v @ axis_angle(c, np.arccos(np.clip(yv @ a, -1, 1))).T
        V.append(v + np.asarray(center))
        F.append(f + off)
        if off:
            F.append(np.array([[off - 1, off, off + 1]]))  # bridge to the previous part
        L.append(np.full(len(v), label, np.int32))
        off += len(v)

    add([0, -0.78, 0], [0.10, 0.12, 0.10], 0)
    add([0, -0.62, 0], [0.05, 0.05, 0.05], 1)
    add([0, -0.35, 0], [0.20, 0.25, 0.11], 14)
    add([0, -0.05, 0], [0.19, 0.10, 0.11], 15)
    for side, (ua, la, ha) in ((+1, (2, 3, 4)), (-1, (5, 6, 7))):
        d = np.array([side * np.sin(arm_angle), np.cos(arm_angle), 0.0])
        sh = np.array([side * 0.24, -0.52, 0.0])
        add(sh + 0.14 * d, [0.05, 0.15, 0.05], ua, axis=d)
        add(sh + 0.42 * d, [0.04, 0.14, 0.04], la, axis=d)
        add(sh + 0.62 * d, [0.04, 0.06, 0.03], ha, axis=d)
    for side, (th, shk, ft) in ((+1, (8, 9, 10)), (-1, (11, 12, 13))):
        d = np.array([side * np.sin(leg_angle), np.cos(leg_angle), 0.0])
        hp = np.array([side * 0.10, 0.05, 0.0])
        add(hp + 0.22 * d, [0.08, 0.22, 0.08], th, axis=d)
        add(hp + 0.64 * d, [0.06, 0.21, 0.06], shk, axis=d)
        add(hp + 0.90 * d + np.array([0, 0, 0.05]), [0.05, 0.04, 0.10], ft)
    if ground is not None:
        ymax = max(v[:, 1].max() for v in V)
        gv, gf = grid_disk(ground[0], ground[1], y=ymax - 0.005)
        V.append(gv)
        F.append(gf + off)
        F.append(np.array([[off - 1, off, off + 1]]))
        L.append(np.full(len(gv), -1, np.int32))
    return (np.concatenate(V, 0), np.concatenate(F, 0).astype(np.int32), np.concatenate(L, 0))


def cube_sphere(n, radius=1.0, bumps=0.05, rot=None):
    """Closed genus-0 triangle mesh with near-uniform vertex density: the 6 faces of a cube, each an
    n x n quad grid (tangent-warped), projected onto the sphere, with the same low-frequency bumps as
    uv_sphere.  6 n^2 + 2 vertices, 12 n^2 faces.  (A UV sphere packs ~1000x the mean density into
    its poles, which no scanner or Poisson mesh does.)"""
    t = np.tan(np.linspace(-np.pi / 4, np.pi / 4, n + 1))
    U, V = np.meshgrid(t, t, indexing="ij")
    one = np.ones_like(U)
    faces_xyz = [(one, U, V), (-one, V, U), (V, one, U), (U, -one, V), (U, V, one), (V, U, -one)]
    pts, quads = [], []
    for k, (X, Y, Z) in enumerate(faces_xyz):
        P = np.stack([X, Y, Z], -1).reshape(-1, 3)
        pts.append(P / np.linalg.norm(P, axis=1, keepdims=True))
        idx = (np.arange((n + 1) ** 2).reshape(n + 1, n + 1) + k * (n + 1) ** 2)
        a, b, c, d = idx[:-1, :-1], idx[1:, :-1], idx[:-1, 1:], idx[1:, 1:]
        quads.append(np.stack([np.stack([a, b, c], -1), np.stack([b, d, c], -1)], 2).reshape(-1, 3))
    P = np.concatenate(pts, 0)
    F = np.concatenate(quads, 0)
    # weld the duplicated edge / corner vertices (exact grid positions: round to a fine lattice)
    key = np.round(P * (1 << 20)).astype(np.int64) + (1 << 20)  # three 21-bit lattice coordinates
    key = (key[:, 0] << 42) | (key[:, 1] << 21) | key[:, 2]      # packed: a 1-D unique is ~10x faster
    _, first, inv = np.unique(key, return_index=True, return_inverse=True)
    order = np.argsort(first)               # keep the original vertex order (face-major)
    rank = np.empty_like(order)
    rank[order] = np.arange(len(order))
    V3 = P[first[order]]
    F = rank[inv.reshape(-1)][F]
    # consistent outward orientation
    c = V3[F].mean(1)
    nrm = np.cross(V3[F[:, 1]] - V3[F[:, 0]], V3[F[:, 2]] - V3[F[:, 0]])
    flip = np.sum(nrm * c, 1) < 0
    F[flip] = F[flip][:, [0, 2, 1]]
    if rot is not None:  # turn the tessellation, not the shape: the bumps are a function of position
        V3 = V3 @ np.asarray(rot, np.float64).T
    th = np.arccos(np.clip(V3[:, 2], -1, 1))
    ph = np.arctan2(V3[:, 1], V3[:, 0])
    r = radius * (1 + bumps * (np.sin(3 * th) * np.cos(2 * ph) + 0.5 * np.cos(5 * th + 1.0) * np.sin(3 * ph)))
    return (V3 * r[:, None]).astype(np.float64), F.astype(np.int32)
