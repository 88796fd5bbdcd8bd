// nn_pyramid.cuh — sparse occupancy pyramid for exact 1-NN / fixed-radius queries against a point
// set (the target scan of the correspondence search, Deformation.cpp:238-301 and
// PartRecognition.cpp:50-77 in the reference, which use FLANN kd-trees).
//
// Layout (all arrays dense per level, but tiny: the fine cell table of a 1M-point surface would be
// ~40M cells, the pyramid over it is < 15 MB and stays L2 resident):
//   fine cells of edge h            (never materialised; a point's cell = floor((p - bmin)/h))
//   level 1 "bricks"   4x4x4 cells   b1[brick] = {64-bit occupancy of its cells,
//                                                 index of its first occupied cell in occ_start}
//   level 2..4 nodes   4x4x4 children each, m2/m3/m4 64-bit occupancy of the children
//   occ_start[n_occ+1]               point ranges of the occupied cells, brick-major order
//   pts[n]                           float4 points sorted in the same order, .w = original index
// A query walks the four levels with the box of cells that can hold a point within the current
// bound: the box is AND-ed with the occupancy masks (empty space costs one 64-bit AND per node),
// the child holding (or nearest to) the query is taken first so the bound tightens at once, and
// nodes, bricks and cells are pruned by their distance to the query against the shrinking bound.
// Exactness contract as nn_grid.cuh: results equal brute force in FMA-free float32, ties to the
// lowest index — the structure only prunes regions that provably hold no better candidate.
#pragma once
#include "nn_grid.cuh"

#ifndef MVS_PYR_GROW
#define MVS_PYR_GROW 2.f
#endif

namespace mvs {

__device__ __forceinline__ unsigned long long box_mask(int xl, int xh, int yl, int yh, int zl,
                                                       int zh) {
  unsigned xm = ((2u << xh) - 1u) & ~((1u << xl) - 1u);
  unsigned ym = (0x1111u & ((2u << (4 * yh)) - 1u)) & ~((1u << (4 * yl)) - 1u);
  unsigned long long zm = (0x0001000100010001ull & ((2ull << (16 * zh)) - 1ull)) &
                          ~((1ull << (16 * zl)) - 1ull);
  return (unsigned long long)(xm * ym) * zm;
}

// squared distance from q to the axis-aligned box [lo, lo + s]^3 deflated by `mg`
__device__ __forceinline__ float aabb_d2(const float q[3], float lx, float ly, float lz, float s,
                                         float mg) {
  float gx = fmaxf(fmaxf(lx - q[0], q[0] - (lx + s)) - mg, 0.f);
  float gy = fmaxf(fmaxf(ly - q[1], q[1] - (ly + s)) - mg, 0.f);
  float gz = fmaxf(fmaxf(lz - q[2], q[2] - (lz + s)) - mg, 0.f);
  return gx * gx + gy * gy + gz * gz;
}

// Lower bound of the squared distance from q to any point of a set bounded by an oriented slab
// record {n (unit), R} {c, t}: every point p of the set satisfies |n.(p - c)| <= t and
// |(p - c) - (n.(p - c)) n| <= R (both inflated at build time).  With v = q - c:
//   |p - q|^2 = (n.(p - q))^2 + |P (p - q)|^2 >= max(|n.v| - t, 0)^2 + max(|P v| - R, 0)^2.
// A scan surface is a thin sheet inside a cell or brick: the axis-aligned box of the cell reaches
// half a cell towards a query above the sheet, the slab does not — for a query at distance d from the
// surface the box test keeps the ~2 pi d / h cells of a disc of radius sqrt(2 d h), the slab test
// the handful under the query.  `e` absorbs the float rounding of the evaluation.
__device__ __forceinline__ float slab_d2(const float q[3], float mg, float4 a, float4 b) {
  float vx = q[0] - b.x, vy = q[1] - b.y, vz = q[2] - b.z;
  float s = a.x * vx + a.y * vy + a.z * vz;
  float ux = vx - s * a.x, uy = vy - s * a.y, uz = vz - s * a.z;
  float lat = sqrtf(ux * ux + uy * uy + uz * uz);
  float e = 4e-6f * (fabsf(vx) + fabsf(vy) + fabsf(vz)) + mg;
  float gap = fmaxf(fabsf(s) - b.w - e, 0.f);
  float l = fmaxf(lat - a.w - e, 0.f);
  return gap * gap + l * l;
}

// state of one query's walk
struct PyrWalk {
  float q[3];
  float mg;     // conservative margin against the float rounding of the cell assignment
  float cap2;   // search radius^2 of this walk (INFINITY: the visitor's own bound only)
  int lo[3], hi[3];  // box of fine cells that can hold a point within the bound
  int hint[3];       // the query's own cell (clamped to the grid): ring centre
  float slab_r2;     // squared search radius above which the slab records are consulted
#ifdef MVS_PYR_STATS
  int st[8];         // walks, upper nodes, brick box tests, brick slab tests, cell box, cell slab, cells scanned, -
#endif
};
#ifdef MVS_PYR_STATS
template <class V>
__device__ __forceinline__ void pyr_stats_out(V&, const int*) {}
#define PYR_STAT(w, k) (++const_cast<PyrWalk&>(w).st[k])
#else
#define PYR_STAT(w, k) ((void)0)
#endif

// Children (bits of `occ`) of node (nx,ny,nz) that lie in the walk's box, nearest ring first.
// `sh`: log2 of the child edge in fine cells.  f(cx, cy, cz, bit) is called per child.
template <class F>
__device__ __forceinline__ void pyr_children(const PyrWalk& w, unsigned long long occ, int nx,
                                             int ny, int nz, int sh, F&& f) {
  int xl = max((w.lo[0] >> sh) - 4 * nx, 0), xh = min((w.hi[0] >> sh) - 4 * nx, 3);
  int yl = max((w.lo[1] >> sh) - 4 * ny, 0), yh = min((w.hi[1] >> sh) - 4 * ny, 3);
  int zl = max((w.lo[2] >> sh) - 4 * nz, 0), zh = min((w.hi[2] >> sh) - 4 * nz, 3);
  unsigned long long m = occ & box_mask(xl, xh, yl, yh, zl, zh);
  if (!m) return;
  int hx = min(max((w.hint[0] >> sh) - 4 * nx, 0), 3);
  int hy = min(max((w.hint[1] >> sh) - 4 * ny, 0), 3);
  int hz = min(max((w.hint[2] >> sh) - 4 * nz, 0), 3);
  // the child nearest to the query first (it tightens the bound at once), the rest in bit order
  int tc = (hz << 4) | (hy << 2) | hx;
  bool pend = (m >> tc) & 1ull;
  while (m) {
    int t = pend ? tc : (__ffsll((long long)m) - 1);
    pend = false;
    m &= ~(1ull << t);
    f(4 * nx + (t & 3), 4 * ny + ((t >> 2) & 3), 4 * nz + (t >> 4), t);
  }
}

template <class Visitor>
__device__ __forceinline__ void pyr_scan_cell(const PyrParams& P, unsigned long long m, int base,
                                              int bit, Visitor& vis, bool slabs = false,
                                              const float* q = nullptr, float mg = 0.f,
                                              float bound = 0.f) {
  int oi = base + __popcll(m & ((1ull << bit) - 1ull));
  if (slabs) {
    float4 a = __ldg(P.cslab + 2 * (size_t)oi), b = __ldg(P.cslab + 2 * (size_t)oi + 1);
    if (slab_d2(q, mg, a, b) > bound) return;
  }
  int s = __ldg(P.occ_start + oi), e = __ldg(P.occ_start + oi + 1);
  int i = s;
  for (; i + 2 <= e; i += 2) {  // two points in flight
    float4 p0 = __ldg(P.pts + i), p1 = __ldg(P.pts + i + 1);
    vis.point(p0);
    vis.point(p1);
  }
  if (i < e) vis.point(__ldg(P.pts + i));
}

// Scan the cells of one brick that lie in the box and can still hold a better candidate.
template <class Visitor>
__device__ __forceinline__ void pyr_scan_brick(const PyrParams& P, const PyrWalk& w, int x1, int y1,
                                               int z1, Visitor& vis) {
  float s1 = 4.f * P.h;
  PYR_STAT(w, 2);
  if (aabb_d2(w.q, P.bmin[0] + x1 * s1, P.bmin[1] + y1 * s1, P.bmin[2] + z1 * s1, s1, w.mg) >
      fminf(vis.worst2(), w.cap2))
    return;
  ulonglong2 rec = __ldg(P.b1 + ((size_t)z1 * P.d1[1] + y1) * P.d1[0] + x1);
  if (!rec.x) return;
  // slab bounds pay once the search radius exceeds about a cell (far queries: the first
  // Gauss-Newton iteration); below that the boxes prune as well and the record loads cost more
  const bool slabs = P.cslab != nullptr && fminf(vis.worst2(), w.cap2) > w.slab_r2;
  if (slabs) {
    PYR_STAT(w, 3);
    float4 a = __ldg(P.bslab + 2 * (size_t)rec.y), b = __ldg(P.bslab + 2 * (size_t)rec.y + 1);
    if (slab_d2(w.q, w.mg, a, b) > fminf(vis.worst2(), w.cap2)) return;
  }
  pyr_children(w, rec.x, x1, y1, z1, 0, [&](int cx, int cy, int cz, int t) {
    const float bound = fminf(vis.worst2(), w.cap2);
    PYR_STAT(w, 4);
#ifdef MVS_PYR_SKIP_CELL_BOX
    if (!slabs)
#endif
    if (aabb_d2(w.q, P.bmin[0] + cx * P.h, P.bmin[1] + cy * P.h, P.bmin[2] + cz * P.h, P.h, w.mg) >
        bound)
      return;
    if (slabs) PYR_STAT(w, 5);
    pyr_scan_cell(P, rec.x, (int)rec.y, t, vis, slabs, w.q, w.mg, bound);
  });
}

// One bounded walk: every point within sqrt(min(vis.worst2(), cap2)) of q (inclusive) is offered to
// the visitor.
template <class Visitor>
__device__ __forceinline__ void pyr_walk(const PyrParams& P, PyrWalk& w, float cap2, Visitor& vis) {
  w.cap2 = cap2;
  PYR_STAT(w, 0);
  float w2 = fminf(vis.worst2(), cap2);
  if (w2 < INFINITY) {
    float r = sqrtf(w2) * 1.000001f + w.mg;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      w.lo[a] = grid_coord(w.q[a] - r, P.bmin[a], P.inv_h, P.dims[a]);
      w.hi[a] = grid_coord(w.q[a] + r, P.bmin[a], P.inv_h, P.dims[a]);
    }
  } else {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      w.lo[a] = 0;
      w.hi[a] = P.dims[a] - 1;
    }
  }
  // small boxes (the seeded, near-surface case): at most 2 bricks per axis, go to them directly,
  // the query's own brick first
  if ((w.hi[0] >> 2) - (w.lo[0] >> 2) <= 1 && (w.hi[1] >> 2) - (w.lo[1] >> 2) <= 1 &&
      (w.hi[2] >> 2) - (w.lo[2] >> 2) <= 1) {
    int hb[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) hb[a] = min(max(w.hint[a] >> 2, w.lo[a] >> 2), w.hi[a] >> 2);
    pyr_scan_brick(P, w, hb[0], hb[1], hb[2], vis);
    for (int z1 = w.lo[2] >> 2; z1 <= (w.hi[2] >> 2); ++z1)
      for (int y1 = w.lo[1] >> 2; y1 <= (w.hi[1] >> 2); ++y1)
        for (int x1 = w.lo[0] >> 2; x1 <= (w.hi[0] >> 2); ++x1) {
          if (x1 == hb[0] && y1 == hb[1] && z1 == hb[2]) continue;
          pyr_scan_brick(P, w, x1, y1, z1, vis);
        }
    return;
  }
  const float s3 = 64.f * P.h, s2 = 16.f * P.h;
  for (int z4 = w.lo[2] >> 8; z4 <= (w.hi[2] >> 8); ++z4)
    for (int y4 = w.lo[1] >> 8; y4 <= (w.hi[1] >> 8); ++y4)
      for (int x4 = w.lo[0] >> 8; x4 <= (w.hi[0] >> 8); ++x4) {
        unsigned long long o4 = __ldg(P.m4 + ((size_t)z4 * P.d4[1] + y4) * P.d4[0] + x4);
        pyr_children(w, o4, x4, y4, z4, 6, [&](int x3, int y3, int z3, int) {
          PYR_STAT(w, 1);
          if (aabb_d2(w.q, P.bmin[0] + x3 * s3, P.bmin[1] + y3 * s3, P.bmin[2] + z3 * s3, s3, w.mg) >
              fminf(vis.worst2(), w.cap2))
            return;
          unsigned long long o3 = __ldg(P.m3 + ((size_t)z3 * P.d3[1] + y3) * P.d3[0] + x3);
          pyr_children(w, o3, x3, y3, z3, 4, [&](int x2, int y2, int z2, int) {
            PYR_STAT(w, 1);
            if (aabb_d2(w.q, P.bmin[0] + x2 * s2, P.bmin[1] + y2 * s2, P.bmin[2] + z2 * s2, s2,
                        w.mg) > fminf(vis.worst2(), w.cap2))
              return;
            unsigned long long o2 = __ldg(P.m2 + ((size_t)z2 * P.d2[1] + y2) * P.d2[0] + x2);
            pyr_children(w, o2, x2, y2, z2, 2,
                         [&](int x1, int y1, int z1, int) { pyr_scan_brick(P, w, x1, y1, z1, vis); });
          });
        });
      }
}

// Exact search: every point within sqrt(vis.worst2()) of q (inclusive) is offered to the visitor.
// A visitor without a finite bound (unseeded nearest-neighbour query) is served by bounded walks
// of doubling radius: a walk of radius r offers every point within r, so once the visitor's bound
// is <= r the answer is exact; empty space costs one masked AND per pyramid node per walk.
template <class Visitor>
__device__ __forceinline__ void pyr_search(const PyrParams& P, float qx, float qy, float qz,
                                           Visitor& vis) {
  PyrWalk w;
  w.q[0] = qx; w.q[1] = qy; w.q[2] = qz;
#pragma unroll
  for (int a = 0; a < 3; ++a) w.hint[a] = grid_coord(w.q[a], P.bmin[a], P.inv_h, P.dims[a]);
  w.mg = 1e-3f * P.h + 1e-6f * (fabsf(qx) + fabsf(qy) + fabsf(qz));
  w.slab_r2 = kPyrSlabRadius * kPyrSlabRadius * P.h * P.h;
#ifdef MVS_PYR_STATS
  for (int k = 0; k < 8; ++k) w.st[k] = 0;
  struct StatOut {
    PyrWalk& w;
    Visitor& v;
    __device__ ~StatOut() { pyr_stats_out(v, w.st); }
  } stat_out{w, vis};
#endif
  if (vis.worst2() < INFINITY) {
    pyr_walk(P, w, INFINITY, vis);
    return;
  }
  // distance from q to the grid's bounding box (queries outside start beyond it)
  float gx = fmaxf(fmaxf(P.bmin[0] - qx, qx - (P.bmin[0] + P.dims[0] * P.h)), 0.f);
  float gy = fmaxf(fmaxf(P.bmin[1] - qy, qy - (P.bmin[1] + P.dims[1] * P.h)), 0.f);
  float gz = fmaxf(fmaxf(P.bmin[2] - qz, qz - (P.bmin[2] + P.dims[2] * P.h)), 0.f);
  float out = sqrtf(gx * gx + gy * gy + gz * gz);
  float r = kPyrFirstRadius * P.h + out;
  const float rmax = P.h * (float)(P.dims[0] + P.dims[1] + P.dims[2]) + out;
  for (;;) {
    float r2 = r * r;
    pyr_walk(P, w, r2, vis);
    float w2 = vis.worst2();
    if (w2 < INFINITY) {
      // a candidate from a box corner (beyond r): one more walk with the visitor's own bound
      if (w2 > r2) pyr_walk(P, w, INFINITY, vis);
      return;
    }
    if (!(r <= rmax)) {  // nothing within reach of the whole grid: unbounded walk (NaN query)
      pyr_walk(P, w, INFINITY, vis);
      return;
    }
    r *= MVS_PYR_GROW;
  }
}

}  // namespace mvs
