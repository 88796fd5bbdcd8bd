// nn_grid.cuh — uniform-grid exact nearest-neighbour structure: device-side traversal shared by the
// kNN kernels (nn_grid.cu), the per-vertex node weights and the node correspondence stage
// (deform_graph.cu).
//
// Exactness contract (include/mvs_b200.h, "exact kNN"): results equal brute force over the same
// float32 points with d2 = (dx*dx + dy*dy) + dz*dz evaluated in round-to-nearest without FMA and
// ties resolved by the lowest point index.  The grid only prunes: ring r of cells around the
// query's cell is scanned completely, and the search stops after a ring only when the k-th best
// distance is strictly below a conservative lower bound on the distance to every unscanned cell.
#pragma once
#include "common.cuh"

namespace mvs {

struct GridParams {
  float bmin[3];
  float h;      // cell edge
  float inv_h;
  int dims[3];
  int64_t ncells;
  int64_t n;  // points
};

// d2 with a fixed, FMA-free evaluation order (bit-identical to oracle/nn.cpp: dist2_f32).
__device__ __forceinline__ float dist2_rn(float ax, float ay, float az, float bx, float by,
                                          float bz) {
  float dx = __fsub_rn(ax, bx);
  float dy = __fsub_rn(ay, by);
  float dz = __fsub_rn(az, bz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

__device__ __forceinline__ int grid_coord(float x, float bmin, float inv_h, int dim) {
  float f = (x - bmin) * inv_h;
  int c = (f > 0.f) ? (int)fminf(f, 2.0e9f) : 0;  // NaN -> 0
  return c < dim - 1 ? c : dim - 1;
}

__device__ __forceinline__ int64_t grid_cell(const GridParams& g, float x, float y, float z) {
  int cx = grid_coord(x, g.bmin[0], g.inv_h, g.dims[0]);
  int cy = grid_coord(y, g.bmin[1], g.inv_h, g.dims[1]);
  int cz = grid_coord(z, g.bmin[2], g.inv_h, g.dims[2]);
  return ((int64_t)cz * g.dims[1] + cy) * g.dims[0] + cx;
}

// Lower bound on the distance from q to any point stored in a cell outside the cube of
// Chebyshev radius r around cell c (conservative by 1e-3 cells against float rounding of the
// cell assignment).  Returns +inf when the cube already covers the whole grid.
__device__ __forceinline__ float ring_lower_bound(const GridParams& g, const float q[3],
                                                  const int c[3], int r) {
  float best = INFINITY;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    if (c[a] - r > 0) {
      float plane = g.bmin[a] + (float)(c[a] - r) * g.h;
      best = fminf(best, q[a] - plane);
    }
    if (c[a] + r < g.dims[a] - 1) {
      float plane = g.bmin[a] + (float)(c[a] + r + 1) * g.h;
      best = fminf(best, plane - q[a]);
    }
  }
  if (best == INFINITY) return INFINITY;
  best -= 1e-3f * g.h + 1e-6f * (fabsf(q[0]) + fabsf(q[1]) + fabsf(q[2]));
  return best;
}

// Visit every point of ring r (cells at Chebyshev distance exactly r from c).  `pts` is the
// cell-sorted array; .w carries the original index (int bits).
template <class Visitor>
__device__ __forceinline__ void grid_visit_ring(const GridParams& g,
                                                const int* __restrict__ cell_start,
                                                const float4* __restrict__ pts, const int c[3],
                                                int r, Visitor& vis) {
  int z0 = max(c[2] - r, 0), z1 = min(c[2] + r, g.dims[2] - 1);
  int y0 = max(c[1] - r, 0), y1 = min(c[1] + r, g.dims[1] - 1);
  int x0 = max(c[0] - r, 0), x1 = min(c[0] + r, g.dims[0] - 1);
  for (int z = z0; z <= z1; ++z) {
    bool zface = (z == c[2] - r) || (z == c[2] + r);
    for (int y = y0; y <= y1; ++y) {
      bool yface = (y == c[1] - r) || (y == c[1] + r);
      int64_t row = ((int64_t)z * g.dims[1] + y) * g.dims[0];
      if (zface || yface) {
        int s = __ldg(cell_start + row + x0);
        int e = __ldg(cell_start + row + x1 + 1);
        for (int i = s; i < e; ++i) vis.point(__ldg(pts + i));
      } else {
        // interior rows: only the two end cells belong to the ring
        if (c[0] - r >= 0) {
          int s = __ldg(cell_start + row + (c[0] - r));
          int e = __ldg(cell_start + row + (c[0] - r) + 1);
          for (int i = s; i < e; ++i) vis.point(__ldg(pts + i));
        }
        if (r > 0 && c[0] + r <= g.dims[0] - 1) {
          int s = __ldg(cell_start + row + (c[0] + r));
          int e = __ldg(cell_start + row + (c[0] + r) + 1);
          for (int i = s; i < e; ++i) vis.point(__ldg(pts + i));
        }
      }
    }
  }
}

// Full search: scan rings until vis.worst2() (the pruning radius^2) is strictly below the
// lower bound of the unscanned region.
template <class Visitor>
__device__ __forceinline__ void grid_search(const GridParams& g,
                                            const int* __restrict__ cell_start,
                                            const float4* __restrict__ pts, float qx, float qy,
                                            float qz, Visitor& vis) {
  float q[3] = {qx, qy, qz};
  int c[3];
  c[0] = grid_coord(qx, g.bmin[0], g.inv_h, g.dims[0]);
  c[1] = grid_coord(qy, g.bmin[1], g.inv_h, g.dims[1]);
  c[2] = grid_coord(qz, g.bmin[2], g.inv_h, g.dims[2]);
  int rmax = max(max(g.dims[0], g.dims[1]), g.dims[2]);
  for (int r = 0; r <= rmax; ++r) {
    grid_visit_ring(g, cell_start, pts, c, r, vis);
    float lb = ring_lower_bound(g, q, c, r);
    if (lb == INFINITY) break;
    if (lb > 0.f && vis.worst2() < lb * lb) break;
  }
}

// Sorted (d2, idx) list of the K best, fully unrolled so it lives in registers.  Empty slots hold
// (+inf, INT_MAX).  Order is lexicographic on (d2, idx): the lowest index wins ties.
template <int K>
struct KnnList {
  float d2[K];
  int idx[K];
  float qx, qy, qz;
  __device__ __forceinline__ void init(float x, float y, float z) {
    qx = x;
    qy = y;
    qz = z;
#pragma unroll
    for (int i = 0; i < K; ++i) {
      d2[i] = INFINITY;
      idx[i] = 0x7fffffff;
    }
  }
  __device__ __forceinline__ float worst2() const { return d2[K - 1]; }
  __device__ __forceinline__ void insert(float d, int id) {
    bool better = (d < d2[K - 1]) || (d == d2[K - 1] && id < idx[K - 1]);
    if (!better) return;
    d2[K - 1] = d;
    idx[K - 1] = id;
#pragma unroll
    for (int i = K - 1; i >= 1; --i) {
      bool sw = (d2[i] < d2[i - 1]) || (d2[i] == d2[i - 1] && idx[i] < idx[i - 1]);
      if (sw) {
        float td = d2[i];
        d2[i] = d2[i - 1];
        d2[i - 1] = td;
        int ti = idx[i];
        idx[i] = idx[i - 1];
        idx[i - 1] = ti;
      }
    }
  }
  __device__ __forceinline__ void point(float4 p) {
    insert(dist2_rn(p.x, p.y, p.z, qx, qy, qz), __float_as_int(p.w));
  }
};

// First radius, in cells, of an unseeded query's doubling schedule (nn_pyramid.cuh: pyr_search).
// Measured on B200, 1e7 near-surface queries against 1e7 points (C2): 2.0 -> 5.04 ms, 1.0 -> 2.73,
// 0.75 -> 2.43, 0.5 -> 2.66, 0.35 -> 3.02; far queries (first GN iteration of C3) do not care.
constexpr float kPyrFirstRadius = 0.75f;
// Points per occupied cell wanted by the two users of the pyramid: near-surface 1-NN queries are
// fastest with ~3 (C2: 2.22 ms at 3, 2.29 at 2, 2.31 at 4.5); the deformation target, whose first
// iteration queries from up to 5 % of the diagonal away, with ~6 (C3 correspondence passes of one
// solve: 2.82 ms at 3, 2.60 at 4.5, 2.52 at 6, 2.72 at 8, 2.63 at 12).
// search radius, in cells, above which a walk consults the oriented slab records
#ifndef MVS_PYR_SLAB_R
#define MVS_PYR_SLAB_R 1.5f
#endif
constexpr float kPyrSlabRadius = MVS_PYR_SLAB_R;
constexpr float kPyrPpcNear = 3.0f;
#ifndef MVS_PYR_PPC_FAR
#define MVS_PYR_PPC_FAR 6.0f
#endif
constexpr float kPyrPpcFar = MVS_PYR_PPC_FAR;

// sparse occupancy pyramid (nn_pyramid.cuh)
struct PyrParams {
  float bmin[3];
  float h, inv_h;
  int dims[3];    // fine cells
  int d1[3];      // bricks
  int d2[3], d3[3], d4[3];
  const ulonglong2* b1;  // per brick: .x occupancy mask of its cells, .y first occupied-cell index
  const unsigned long long* m2;
  const unsigned long long* m3;
  const unsigned long long* m4;
  const int* occ_start;
  const float4* pts;
  int64_t n;
  // oriented slab bounds (nn_pyramid.cuh: slab_d2): two float4 per record, {unit normal, lateral
  // radius} {centre, half thickness}.  cslab[2 oi] bounds the points of occupied cell oi; bslab[2 oi]
  // bounds ALL points of the brick whose first occupied cell is oi.  Null = boxes only.
  const float4* cslab;
  const float4* bslab;
};

}  // namespace mvs

// Device-resident search structure.
struct mvs_nn_index {
  mvs_ctx* ctx = nullptr;
  mvs::GridParams g{};
  mvs::DevBuf<float4> pts_in;      // staging when built from host memory
  mvs::DevBuf<float4> pts_sorted;  // cell-sorted, .w = original index
  mvs::DevBuf<int> cell_start;     // ncells + 1
  mvs::DevBuf<int> cell_fill;      // scratch
  mvs::DevBuf<float> red;          // bbox reduction scratch
  mvs::DevBuf<int> scan_tmp;
  // query-side scratch
  mvs::DevBuf<float4> q_in;
  mvs::DevBuf<int> q_perm;
  mvs::DevBuf<int> q_cell;
  mvs::DevBuf<int> q_hist;
  mvs::DevBuf<int> o_idx;
  mvs::DevBuf<float> o_d2;
  bool built = false;
  // sparse occupancy pyramid (exact 1-NN / fixed-radius queries; nn_pyramid.cu)
  mvs::PyrParams pyr{};
  bool pyr_built = false;
  int pyr_n_occ = 0;
  float pyr_last_h = 0.f;
  int64_t pyr_last_n = 0;
  double pyr_last_ext = 0.0;
  mvs::DevBuf<unsigned long long> pyr_keys, pyr_keys2, pyr_m2, pyr_m3, pyr_m4;
  mvs::DevBuf<ulonglong2> pyr_b1;
  mvs::DevBuf<int> pyr_vals, pyr_order, pyr_flag, pyr_occ;
  mvs::DevBuf<float4> pyr_pts, pyr_cslab, pyr_bslab;
  mvs::DevBuf<char> pyr_tmp;
};

namespace mvs {
// Build from device float4 points (w ignored).  target_ppc: wanted points per occupied cell.
int nn_build(mvs_nn_index* idx, const float4* pts_dev, int64_t n, float target_ppc = 3.0f);
// Sparse pyramid over device float4 points (w ignored unless w_is_index).
int nn_build_pyramid(mvs_nn_index* idx, const float4* pts_dev, int64_t n,
                     float target_ppc = kPyrPpcNear,
                     bool w_is_index = false);
int nn_nn1_pyramid(mvs_nn_index* idx, const float4* q_dev, const int* perm_dev, int64_t nq,
                   int* out_idx_dev, float* out_d2_dev);
// Exact kNN, device pointers. out_idx: nq*k, out_d2: nq*k (either may be null).
int nn_knn(mvs_nn_index* idx, const float4* q_dev, int64_t nq, int k, int* out_idx_dev,
           float* out_d2_dev);
// Exclusive scan of `n` ints in place (returns total in *total_host if non-null; syncs then).
int exclusive_scan_inplace(mvs_ctx* ctx, int* data_dev, int64_t n, DevBuf<int>& tmp);
// xyz (float or double, host) -> float4 device upload with on-device conversion
int upload_xyz_f32(mvs_ctx* ctx, const float* xyz_host, int64_t n, DevBuf<float4>& out);
int upload_xyz_f64(mvs_ctx* ctx, const double* xyz_host, int64_t n, DevBuf<float4>& out);
}  // namespace mvs
