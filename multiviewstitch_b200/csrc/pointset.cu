// pointset.cu — point-set reductions and the rigid+scale pre-alignment on the device:
//   PointSetUtils (SetUtils/PointSetUtils.cpp:3-62): barycenter, bounding box, PCA pivots
//   Alignment::Align / RemoveGround / InitAlignment / LocalAlignment(Core)
//                                           (Alignment/Alignment.cpp:11-76, 79-233, 235-314, 316-546)
//   per-point similarity transforms (Alignment.cpp:31-34, Processor.cpp:1025-1026, 1177-1189)
//   RetainConnectRegion (Alignment.cpp:618-654) — host union-find, sequential by nature.
//
// Every O(N) sweep of the reference is one streaming kernel here (16-24 B/point/pass, double
// accumulators, fixed reduction trees from reduce.cuh so results are run-to-run reproducible); the
// 3x3 eigen/inverse algebra and the control flow between sweeps stay on the host, exactly where the
// reference has them.  Points stay in the caller's layout (n x 3 doubles) on the device: a warp reads
// 768 contiguous bytes per pass, every byte of which is used.
//
// Conventions where the reference is unpinned (same as oracle/align.cpp): pivots are sign-normalised
// (largest-magnitude component positive) before the reference's own flips; max-reductions start from
// lowest(); LocalAlignmentCore erases the smallest missing label; CalcRotation is clamped.
#include <cmath>
#include <cstring>
#include <limits>
#include <numeric>
#include <vector>

#include "linalg3.h"
#include "nn_grid.cuh"
#include "reduce.cuh"

namespace mvs {

constexpr int kPsThreads = 256;

// selection predicate: labels == nullptr -> all points; eq >= 0 -> labels[i] == eq; else bit test
struct Sel {
  const int* labels;
  unsigned mask;
  int eq;
  __device__ __forceinline__ bool operator()(int64_t i) const {
    if (!labels) return true;
    int l = __ldg(labels + i);
    if (eq >= 0) return l == eq;
    return l >= 0 && l <= 31 && ((mask >> l) & 1u);
  }
};

__device__ __forceinline__ void load3(const double* __restrict__ p, int64_t i, double v[3]) {
  v[0] = __ldg(p + 3 * i);
  v[1] = __ldg(p + 3 * i + 1);
  v[2] = __ldg(p + 3 * i + 2);
}

// grid-wide max of N doubles (same last-CTA pattern as grid_reduce_sum)
template <int N>
__device__ __forceinline__ bool grid_reduce_max(double (&v)[N], double* smem, double* partials,
                                                unsigned* ticket, double* out) {
  __shared__ bool is_last;
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int j = 0; j < N; ++j)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[j] = fmax(v[j], __shfl_xor_sync(0xffffffffu, v[j], o));
  if (lane == 0)
#pragma unroll
    for (int j = 0; j < N; ++j) smem[warp * N + j] = v[j];
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < nw; ++w)
#pragma unroll
      for (int j = 0; j < N; ++j) v[j] = fmax(v[j], smem[w * N + j]);
#pragma unroll
    for (int j = 0; j < N; ++j) partials[(size_t)blockIdx.x * N + j] = v[j];
    __threadfence();
    is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return false;
  __threadfence();
  if (threadIdx.x < N) {
    double s = -INFINITY;
    for (unsigned b = 0; b < gridDim.x; ++b) s = fmax(s, __ldcg(partials + (size_t)b * N + threadIdx.x));
    out[threadIdx.x] = s;
  }
  if (threadIdx.x == 0) *ticket = 0u;
  __threadfence();
  __syncthreads();
  return true;
}

// pass 1: sum, count  -> out[0..3]; pass 1b: bbox -> out[4..9] as (-min, max)
__global__ void __launch_bounds__(kPsThreads)
    k_ps_sum(const double* __restrict__ pts, int64_t n, Sel sel, double* partials,
             unsigned* ticket, double* out) {
  __shared__ double smem[32 * 4];
  double v[4] = {0, 0, 0, 0};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    if (!sel(i)) continue;
    double p[3];
    load3(pts, i, p);
    v[0] += p[0]; v[1] += p[1]; v[2] += p[2]; v[3] += 1.0;
  }
  grid_reduce_sum<4>(v, smem, partials, ticket, out);
}

__global__ void __launch_bounds__(kPsThreads)
    k_ps_bbox(const double* __restrict__ pts, int64_t n, Sel sel, double* partials,
              unsigned* ticket, double* out) {
  __shared__ double smem[32 * 6];
  double v[6] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY, -INFINITY, -INFINITY};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    if (!sel(i)) continue;
    double p[3];
    load3(pts, i, p);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      v[a] = fmax(v[a], -p[a]);
      v[3 + a] = fmax(v[3 + a], p[a]);
    }
  }
  grid_reduce_max<6>(v, smem, partials, ticket, out);
}

// pass 2: sum (p-c)(p-c)^T, upper triangle xx xy xz yy yz zz
__global__ void __launch_bounds__(kPsThreads)
    k_ps_cov(const double* __restrict__ pts, int64_t n, Sel sel, double c0, double c1, double c2,
             double* partials, unsigned* ticket, double* out) {
  __shared__ double smem[32 * 6];
  double v[6] = {0, 0, 0, 0, 0, 0};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    if (!sel(i)) continue;
    double p[3];
    load3(pts, i, p);
    double d0 = p[0] - c0, d1 = p[1] - c1, d2 = p[2] - c2;
    v[0] += d0 * d0; v[1] += d0 * d1; v[2] += d0 * d2;
    v[3] += d1 * d1; v[4] += d1 * d2; v[5] += d2 * d2;
  }
  grid_reduce_sum<6>(v, smem, partials, ticket, out);
}

// min/max with first-index tie-break of t = axis.(p - origin)/|axis|^2
struct RangeAcc {
  double tmin, tmax;
  long long imin, imax;
};
__device__ __forceinline__ void range_merge(RangeAcc& a, double tmin, long long imin, double tmax,
                                            long long imax) {
  if (imin >= 0 && (a.imin < 0 || tmin < a.tmin || (tmin == a.tmin && imin < a.imin))) {
    a.tmin = tmin;
    a.imin = imin;
  }
  if (imax >= 0 && (a.imax < 0 || tmax > a.tmax || (tmax == a.tmax && imax < a.imax))) {
    a.tmax = tmax;
    a.imax = imax;
  }
}

__global__ void __launch_bounds__(kPsThreads)
    k_ps_range(const double* __restrict__ pts, int64_t n, Sel sel, double o0, double o1, double o2,
               double a0, double a1, double a2, double inv_nn_unused, double nn,
               RangeAcc* partials, unsigned* ticket, RangeAcc* out) {
  __shared__ RangeAcc sm[32];
  __shared__ bool is_last;
  RangeAcc r{0.0, 0.0, -1, -1};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    if (!sel(i)) continue;
    double p[3];
    load3(pts, i, p);
    // FMA-free in the reference's order: the arg-extrema are index decisions
    double t = __ddiv_rn(__dadd_rn(__dadd_rn(__dmul_rn(a0, __dsub_rn(p[0], o0)),
                                             __dmul_rn(a1, __dsub_rn(p[1], o1))),
                                   __dmul_rn(a2, __dsub_rn(p[2], o2))),
                         nn);
    range_merge(r, t, i, t, i);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double tmin = __shfl_xor_sync(0xffffffffu, r.tmin, o), tmax = __shfl_xor_sync(0xffffffffu, r.tmax, o);
    long long imin = __shfl_xor_sync(0xffffffffu, r.imin, o), imax = __shfl_xor_sync(0xffffffffu, r.imax, o);
    range_merge(r, tmin, imin, tmax, imax);
  }
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  if (lane == 0) sm[warp] = r;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < nw; ++w) range_merge(r, sm[w].tmin, sm[w].imin, sm[w].tmax, sm[w].imax);
    partials[blockIdx.x] = r;
    __threadfence();
    is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    __threadfence();
    RangeAcc t{0.0, 0.0, -1, -1};
    for (unsigned b = 0; b < gridDim.x; ++b) {
      RangeAcc q = partials[b];
      range_merge(t, q.tmin, q.imin, q.tmax, q.imax);
    }
    *out = t;
    *ticket = 0u;
  }
}

// p <- s R p + t, n <- R n (forward)  |  p <- (1/s) R^T (p - t), n <- R^T n (inverse)
__global__ void k_ps_apply(double* __restrict__ pts, double* __restrict__ nrm, int64_t n, Sel sel,
                           const double* __restrict__ srt /* s, R[9], t[3] */, int inverse) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n || !sel(i)) return;
  double s = srt[0];
  const double* R = srt + 1;
  const double* t = srt + 10;
  double p[3] = {pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
  if (!inverse) {
#pragma unroll
    for (int a = 0; a < 3; ++a)
      pts[3 * i + a] = s * (R[3 * a] * p[0] + R[3 * a + 1] * p[1] + R[3 * a + 2] * p[2]) + t[a];
  } else {
    double d[3] = {p[0] - t[0], p[1] - t[1], p[2] - t[2]};
#pragma unroll
    for (int a = 0; a < 3; ++a) pts[3 * i + a] = (R[a] * d[0] + R[3 + a] * d[1] + R[6 + a] * d[2]) / s;
  }
  if (nrm) {
    double q[3] = {nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]};
#pragma unroll
    for (int a = 0; a < 3; ++a)
      nrm[3 * i + a] = inverse ? (R[a] * q[0] + R[3 + a] * q[1] + R[6 + a] * q[2])
                               : (R[3 * a] * q[0] + R[3 * a + 1] * q[1] + R[3 * a + 2] * q[2]);
  }
}

__global__ void k_ps_translate(double* __restrict__ pts, int64_t n, double o0, double o1,
                               double o2) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  pts[3 * i] += o0;
  pts[3 * i + 1] += o1;
  pts[3 * i + 2] += o2;
}

__global__ void k_apply_f4(const float4* __restrict__ in, float4* __restrict__ out, int64_t n,
                           const double* __restrict__ srt) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  double s = srt[0];
  const double* R = srt + 1;
  const double* t = srt + 10;
  for (; i < n; i += stride) {
    float4 p = __ldg(in + i);
    double x = p.x, y = p.y, z = p.z;
    out[i] = make_float4((float)(s * (R[0] * x + R[1] * y + R[2] * z) + t[0]),
                         (float)(s * (R[3] * x + R[4] * y + R[5] * z) + t[1]),
                         (float)(s * (R[6] * x + R[7] * y + R[8] * z) + t[2]), p.w);
  }
}

__global__ void k_f64xyz_to_f4(const double* __restrict__ in, float4* __restrict__ out, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = make_float4((float)in[3 * i], (float)in[3 * i + 1], (float)in[3 * i + 2], 0.f);
}

__global__ void k_gather_labels(const int* __restrict__ nn, const int* __restrict__ src_labels,
                                int64_t n, int* __restrict__ out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = src_labels[nn[i]];
}

__global__ void k_label_presence(const int* __restrict__ labels, int64_t n, unsigned mask,
                                 unsigned* __restrict__ out) {
  unsigned m = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    int l = labels[i];
    if (l >= 0 && l <= 31 && ((mask >> l) & 1u)) m |= 1u << l;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m |= __shfl_xor_sync(0xffffffffu, m, o);
  if ((threadIdx.x & 31) == 0 && m) atomicOr(out, m);
}

// RemoveGround pass B (Alignment.cpp:103-126,148-153): per side count, sum c, sum c c^T
__global__ void __launch_bounds__(kPsThreads)
    k_ground_sides(const double* __restrict__ pts, int64_t n, double b0, double b1, double b2,
                   double a0, double a1, double a2, double nn, double thr1, double thr2,
                   double* partials, unsigned* ticket, double* out) {
  __shared__ double smem[32 * 20];
  double v[20];
#pragma unroll
  for (int k = 0; k < 20; ++k) v[k] = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    double p[3];
    load3(pts, i, p);
    double t = __ddiv_rn(__dadd_rn(__dadd_rn(__dmul_rn(a0, __dsub_rn(p[0], b0)),
                                             __dmul_rn(a1, __dsub_rn(p[1], b1))),
                                   __dmul_rn(a2, __dsub_rn(p[2], b2))),
                         nn);
    int side = -1;
    if (t < 0) {
      if (-t > thr1) side = 0;
    } else if (t > thr2)
      side = 1;
    if (side < 0) continue;
    double* w = v + 10 * side;
    w[0] += 1.0;
    w[1] += p[0]; w[2] += p[1]; w[3] += p[2];
    w[4] += p[0] * p[0]; w[5] += p[0] * p[1]; w[6] += p[0] * p[2];
    w[7] += p[1] * p[1]; w[8] += p[1] * p[2]; w[9] += p[2] * p[2];
  }
  grid_reduce_sum<20>(v, smem, partials, ticket, out);
}

// RemoveGround pass C/D (Alignment.cpp:178-193): max plane distance over the candidates of the
// chosen side; then the remove flags.
struct GroundSel {
  double b[3], a[3], nn, thr;
  int side;  // 0: t < 0 branch, 1: t >= 0 branch
  __device__ __forceinline__ bool candidate(const double p[3]) const {
    double t = __ddiv_rn(__dadd_rn(__dadd_rn(__dmul_rn(a[0], __dsub_rn(p[0], b[0])),
                                             __dmul_rn(a[1], __dsub_rn(p[1], b[1]))),
                                   __dmul_rn(a[2], __dsub_rn(p[2], b[2]))),
                         nn);
    return side == 0 ? (t < 0 && -t > thr) : (!(t < 0) && t > thr);
  }
};
__device__ __forceinline__ double plane_dist(const double p[3], const double pl[4]) {
  return fabs(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(pl[0], p[0]), __dmul_rn(pl[1], p[1])),
                                  __dmul_rn(pl[2], p[2])),
                        pl[3]));
}

__global__ void __launch_bounds__(kPsThreads)
    k_ground_maxdist(const double* __restrict__ pts, int64_t n, GroundSel gs, double p0, double p1,
                     double p2, double p3, double* partials, unsigned* ticket, double* out) {
  __shared__ double smem[32];
  double pl[4] = {p0, p1, p2, p3};
  double v[1] = {-INFINITY};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    double p[3];
    load3(pts, i, p);
    if (gs.candidate(p)) v[0] = fmax(v[0], plane_dist(p, pl));
  }
  grid_reduce_max<1>(v, smem, partials, ticket, out);
}

__global__ void k_ground_flags(const double* __restrict__ pts, int64_t n, GroundSel gs, double p0,
                               double p1, double p2, double p3, double threshold,
                               unsigned char* __restrict__ remove) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double pl[4] = {p0, p1, p2, p3};
  double p[3];
  load3(pts, i, p);
  remove[i] = (gs.candidate(p) && plane_dist(p, pl) < threshold) ? 1 : 0;
}

// ------------------------------------------------------------------ host-side drivers
struct PsWork {
  mvs_ctx* ctx;
  DevBuf<double> partials, out;
  DevBuf<RangeAcc> rpart, rout;
  DevBuf<unsigned> ticket;
  DevBuf<double> srt;
  int grid = 1;
  int init(mvs_ctx* c, int64_t nmax) {
    ctx = c;
    grid = grid_for(c, nmax > 0 ? nmax : 1, kPsThreads, 4);
    MVS_TRY(partials.reserve((size_t)grid * 20 + 32));
    MVS_TRY(out.reserve(32));
    MVS_TRY(rpart.reserve((size_t)grid));
    MVS_TRY(rout.reserve(1));
    MVS_TRY(ticket.reserve(1));
    MVS_TRY(srt.reserve(16));
    MVS_CUDA(cudaMemsetAsync(ticket.p, 0, sizeof(unsigned), c->stream));
    return MVS_OK;
  }
  int grid_n(int64_t n) const { return grid_for(ctx, n > 0 ? n : 1, kPsThreads, 4); }
};

struct HostStats {
  double bary[3] = {0, 0, 0}, bmin[3] = {0, 0, 0}, bmax[3] = {0, 0, 0};
  double pivots[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};  // row-major, columns = axes
  double evals[3] = {0, 0, 0};
  int64_t count = 0;
};

static int ps_stats(PsWork& w, const double* pts, int64_t n, Sel sel, bool want_bbox,
                    HostStats* hs) {
  mvs_ctx* ctx = w.ctx;
  cudaStream_t st = ctx->stream;
  double h[16];
  int grid = w.grid_n(n);
  {
    ProfScope ps(ctx, "ps_sum");
    k_ps_sum<<<grid, kPsThreads, 0, st>>>(pts, n, sel, w.partials.p, w.ticket.p, w.out.p);
  }
  if (want_bbox) {
    ProfScope ps(ctx, "ps_bbox");
    k_ps_bbox<<<grid, kPsThreads, 0, st>>>(pts, n, sel, w.partials.p, w.ticket.p, w.out.p + 4);
  }
  MVS_CUDA(cudaGetLastError());
  MVS_TRY(d2h_sync(ctx, h, w.out.p, sizeof(double) * 10));
  hs->count = (int64_t)h[3];
  if (hs->count <= 0) return MVS_OK;
  for (int a = 0; a < 3; ++a) {
    hs->bary[a] = h[a] / (double)hs->count;
    if (want_bbox) {
      hs->bmin[a] = -h[4 + a];
      hs->bmax[a] = h[7 + a];
    }
  }
  {
    ProfScope ps(ctx, "ps_cov");
    k_ps_cov<<<grid, kPsThreads, 0, st>>>(pts, n, sel, hs->bary[0], hs->bary[1], hs->bary[2],
                                          w.partials.p, w.ticket.p, w.out.p);
  }
  MVS_CUDA(cudaGetLastError());
  MVS_TRY(d2h_sync(ctx, h, w.out.p, sizeof(double) * 6));
  double denom = hs->count > 1 ? (double)(hs->count - 1) : 1.0;  // PointSetUtils.cpp:26
  double C[9] = {h[0], h[1], h[2], h[1], h[3], h[4], h[2], h[4], h[5]};
  for (int k = 0; k < 9; ++k) C[k] /= denom;
  double ev[3], Q[9];
  sym_eig3(C, ev, Q);  // ascending
  for (int c = 0; c < 3; ++c) {
    int s = 2 - c;  // PointSetUtils.cpp:35: descending eigenvalue order
    double v[3] = {Q[s], Q[3 + s], Q[6 + s]};
    double l = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    int big = 0;
    for (int a = 1; a < 3; ++a)
      if (fabs(v[a]) > fabs(v[big])) big = a;
    double sg = v[big] < 0 ? -1.0 : 1.0;
    for (int a = 0; a < 3; ++a) hs->pivots[3 * a + c] = sg * v[a] / l;
    hs->evals[c] = ev[s];
  }
  return MVS_OK;
}

struct HostRange {
  double tmin = std::numeric_limits<double>::max();
  double tmax = std::numeric_limits<double>::lowest();
  int64_t imin = -1, imax = -1;
};

static int ps_range(PsWork& w, const double* pts, int64_t n, Sel sel, const double origin[3],
                    const double axis[3], HostRange* hr) {
  mvs_ctx* ctx = w.ctx;
  double l = sqrt(axis[0] * axis[0] + axis[1] * axis[1] + axis[2] * axis[2]);
  double nn = l * l;  // pivot.norm() * pivot.norm()
  {
    ProfScope ps(ctx, "ps_range");
    k_ps_range<<<w.grid_n(n), kPsThreads, 0, ctx->stream>>>(pts, n, sel, origin[0], origin[1],
                                                            origin[2], axis[0], axis[1], axis[2],
                                                            0.0, nn, w.rpart.p, w.ticket.p,
                                                            w.rout.p);
  }
  MVS_CUDA(cudaGetLastError());
  RangeAcc r;
  MVS_TRY(d2h_sync(ctx, &r, w.rout.p, sizeof(r)));
  *hr = HostRange();
  if (r.imin >= 0) {
    hr->tmin = r.tmin;
    hr->imin = r.imin;
  }
  if (r.imax >= 0) {
    hr->tmax = r.tmax;
    hr->imax = r.imax;
  }
  return MVS_OK;
}

static int ps_apply(PsWork& w, double* pts, double* nrm, int64_t n, Sel sel, double s,
                    const double R[9], const double t[3], int inverse) {
  mvs_ctx* ctx = w.ctx;
  double h[13];
  h[0] = s;
  memcpy(h + 1, R, sizeof(double) * 9);
  memcpy(h + 10, t, sizeof(double) * 3);
  // stream-ordered: the previous user of w.srt has been launched already
  MVS_CUDA(cudaMemcpyAsync(w.srt.p, h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream));
  MVS_CUDA(cudaStreamSynchronize(ctx->stream));  // h is on the stack
  ProfScope ps(ctx, "ps_apply");
  k_ps_apply<<<div_up(n, 256), 256, 0, ctx->stream>>>(pts, nrm, n, sel, w.srt.p, inverse);
  MVS_CUDA(cudaGetLastError());
  return MVS_OK;
}

static const Sel kAll{nullptr, 0, -1};

// host union-find, Alignment.cpp:618-654 + SetUtils/UnionSetUtils.cpp:4-45 (rank == size)
static void retain_connected_host(double* pts, double* normals, int64_t* n_io, int32_t* facets,
                                  int64_t* nf_io) {
  int64_t n = *n_io, nf = *nf_io;
  if (n <= 0) return;
  std::vector<int32_t> par((size_t)n), rank((size_t)n, 1);
  std::iota(par.begin(), par.end(), 0);
  auto find = [&](int32_t x) {
    int32_t r = x;
    while (par[r] != r) r = par[r];
    while (x != r) {
      int32_t nx = par[x];
      par[x] = r;
      x = nx;
    }
    return r;
  };
  for (int64_t f = 0; f < nf; ++f)
    for (int e = 1; e <= 2; ++e) {
      int32_t a = find(facets[3 * f]), b = find(facets[3 * f + e]);
      if (a == b) continue;
      if (rank[a] > rank[b]) {
        par[b] = a;
        rank[a] += rank[b];
      } else {
        par[a] = b;
        rank[b] += rank[a];
      }
    }
  int32_t root = 0;
  for (int64_t i = 1; i < n; ++i)
    if (rank[root] < rank[i]) root = (int32_t)i;  // ProminentRepresent: first maximum
  std::vector<int32_t> remap((size_t)n, -1);
  int64_t m = 0;
  for (int64_t i = 0; i < n; ++i) {
    if (find((int32_t)i) != root) continue;
    if (m != i) {
      memcpy(pts + 3 * m, pts + 3 * i, sizeof(double) * 3);
      if (normals) memcpy(normals + 3 * m, normals + 3 * i, sizeof(double) * 3);
    }
    remap[i] = (int32_t)m++;
  }
  int64_t mf = 0;
  for (int64_t f = 0; f < nf; ++f) {
    if (remap[facets[3 * f]] < 0) continue;
    for (int e = 0; e < 3; ++e) facets[3 * mf + e] = remap[facets[3 * f + e]];
    ++mf;
  }
  *n_io = m;
  *nf_io = mf;
}

// Alignment.cpp:79-233 on a device copy `d_pts` of the host points; compaction + connected
// component on the host arrays; returns the compacted set re-uploaded into d_pts.
static int remove_ground_impl(PsWork& w, DevBuf<double>& d_pts, double* pts, double* normals,
                              int64_t* n_io, int32_t* facets, int64_t* nf_io, double dist_thres,
                              double ground_ratio, double ground_ray[3]) {
  mvs_ctx* ctx = w.ctx;
  cudaStream_t st = ctx->stream;
  int64_t n = *n_io, nf = *nf_io;
  HostStats hs;
  MVS_TRY(ps_stats(w, d_pts.p, n, kAll, false, &hs));
  double pivot[3] = {hs.pivots[0], hs.pivots[3], hs.pivots[6]};
  HostRange hr;
  MVS_TRY(ps_range(w, d_pts.p, n, kAll, hs.bary, pivot, &hr));
  const double dmin = std::numeric_limits<double>::min();
  double tMax1 = hr.tmin < 0 ? std::max(-hr.tmin, dmin) : dmin;
  double tMax2 = hr.tmax >= 0 ? std::max(hr.tmax, dmin) : dmin;
  double l = sqrt(pivot[0] * pivot[0] + pivot[1] * pivot[1] + pivot[2] * pivot[2]);
  double nn = l * l;
  double h[20];
  {
    ProfScope ps(ctx, "ground_sides");
    k_ground_sides<<<w.grid_n(n), kPsThreads, 0, st>>>(d_pts.p, n, hs.bary[0], hs.bary[1],
                                                       hs.bary[2], pivot[0], pivot[1], pivot[2], nn,
                                                       tMax1 * dist_thres, tMax2 * dist_thres,
                                                       w.partials.p, w.ticket.p, w.out.p);
  }
  MVS_CUDA(cudaGetLastError());
  MVS_TRY(d2h_sync(ctx, h, w.out.p, sizeof(h)));
  int side = h[0] > h[10] ? 0 : 1;  // candidates1.size() > candidates2.size() (:129)
  for (int a = 0; a < 3; ++a) ground_ray[a] = side == 0 ? -pivot[a] : pivot[a];
  const double* s = h + 10 * side;
  double A[9] = {s[4], s[5], s[6], s[5], s[7], s[8], s[6], s[8], s[9]}, Ai[9];
  mat3_inverse(A, Ai);
  double ans[3];
  for (int a = 0; a < 3; ++a) ans[a] = -(Ai[3 * a] * s[1] + Ai[3 * a + 1] * s[2] + Ai[3 * a + 2] * s[3]);
  double la = sqrt(ans[0] * ans[0] + ans[1] * ans[1] + ans[2] * ans[2]);
  double d = 1.0 / la;
  for (int a = 0; a < 3; ++a) ans[a] /= la;
  if (ans[0] * pivot[0] + ans[1] * pivot[1] + ans[2] * pivot[2] < 0) {
    for (int a = 0; a < 3; ++a) ans[a] = -ans[a];
    d = -d;
  }
  GroundSel gs;
  for (int a = 0; a < 3; ++a) {
    gs.b[a] = hs.bary[a];
    gs.a[a] = pivot[a];
  }
  gs.nn = nn;
  gs.side = side;
  gs.thr = (side == 0 ? tMax1 : tMax2) * dist_thres;
  {
    ProfScope ps(ctx, "ground_maxdist");
    k_ground_maxdist<<<w.grid_n(n), kPsThreads, 0, st>>>(d_pts.p, n, gs, ans[0], ans[1], ans[2], d,
                                                         w.partials.p, w.ticket.p, w.out.p);
  }
  MVS_CUDA(cudaGetLastError());
  double maxDist = 0;
  MVS_TRY(d2h_sync(ctx, &maxDist, w.out.p, sizeof(double)));
  maxDist = std::max(maxDist, dmin);  // :181 initial value
  double threshold = maxDist * ground_ratio;
  DevBuf<unsigned char> flags;
  MVS_TRY(flags.reserve((size_t)n));
  {
    ProfScope ps(ctx, "ground_flags");
    k_ground_flags<<<div_up(n, 256), 256, 0, st>>>(d_pts.p, n, gs, ans[0], ans[1], ans[2], d,
                                                   threshold, flags.p);
  }
  MVS_CUDA(cudaGetLastError());
  std::vector<unsigned char> rem((size_t)n);
  MVS_TRY(d2h_sync(ctx, rem.data(), flags.p, (size_t)n));
  // compaction (:196-219) + largest connected component (:227) on the caller's host arrays
  std::vector<int32_t> remap((size_t)n, -1);
  int64_t m = 0;
  for (int64_t i = 0; i < n; ++i) {
    if (rem[i]) continue;
    if (m != i) {
      memcpy(pts + 3 * m, pts + 3 * i, sizeof(double) * 3);
      if (normals) memcpy(normals + 3 * m, normals + 3 * i, sizeof(double) * 3);
    }
    remap[i] = (int32_t)m++;
  }
  int64_t mf = 0;
  for (int64_t f = 0; f < nf; ++f) {
    int32_t a = remap[facets[3 * f]], b = remap[facets[3 * f + 1]], c = remap[facets[3 * f + 2]];
    if (a < 0 || b < 0 || c < 0) continue;
    facets[3 * mf] = a;
    facets[3 * mf + 1] = b;
    facets[3 * mf + 2] = c;
    ++mf;
  }
  *n_io = m;
  *nf_io = mf;
  retain_connected_host(pts, normals, n_io, facets, nf_io);
  if (*n_io > 0) MVS_TRY(h2d(ctx, d_pts.p, pts, sizeof(double) * 3 * (*n_io)));
  MVS_CUDA(cudaStreamSynchronize(st));
  return MVS_OK;
}

// Alignment.cpp:235-314
static int init_alignment_impl(PsWork& w, const double* d_src, int64_t ns, const double* d_tgt,
                               int64_t nt, const double ground_ray[3], const double view_ray[3],
                               double R[9], double tr[3], double* scale) {
  HostStats s1, s2;
  MVS_TRY(ps_stats(w, d_src, ns, kAll, false, &s1));
  MVS_TRY(ps_stats(w, d_tgt, nt, kAll, false, &s2));
  double* tp = s2.pivots;
  auto flip_if = [&](int c, const double ray[3]) {
    if (ray[0] * tp[c] + ray[1] * tp[3 + c] + ray[2] * tp[6 + c] < 0)
      for (int a = 0; a < 3; ++a) tp[3 * a + c] = -tp[3 * a + c];
  };
  flip_if(0, ground_ray);
  flip_if(2, view_ray);
  double ax1[3] = {s1.pivots[0], s1.pivots[3], s1.pivots[6]}, ax2[3] = {tp[0], tp[3], tp[6]};
  HostRange r1, r2;
  MVS_TRY(ps_range(w, d_src, ns, kAll, s1.bary, ax1, &r1));
  MVS_TRY(ps_range(w, d_tgt, nt, kAll, s2.bary, ax2, &r2));
  double sc = (r2.tmax - r2.tmin) / (r1.tmax - r1.tmin);
  double Pi[9];
  mat3_inverse(s1.pivots, Pi);
  mat3_mul(tp, Pi, R);
  double k = r2.tmax - r1.tmax * sc;
  for (int a = 0; a < 3; ++a) {
    double Rb = R[3 * a] * s1.bary[0] + R[3 * a + 1] * s1.bary[1] + R[3 * a + 2] * s1.bary[2];
    tr[a] = ax2[a] * k + s2.bary[a] - sc * Rb;
  }
  *scale = sc;
  return MVS_OK;
}

// Common/Utils.h:124-149 (clamped; parallel axes -> identity)
static void calc_rotation(const double before[3], const double after[3], double R[9]) {
  double lb = sqrt(before[0] * before[0] + before[1] * before[1] + before[2] * before[2]);
  double la = sqrt(after[0] * after[0] + after[1] * after[1] + after[2] * after[2]);
  double b[3] = {before[0] / lb, before[1] / lb, before[2] / lb};
  double a[3] = {after[0] / la, after[1] / la, after[2] / la};
  double value = std::max(-1.0, std::min(1.0, b[0] * a[0] + b[1] * a[1] + b[2] * a[2]));
  double angle = acos(value);
  double u[3] = {b[1] * a[2] - b[2] * a[1], b[2] * a[0] - b[0] * a[2], b[0] * a[1] - b[1] * a[0]};
  double lu = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
  for (int k = 0; k < 9; ++k) R[k] = (k % 4 == 0) ? 1.0 : 0.0;
  if (!(lu > 0.0)) return;
  for (int k = 0; k < 3; ++k) u[k] /= lu;
  double c = cos(angle), s = sin(angle);
  R[0] = c + u[0] * u[0] * (1 - c);
  R[1] = u[0] * u[1] * (1 - c) - u[2] * s;
  R[2] = u[1] * s + u[0] * u[2] * (1 - c);
  R[3] = u[2] * s + u[0] * u[1] * (1 - c);
  R[4] = c + u[1] * u[1] * (1 - c);
  R[5] = -u[0] * s + u[1] * u[2] * (1 - c);
  R[6] = -u[1] * s + u[0] * u[2] * (1 - c);
  R[7] = u[0] * s + u[1] * u[2] * (1 - c);
  R[8] = c + u[2] * u[2] * (1 - c);
}

static int label_presence(PsWork& w, const int* labels, int64_t n, unsigned mask, unsigned* out) {
  mvs_ctx* ctx = w.ctx;
  unsigned* d = reinterpret_cast<unsigned*>(w.rout.p);  // 4 spare bytes of scratch
  MVS_CUDA(cudaMemsetAsync(d, 0, sizeof(unsigned), ctx->stream));
  {
    ProfScope ps(ctx, "label_presence");
    k_label_presence<<<w.grid_n(n), kPsThreads, 0, ctx->stream>>>(labels, n, mask, d);
  }
  MVS_CUDA(cudaGetLastError());
  return d2h_sync(ctx, out, d, sizeof(unsigned));
}

// Alignment.cpp:423-546 for one limb; *ok = false leaves the limb untouched
static int local_core_impl(PsWork& w, const double* d_src, const int* d_sl, int64_t ns,
                           const int* h_sl, const double* d_tgt, const int* d_tl, int64_t nt,
                           const int* h_tl, unsigned sel_mask, int label, double* scale,
                           double R[9], double tr[3], bool* ok) {
  *ok = false;
  unsigned sp = 0, tp = 0;
  MVS_TRY(label_presence(w, d_sl, ns, sel_mask, &sp));
  MVS_TRY(label_presence(w, d_tl, nt, sel_mask, &tp));
  if (!sp || !tp) return MVS_OK;
  HostStats s1, s2;
  MVS_TRY(ps_stats(w, d_src, ns, Sel{d_sl, sel_mask, -1}, false, &s1));
  MVS_TRY(ps_stats(w, d_tgt, nt, Sel{d_tl, sel_mask, -1}, false, &s2));
  double ax1[3] = {s1.pivots[0], s1.pivots[3], s1.pivots[6]};
  double ax2[3] = {s2.pivots[0], s2.pivots[3], s2.pivots[6]};
  if (ax1[0] * ax2[0] + ax1[1] * ax2[1] + ax1[2] * ax2[2] < 0)
    for (int a = 0; a < 3; ++a) ax2[a] = -ax2[a];
  unsigned scnt = sp, tcnt = tp;
  int cs = __builtin_popcount(sp), ct = __builtin_popcount(tp);
  if (cs < ct) {
    unsigned miss = tcnt & ~scnt;
    if (miss) tcnt &= ~(1u << __builtin_ctz(miss));
  } else if (cs > ct) {
    unsigned miss = scnt & ~tcnt;
    if (miss) scnt &= ~(1u << __builtin_ctz(miss));
  }
  HostRange r1, r2;
  MVS_TRY(ps_range(w, d_src, ns, Sel{d_sl, scnt, -1}, s1.bary, ax1, &r1));
  MVS_TRY(ps_range(w, d_tgt, nt, Sel{d_tl, tcnt, -1}, s2.bary, ax2, &r2));
  if (r1.imin < 0 || r2.imin < 0) return MVS_OK;
  double a0 = r1.tmin, a1 = r1.tmax, b0 = r2.tmin, b1 = r2.tmax;
  int64_t f1 = r1.imin, n1 = r1.imax, n2 = r2.imax;
  if (h_sl[n1] != label) {
    std::swap(a0, a1);
    std::swap(f1, n1);
  }
  if (h_tl[n2] != label) std::swap(b0, b1);
  *scale = (b1 - b0) / (a1 - a0);
  calc_rotation(ax1, ax2, R);
  double X[3];
  MVS_TRY(d2h_sync(w.ctx, X, d_src + 3 * f1, sizeof(double) * 3));  // the far-end point stays fixed
  for (int a = 0; a < 3; ++a)
    tr[a] = X[a] - (*scale) * (R[3 * a] * X[0] + R[3 * a + 1] * X[1] + R[3 * a + 2] * X[2]);
  *ok = true;
  return MVS_OK;
}

// v = s R p + t in place on device doubles (n x 3): the kernel mvs_apply_srt runs, without the trip
// through host memory (mvs_stitch_pairs)
int apply_srt_dev_f64(mvs_ctx* ctx, double* xyz_dev, int64_t n, double scale, const double R[9],
                      const double t[3]) {
  if (n <= 0) return MVS_OK;
  PsWork w;
  MVS_TRY(w.init(ctx, n));
  return ps_apply(w, xyz_dev, nullptr, n, kAll, scale, R, t, 0);
}

static int upload_f64(mvs_ctx* ctx, const double* host, int64_t n3, DevBuf<double>& d) {
  MVS_TRY(d.reserve((size_t)(n3 > 0 ? n3 : 1)));
  if (n3 > 0) MVS_TRY(h2d(ctx, d.p, host, sizeof(double) * n3));
  return MVS_OK;
}

static int resolve_ctx(mvs_ctx*& ctx) {
  if (!ctx) ctx = default_ctx();
  if (!ctx) return MVS_ERR_NO_DEVICE;
  return ctx_ensure_device(ctx);
}

}  // namespace mvs

using namespace mvs;

extern "C" {

int mvs_align_default_params(mvs_align_params* p) {
  if (!p) return MVS_ERR_INVALID_ARG;
  p->dist_thres = 0.8;     // ParamParser::dist_thres default (Parameter/ParamParser.cpp)
  p->ground_ratio = 0.28;  // Alignment.cpp:187
  p->remove_ground = 1;
  p->local_align = 1;
  p->neck_label = 1;  // NECK, PartRecognition.h:15
  p->reserved = 0;
  return MVS_OK;
}

static int apply_common(mvs_ctx* ctx, double* pts, double* normals, int64_t n, double scale,
                        const double R[9], const double t[3], int inverse) {
  MVS_TRY(resolve_ctx(ctx));
  if (n < 0 || (n > 0 && !pts) || !R || !t || (inverse && scale == 0.0)) {
    set_error("apply_srt: bad argument");
    return MVS_ERR_INVALID_ARG;
  }
  if (n == 0) return MVS_OK;
  PsWork w;
  MVS_TRY(w.init(ctx, n));
  DevBuf<double> dp, dn;
  MVS_TRY(upload_f64(ctx, pts, 3 * n, dp));
  if (normals) MVS_TRY(upload_f64(ctx, normals, 3 * n, dn));
  MVS_TRY(ps_apply(w, dp.p, normals ? dn.p : nullptr, n, kAll, scale, R, t, inverse));
  if (normals)
    MVS_TRY(d2h_async(ctx, normals, dn.p, sizeof(double) * 3 * n));
  return d2h_sync(ctx, pts, dp.p, sizeof(double) * 3 * n);
}

int mvs_apply_srt(mvs_ctx* ctx, double* pts_xyz, double* normals_xyz, int64_t n, double scale,
                  const double R[9], const double t[3]) {
  return apply_common(ctx, pts_xyz, normals_xyz, n, scale, R, t, 0);
}

int mvs_apply_srt_inverse(mvs_ctx* ctx, double* pts_xyz, double* normals_xyz, int64_t n,
                          double scale, const double R[9], const double t[3]) {
  return apply_common(ctx, pts_xyz, normals_xyz, n, scale, R, t, 1);
}

int mvs_apply_srt_dev_f32(mvs_ctx* ctx, const float* in_xyzw_dev, float* out_xyzw_dev, int64_t n,
                          double scale, const double R[9], const double t[3]) {
  MVS_TRY(resolve_ctx(ctx));
  if (n < 0 || (n > 0 && (!in_xyzw_dev || !out_xyzw_dev)) || !R || !t) {
    set_error("apply_srt_dev_f32: bad argument");
    return MVS_ERR_INVALID_ARG;
  }
  if (n == 0) return MVS_OK;
  DevBuf<double> srt;
  MVS_TRY(srt.reserve(16));
  double h[13];
  h[0] = scale;
  memcpy(h + 1, R, sizeof(double) * 9);
  memcpy(h + 10, t, sizeof(double) * 3);
  MVS_CUDA(cudaMemcpyAsync(srt.p, h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream));
  {
    ProfScope ps(ctx, "apply_srt_f4");
    k_apply_f4<<<grid_for(ctx, n, 256, 8), 256, 0, ctx->stream>>>(
        reinterpret_cast<const float4*>(in_xyzw_dev), reinterpret_cast<float4*>(out_xyzw_dev), n,
        srt.p);
  }
  MVS_CUDA(cudaGetLastError());
  MVS_CUDA(cudaStreamSynchronize(ctx->stream));  // srt / h go out of scope
  return MVS_OK;
}

int mvs_pointset_stats(mvs_ctx* ctx, const double* pts_xyz, int64_t n, double barycenter[3],
                       double bbox_min[3], double bbox_max[3], double pivots[9],
                       double eigenvalues[3]) {
  MVS_TRY(resolve_ctx(ctx));
  if (n <= 0 || !pts_xyz) {
    set_error("pointset_stats: empty point set");
    return MVS_ERR_DEGENERATE;
  }
  PsWork w;
  MVS_TRY(w.init(ctx, n));
  DevBuf<double> dp;
  MVS_TRY(upload_f64(ctx, pts_xyz, 3 * n, dp));
  HostStats hs;
  MVS_TRY(ps_stats(w, dp.p, n, kAll, true, &hs));
  for (int a = 0; a < 3; ++a) {
    if (barycenter) barycenter[a] = hs.bary[a];
    if (bbox_min) bbox_min[a] = hs.bmin[a];
    if (bbox_max) bbox_max[a] = hs.bmax[a];
    if (eigenvalues) eigenvalues[a] = hs.evals[a];
  }
  if (pivots) memcpy(pivots, hs.pivots, sizeof(double) * 9);
  return MVS_OK;
}

int mvs_project_range(mvs_ctx* ctx, const double* pts_xyz, int64_t n, const int32_t* labels,
                      int32_t label, const double origin[3], const double axis[3], double* tmin,
                      double* tmax, int64_t* argmin, int64_t* argmax) {
  MVS_TRY(resolve_ctx(ctx));
  if (n <= 0 || !pts_xyz || !origin || !axis) {
    set_error("project_range: bad argument");
    return MVS_ERR_INVALID_ARG;
  }
  if (labels && label < 0) {
    set_error("project_range: label must be >= 0");
    return MVS_ERR_INVALID_ARG;
  }
  PsWork w;
  MVS_TRY(w.init(ctx, n));
  DevBuf<double> dp;
  DevBuf<int> dl;
  MVS_TRY(upload_f64(ctx, pts_xyz, 3 * n, dp));
  if (labels) {
    MVS_TRY(dl.reserve((size_t)n));
    MVS_TRY(h2d(ctx, dl.p, labels, sizeof(int) * n));
  }
  HostRange hr;
  MVS_TRY(ps_range(w, dp.p, n, Sel{labels ? dl.p : nullptr, 0, labels ? label : -1}, origin, axis,
                   &hr));
  if (tmin) *tmin = hr.tmin;
  if (tmax) *tmax = hr.tmax;
  if (argmin) *argmin = hr.imin;
  if (argmax) *argmax = hr.imax;
  return MVS_OK;
}

int mvs_retain_connected(double* pts_xyz, double* normals_xyz, int64_t* n, int32_t* facets,
                         int64_t* n_facets) {
  if (!pts_xyz || !n || !n_facets || (*n_facets > 0 && !facets) || *n < 0 || *n_facets < 0) {
    set_error("retain_connected: bad argument");
    return MVS_ERR_INVALID_ARG;
  }
  for (int64_t i = 0; i < 3 * (*n_facets); ++i)
    if (facets[i] < 0 || facets[i] >= *n) {
      set_error("retain_connected: facet index %d out of range", facets[i]);
      return MVS_ERR_INVALID_ARG;
    }
  retain_connected_host(pts_xyz, normals_xyz, n, facets, n_facets);
  return MVS_OK;
}

static int check_mesh(const double* pts, const int64_t* n, const int32_t* facets,
                      const int64_t* nf) {
  if (!pts || !n || *n <= 0) {
    set_error("empty target point set");
    return MVS_ERR_DEGENERATE;
  }
  if (!nf || *nf < 0 || (*nf > 0 && !facets)) {
    set_error("bad facet array");
    return MVS_ERR_INVALID_ARG;
  }
  for (int64_t i = 0; i < 3 * (*nf); ++i)
    if (facets[i] < 0 || facets[i] >= *n) {
      set_error("facet index %d out of range", facets[i]);
      return MVS_ERR_INVALID_ARG;
    }
  return MVS_OK;
}

int mvs_remove_ground(mvs_ctx* ctx, double* pts_xyz, double* normals_xyz, int64_t* n,
                      int32_t* facets, int64_t* n_facets, const mvs_align_params* params,
                      double ground_ray[3]) {
  MVS_TRY(resolve_ctx(ctx));
  MVS_TRY(check_mesh(pts_xyz, n, facets, n_facets));
  if (!ground_ray) {
    set_error("remove_ground: null ground_ray");
    return MVS_ERR_INVALID_ARG;
  }
  mvs_align_params def;
  if (!params) {
    mvs_align_default_params(&def);
    params = &def;
  }
  PsWork w;
  MVS_TRY(w.init(ctx, *n));
  DevBuf<double> dp;
  MVS_TRY(upload_f64(ctx, pts_xyz, 3 * (*n), dp));
  return remove_ground_impl(w, dp, pts_xyz, normals_xyz, n, facets, n_facets, params->dist_thres,
                            params->ground_ratio, ground_ray);
}

int mvs_align(mvs_ctx* ctx, double* src_xyz, double* src_normals, int64_t n_src,
              const int32_t* src_labels, double* tgt_xyz, double* tgt_normals, int64_t* n_tgt,
              int32_t* tgt_facets, int64_t* n_tgt_facets, const double view_ray[3],
              const mvs_align_params* params, int32_t* out_tgt_labels, double ground_ray[3]) {
  MVS_TRY(resolve_ctx(ctx));
  if (!src_xyz || n_src <= 0 || !src_labels || !view_ray || !ground_ray) {
    set_error("align: bad argument");
    return MVS_ERR_INVALID_ARG;
  }
  MVS_TRY(check_mesh(tgt_xyz, n_tgt, tgt_facets, n_tgt_facets));
  mvs_align_params def;
  if (!params) {
    mvs_align_default_params(&def);
    params = &def;
  }
  cudaStream_t st = ctx->stream;
  PsWork w;
  MVS_TRY(w.init(ctx, std::max<int64_t>(n_src, *n_tgt)));
  DevBuf<double> ds, dsn, dt;
  DevBuf<int> dsl, dtl, dnn;
  MVS_TRY(upload_f64(ctx, src_xyz, 3 * n_src, ds));
  if (src_normals) MVS_TRY(upload_f64(ctx, src_normals, 3 * n_src, dsn));
  MVS_TRY(upload_f64(ctx, tgt_xyz, 3 * (*n_tgt), dt));
  MVS_TRY(dsl.reserve((size_t)n_src));
  MVS_TRY(h2d(ctx, dsl.p, src_labels, sizeof(int) * n_src));
  // 1. RemoveGround (Alignment.cpp:21)
  if (params->remove_ground)
    MVS_TRY(remove_ground_impl(w, dt, tgt_xyz, tgt_normals, n_tgt, tgt_facets, n_tgt_facets,
                               params->dist_thres, params->ground_ratio, ground_ray));
  int64_t nt = *n_tgt;
  if (nt <= 0) {
    set_error("align: target empty after ground removal");
    return MVS_ERR_DEGENERATE;
  }
  // 2. InitAlignment + apply (Alignment.cpp:27-34)
  double R[9], tr[3], scale = 1.0;
  MVS_TRY(init_alignment_impl(w, ds.p, n_src, dt.p, nt, ground_ray, view_ray, R, tr, &scale));
  MVS_TRY(ps_apply(w, ds.p, src_normals ? dsn.p : nullptr, n_src, kAll, scale, R, tr, 0));
  // 3. PartRecog: exact 1-NN of every target point in the transformed template (Alignment.cpp:49)
  {
    DevBuf<float4> s4, t4;
    MVS_TRY(s4.reserve((size_t)n_src));
    MVS_TRY(t4.reserve((size_t)nt));
    MVS_TRY(dnn.reserve((size_t)nt));
    MVS_TRY(dtl.reserve((size_t)nt));
    {
      ProfScope ps(ctx, "f64_to_f4", 2);
      k_f64xyz_to_f4<<<div_up(n_src, 256), 256, 0, st>>>(ds.p, s4.p, n_src);
      k_f64xyz_to_f4<<<div_up(nt, 256), 256, 0, st>>>(dt.p, t4.p, nt);
    }
    mvs_nn_index* idx = nullptr;
    MVS_TRY(mvs_nn_index_create(ctx, &idx));
    int rc = nn_build(idx, s4.p, n_src);
    if (rc == MVS_OK) rc = nn_knn(idx, t4.p, nt, 1, dnn.p, nullptr);
    if (rc == MVS_OK) {
      ProfScope ps(ctx, "gather_labels");
      k_gather_labels<<<div_up(nt, 256), 256, 0, st>>>(dnn.p, dsl.p, nt, dtl.p);
    }
    cudaStreamSynchronize(st);
    mvs_nn_index_destroy(idx);
    MVS_TRY(rc);
    MVS_CUDA(cudaGetLastError());
  }
  std::vector<int32_t> h_tl((size_t)nt);
  MVS_TRY(d2h_sync(ctx, h_tl.data(), dtl.p, sizeof(int) * nt));
  // 4. neck-centroid translation (Alignment.cpp:54-64)
  {
    double a[4], b[4];
    int neck = params->neck_label;
    {
      ProfScope ps(ctx, "ps_sum", 2);
      k_ps_sum<<<w.grid_n(n_src), kPsThreads, 0, st>>>(ds.p, n_src, Sel{dsl.p, 0, neck},
                                                       w.partials.p, w.ticket.p, w.out.p);
      k_ps_sum<<<w.grid_n(nt), kPsThreads, 0, st>>>(dt.p, nt, Sel{dtl.p, 0, neck}, w.partials.p,
                                                    w.ticket.p, w.out.p + 4);
    }
    MVS_CUDA(cudaGetLastError());
    double h[8];
    MVS_TRY(d2h_sync(ctx, h, w.out.p, sizeof(h)));
    memcpy(a, h, sizeof(a));
    memcpy(b, h + 4, sizeof(b));
    if (neck >= 0 && a[3] > 0 && b[3] > 0) {
      ProfScope ps(ctx, "ps_translate");
      k_ps_translate<<<div_up(n_src, 256), 256, 0, st>>>(ds.p, n_src, b[0] / b[3] - a[0] / a[3],
                                                         b[1] / b[3] - a[1] / a[3],
                                                         b[2] / b[3] - a[2] / a[3]);
      MVS_CUDA(cudaGetLastError());
    }
  }
  // 5. LocalAlignment (Alignment.cpp:316-421): label sets per limb, PART enum values
  if (params->local_align) {
    struct Limb { unsigned sel; int label; unsigned apply; };
    const Limb limbs[4] = {
        {(1u << 2) | (1u << 3) | (1u << 4), 4, (1u << 2) | (1u << 3) | (1u << 4)},
        {(1u << 5) | (1u << 6) | (1u << 7), 7, (1u << 5) | (1u << 6) | (1u << 7)},
        {(1u << 8) | (1u << 9), 9, (1u << 8) | (1u << 9) | (1u << 10)},
        {(1u << 11) | (1u << 12), 12, (1u << 11) | (1u << 12) | (1u << 13)}};
    for (const Limb& L : limbs) {
      double Rl[9], tl[3], sc = 1.0;
      bool ok = false;
      MVS_TRY(local_core_impl(w, ds.p, dsl.p, n_src, src_labels, dt.p, dtl.p, nt, h_tl.data(),
                              L.sel, L.label, &sc, Rl, tl, &ok));
      if (!ok) continue;
      MVS_TRY(ps_apply(w, ds.p, src_normals ? dsn.p : nullptr, n_src, Sel{dsl.p, L.apply, -1}, sc,
                       Rl, tl, 0));
    }
  }
  if (out_tgt_labels) memcpy(out_tgt_labels, h_tl.data(), sizeof(int32_t) * nt);
  if (src_normals)
    MVS_TRY(d2h_async(ctx, src_normals, dsn.p, sizeof(double) * 3 * n_src));
  return d2h_sync(ctx, src_xyz, ds.p, sizeof(double) * 3 * n_src);
}

}  // extern "C"
