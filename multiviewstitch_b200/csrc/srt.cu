// srt.cu — scale-rotation-translation solver (Solver/SRTSolver.cpp) on the device.
//
//   closed form  = 2 streaming passes over the matches (centroids; ratio-of-norms scale + 3x3
//                  cross-covariance), 3x3 Jacobi SVD on one thread.           64 B/pair (float4)
//   RANSAC       = the same 2 passes, one thread per hypothesis for the 3-point Kabsch, then a
//                  (match tile x hypothesis chunk) scoring grid that replicates the reference's
//                  (int)(f*x/z + c + 0.5) pixel rounding in FMA-free double arithmetic.
// All sums are double precision with a fixed reduction tree (reduce.cuh).
#include <cmath>

#include "linalg3.h"
#include "nn_grid.cuh"
#include "reduce.cuh"

namespace mvs {

struct CamD {
  double fx, cx, fy, cy;
  double R[9];
  double t[3];
};

static CamD to_camd(const mvs_camera& c) {
  CamD d;
  d.fx = c.K[0];
  d.cx = c.K[2];
  d.fy = c.K[4];
  d.cy = c.K[5];
  for (int i = 0; i < 9; ++i) d.R[i] = c.R[i];
  for (int i = 0; i < 3; ++i) d.t[i] = c.t[i];
  return d;
}

// Match accessors: float4 AoS (C2 bandwidth layout) or six double planes (drop-in precision).
struct AccF4 {
  const float4* p;
  const float4* q;
  __device__ __forceinline__ void load(int64_t i, double P[3], double Q[3]) const {
    float4 a = __ldg(p + i), b = __ldg(q + i);
    P[0] = a.x; P[1] = a.y; P[2] = a.z;
    Q[0] = b.x; Q[1] = b.y; Q[2] = b.z;
  }
};
struct AccF64 {
  const double* base;  // px,py,pz,qx,qy,qz planes of length n
  int64_t n;
  __device__ __forceinline__ void load(int64_t i, double P[3], double Q[3]) const {
    P[0] = __ldg(base + i);
    P[1] = __ldg(base + n + i);
    P[2] = __ldg(base + 2 * n + i);
    Q[0] = __ldg(base + 3 * n + i);
    Q[1] = __ldg(base + 4 * n + i);
    Q[2] = __ldg(base + 5 * n + i);
  }
};

__global__ void k_pack_planes(const double* __restrict__ p, const double* __restrict__ q,
                              double* __restrict__ planes, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      planes[a * n + i] = p[3 * i + a];
      planes[(3 + a) * n + i] = q[3 * i + a];
    }
  }
}

constexpr int kSrtThreads = 256;

// device result block layout (doubles)
enum { R_CENT = 0, R_SUMS = 6, R_ACC = 12, R_SCALE = 22, R_ROT = 23, R_T = 32, R_BEST_ERR = 35,
       R_BEST_IT = 36, R_SIZE = 40 };

// pass 1: centroids (Solver/SRTSolver.cpp:34-41)
template <class Acc>
__global__ void __launch_bounds__(kSrtThreads)
    k_srt_centroids(Acc acc, int64_t n, double* partials, unsigned* ticket, double* res) {
  __shared__ double smem[32 * 6];
  double v[6] = {0, 0, 0, 0, 0, 0};
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    double P[3], Q[3];
    acc.load(i, P, Q);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      v[a] += P[a];
      v[3 + a] += Q[a];
    }
  }
  if (grid_reduce_sum<6>(v, smem, partials, ticket, res + R_SUMS)) {
    if (threadIdx.x < 6) res[R_CENT + threadIdx.x] = res[R_SUMS + threadIdx.x] / (double)n;
  }
}

// |b| / |a| from the squared norms.  Two IEEE float64 square roots and a division are ~90
// instructions of software sequences and made the pass compute bound (125 us per 1e7 pairs, 39 % of
// the HBM roofline); here 1/na2 and 1/sqrt(t) start from float32 hardware seeds and take two Newton
// steps each in float64 (relative error 6e-8 -> 4e-15 -> rounding level), ~25 instructions.
// Arguments outside the float32 seed range (zero, subnormal, huge, NaN) take the plain form, which
// also keeps the reference's behaviour for p_i = centroid (division by zero, SRTSolver.cpp:44).
__device__ __forceinline__ double norm_ratio(double nb2, double na2) {
  if (!(na2 > 1e-30 && na2 < 1e30 && nb2 > 1e-30 && nb2 < 1e30)) return sqrt(nb2) / sqrt(na2);
  double y = (double)__frcp_rn((float)na2);
  y = y * (2.0 - na2 * y);
  y = y * (2.0 - na2 * y);
  const double t = nb2 * y;
  double r = (double)rsqrtf((float)t);
  const double h = 0.5 * t;
  r = r * (1.5 - h * r * r);
  r = r * (1.5 - h * r * r);
  return t * r;
}

// pass 2: sum of |q-c2|/|p-c1| (SRTSolver.cpp:42-45) and raw cross-covariance (p-c1)(q-c2)^T
// (:75-107, the scale factor is applied once at the end); the last CTA finishes the closed form.
template <class Acc>
__global__ void __launch_bounds__(kSrtThreads, 4)
    k_srt_scale_cov(Acc acc, int64_t n, double* partials, unsigned* ticket, double* res) {
  __shared__ double smem[32 * 10];
  double c1[3], c2[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    c1[a] = res[R_CENT + a];
    c2[a] = res[R_CENT + 3 + a];
  }
  double v[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  auto add = [&](const double (&P)[3], const double (&Q)[3]) {
    double a0 = P[0] - c1[0], a1 = P[1] - c1[1], a2 = P[2] - c1[2];
    double b0 = Q[0] - c2[0], b1 = Q[1] - c2[1], b2 = Q[2] - c2[2];
    v[0] += norm_ratio(b0 * b0 + b1 * b1 + b2 * b2, a0 * a0 + a1 * a1 + a2 * a2);
    v[1] += a0 * b0; v[2] += a0 * b1; v[3] += a0 * b2;
    v[4] += a1 * b0; v[5] += a1 * b1; v[6] += a1 * b2;
    v[7] += a2 * b0; v[8] += a2 * b1; v[9] += a2 * b2;
  };
  // two pairs per trip, both loads issued before the float64 work of either
  for (; i + stride < n; i += 2 * stride) {
    double P0[3], Q0[3], P1[3], Q1[3];
    acc.load(i, P0, Q0);
    acc.load(i + stride, P1, Q1);
    add(P0, Q0);
    add(P1, Q1);
  }
  if (i < n) {
    double P[3], Q[3];
    acc.load(i, P, Q);
    add(P, Q);
  }
  grid_reduce_sum<10>(v, smem, partials, ticket, res + R_ACC);
}

// The closed form (3x3 Jacobi SVD in float64) is its own one-thread launch: inlined into the
// streaming pass it raised that kernel to 128 registers + 168 B of spill, i.e. 2 CTAs per SM instead
// of 4 and half the loads in flight (125 us per 1e7 pairs; 39 % of the HBM roofline).
__global__ void k_srt_finish(int64_t n, double* res) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double c1[3], c2[3];
  for (int a = 0; a < 3; ++a) {
    c1[a] = res[R_CENT + a];
    c2[a] = res[R_CENT + 3 + a];
  }
  double s = res[R_ACC] / (double)n;
  double S[9], R[9], t[3];
  for (int k = 0; k < 9; ++k) S[k] = s * res[R_ACC + 1 + k];
  kabsch_from_cov(S, s, c1, c2, R, t);
  res[R_SCALE] = s;
  for (int k = 0; k < 9; ++k) res[R_ROT + k] = R[k];
  for (int k = 0; k < 3; ++k) res[R_T + k] = t[k];
}

// hypothesis record: sR[9], t[3], iR[9] (+3 pad) ; plain R kept separately
constexpr int kHypStride = 24;

__host__ __device__ inline void make_hyp_record(double s, const double* R, const double* t,
                                                double* rec) {
  double inv_s = 1.0 / s;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      rec[3 * i + j] = s * R[3 * i + j];
      rec[12 + 3 * i + j] = inv_s * R[3 * j + i];
    }
  for (int i = 0; i < 3; ++i) rec[9 + i] = t[i];
  rec[21] = rec[22] = rec[23] = 0.0;
}

// one thread per RANSAC hypothesis (SRTSolver.cpp:148-174): 3 globally-centred pairs -> S -> Kabsch
template <class Acc>
__global__ void k_ransac_hyp(Acc acc, const int* __restrict__ triples, int iters,
                             const double* __restrict__ res, double* __restrict__ hyps,
                             double* __restrict__ rts) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= iters) return;
  double c1[3], c2[3];
  for (int a = 0; a < 3; ++a) {
    c1[a] = res[R_CENT + a];
    c2[a] = res[R_CENT + 3 + a];
  }
  double s = res[R_SCALE];
  double S[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int j = 0; j < 3; ++j) {
    double P[3], Q[3], x[3], y[3];
    acc.load(triples[3 * k + j], P, Q);
    for (int a = 0; a < 3; ++a) {
      x[a] = (P[a] - c1[a]) * s;
      y[a] = Q[a] - c2[a];
    }
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) S[3 * a + b] += x[a] * y[b];
  }
  double R[9], t[3];
  kabsch_from_cov(S, s, c1, c2, R, t);
  make_hyp_record(s, R, t, hyps + (size_t)k * kHypStride);
  for (int a = 0; a < 9; ++a) rts[12 * k + a] = R[a];
  for (int a = 0; a < 3; ++a) rts[12 * k + 9 + a] = t[a];
}

// Camera/Camera.cpp:68-72 then :45-48, in round-to-nearest double with no FMA so the truncated
// pixel coordinates are bit-identical to the host evaluation.
__device__ __forceinline__ void project_rn(const CamD& c, double x, double y, double z, int& u,
                                           int& v) {
  double pcx = __dadd_rn(
      __dadd_rn(__dadd_rn(__dmul_rn(c.R[0], x), __dmul_rn(c.R[1], y)), __dmul_rn(c.R[2], z)),
      c.t[0]);
  double pcy = __dadd_rn(
      __dadd_rn(__dadd_rn(__dmul_rn(c.R[3], x), __dmul_rn(c.R[4], y)), __dmul_rn(c.R[5], z)),
      c.t[1]);
  double pcz = __dadd_rn(
      __dadd_rn(__dadd_rn(__dmul_rn(c.R[6], x), __dmul_rn(c.R[7], y)), __dmul_rn(c.R[8], z)),
      c.t[2]);
  u = __double2int_rz(__dadd_rn(__dadd_rn(__ddiv_rn(__dmul_rn(c.fx, pcx), pcz), c.cx), 0.5));
  v = __double2int_rz(__dadd_rn(__dadd_rn(__ddiv_rn(__dmul_rn(c.fy, pcy), pcz), c.cy), 0.5));
}

__device__ __forceinline__ double pix_dist(int u1, int v1, int u2, int v2) {
  double du = __dsub_rn((double)u1, (double)u2), dv = __dsub_rn((double)v1, (double)v2);
  return __dsqrt_rn(__dadd_rn(__dmul_rn(du, du), __dmul_rn(dv, dv)));
}

struct CamPair {
  CamD c1, c2;
};

constexpr int kHypChunk = 8;

// Solver/SRTSolver.cpp:6-29 for a chunk of hypotheses at once.  grid = (match tiles, chunks).
template <class Acc>
__global__ void __launch_bounds__(kSrtThreads)
    k_residual(Acc acc, int64_t n, CamPair cams, const double* __restrict__ hyps, int nhyp,
               double* __restrict__ partials /* [nhyp][gridDim.x] */) {
  __shared__ double sh[kHypChunk * kHypStride];
  __shared__ double smem[32 * kHypChunk];
  int h0 = blockIdx.y * kHypChunk;
  int hc = min(kHypChunk, nhyp - h0);
  for (int i = threadIdx.x; i < kHypChunk * kHypStride; i += blockDim.x)
    sh[i] = (i < hc * kHypStride) ? hyps[(size_t)h0 * kHypStride + i] : 0.0;
  __syncthreads();
  double acc_e[kHypChunk];
#pragma unroll
  for (int j = 0; j < kHypChunk; ++j) acc_e[j] = 0.0;
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    double P[3], Q[3];
    acc.load(i, P, Q);
    int u2, v2, u1_, v1_;
    project_rn(cams.c2, Q[0], Q[1], Q[2], u2, v2);    // cam2 <- p2
    project_rn(cams.c1, P[0], P[1], P[2], u1_, v1_);  // cam1 <- p1
#pragma unroll
    for (int j = 0; j < kHypChunk; ++j) {
      if (j < hc) {
        const double* H = sh + j * kHypStride;
        double tp[3], tq[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          tp[a] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(H[3 * a], P[0]),
                                                __dmul_rn(H[3 * a + 1], P[1])),
                                      __dmul_rn(H[3 * a + 2], P[2])),
                            H[9 + a]);
        }
        double d0 = __dsub_rn(Q[0], H[9]), d1 = __dsub_rn(Q[1], H[10]), d2 = __dsub_rn(Q[2], H[11]);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          tq[a] = __dadd_rn(__dadd_rn(__dmul_rn(H[12 + 3 * a], d0), __dmul_rn(H[12 + 3 * a + 1], d1)),
                            __dmul_rn(H[12 + 3 * a + 2], d2));
        }
        int u1, v1, u2_, v2_;
        project_rn(cams.c2, tp[0], tp[1], tp[2], u1, v1);
        project_rn(cams.c1, tq[0], tq[1], tq[2], u2_, v2_);
        double e1 = pix_dist(u1, v1, u2, v2);
        double e2 = pix_dist(u1_, v1_, u2_, v2_);
        acc_e[j] += __dmul_rn(__dadd_rn(e1, e2), 0.5);
      }
    }
  }
  block_reduce_sum<kHypChunk>(acc_e, smem);
  if (threadIdx.x == 0) {
    for (int j = 0; j < hc; ++j) partials[(size_t)(h0 + j) * gridDim.x + blockIdx.x] = acc_e[j];
  }
}

// sum the per-tile partials in tile order, divide by n, strict-< argmin (first wins)
__global__ void k_residual_final(const double* __restrict__ partials, int tiles, int nhyp,
                                 int64_t n, double* __restrict__ errs,
                                 const double* __restrict__ rts, double* __restrict__ res) {
  for (int h = threadIdx.x; h < nhyp; h += blockDim.x) {
    double s = 0.0;
    for (int b = 0; b < tiles; ++b) s += partials[(size_t)h * tiles + b];
    errs[h] = s / (double)n;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double best = HUGE_VAL;
    int bi = -1;
    for (int h = 0; h < nhyp; ++h)
      if (errs[h] < best) {
        best = errs[h];
        bi = h;
      }
    res[R_BEST_ERR] = best;
    res[R_BEST_IT] = (double)bi;
    if (rts && bi >= 0) {
      for (int a = 0; a < 9; ++a) res[R_ROT + a] = rts[12 * bi + a];
      for (int a = 0; a < 3; ++a) res[R_T + a] = rts[12 * bi + 9 + a];
    }
  }
}

static inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

}  // namespace mvs

using namespace mvs;

struct mvs_srt {
  mvs_ctx* ctx = nullptr;
  int iter_num = 100;  // Solver/SRTSolver.h:10
  bool print = true;   // Solver/SRTSolver.h:10
  int64_t n = 0;
  int mode = 0;  // 0 none, 1 double planes, 2 float4
  DevBuf<double> planes;
  DevBuf<float4> pf, qf;
  const float4* p4 = nullptr;
  const float4* q4 = nullptr;
  bool have_cams = false;
  CamPair cams;
  DevBuf<double> partials, res, hyps, rts, errs;
  DevBuf<unsigned> ticket;
  DevBuf<int> triples;
};

// Processor/Processor.cpp:210-239: per-match symmetric reprojection test under one transform
template <class Acc>
__global__ void __launch_bounds__(kSrtThreads)
    k_srt_inliers(Acc acc, int64_t n, CamPair cams, const double* __restrict__ hyp, double thr,
                  unsigned char* __restrict__ flags, double* partials, unsigned* ticket,
                  double* out) {
  __shared__ double smem[32 * 2];
  __shared__ double H[kHypStride];
  if (threadIdx.x < kHypStride) H[threadIdx.x] = hyp[threadIdx.x];
  __syncthreads();
  double v[2] = {0.0, 0.0};
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    double P[3], Q[3];
    acc.load(i, P, Q);
    int u2, v2, u1_, v1_, u1, v1, u2_, v2_;
    project_rn(cams.c2, Q[0], Q[1], Q[2], u2, v2);
    project_rn(cams.c1, P[0], P[1], P[2], u1_, v1_);
    double tp[3], tq[3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
      tp[a] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(H[3 * a], P[0]), __dmul_rn(H[3 * a + 1], P[1])),
                                  __dmul_rn(H[3 * a + 2], P[2])),
                        H[9 + a]);
    double d0 = __dsub_rn(Q[0], H[9]), d1 = __dsub_rn(Q[1], H[10]), d2 = __dsub_rn(Q[2], H[11]);
#pragma unroll
    for (int a = 0; a < 3; ++a)
      tq[a] = __dadd_rn(__dadd_rn(__dmul_rn(H[12 + 3 * a], d0), __dmul_rn(H[12 + 3 * a + 1], d1)),
                        __dmul_rn(H[12 + 3 * a + 2], d2));
    project_rn(cams.c2, tp[0], tp[1], tp[2], u1, v1);
    project_rn(cams.c1, tq[0], tq[1], tq[2], u2_, v2_);
    double e1 = pix_dist(u1, v1, u2, v2);
    double e2 = pix_dist(u1_, v1_, u2_, v2_);
    bool in = (e1 <= thr) && (e2 <= thr);
    flags[i] = in ? 1 : 0;
    v[0] += __dmul_rn(__dadd_rn(e1, e2), 0.5);
    v[1] += in ? 1.0 : 0.0;
  }
  grid_reduce_sum<2>(v, smem, partials, ticket, out);
}

namespace {

int srt_prepare(mvs_srt* h) {
  MVS_TRY(h->res.reserve(R_SIZE));
  if (!h->ticket.p) {
    MVS_TRY(h->ticket.reserve(1));
    MVS_CUDA(cudaMemsetAsync(h->ticket.p, 0, sizeof(unsigned), h->ctx->stream));
  }
  return MVS_OK;
}

int srt_set_cams(mvs_srt* h, const mvs_camera* c1, const mvs_camera* c2) {
  h->have_cams = (c1 && c2);
  if (h->have_cams) {
    h->cams.c1 = to_camd(*c1);
    h->cams.c2 = to_camd(*c2);
  }
  return MVS_OK;
}

template <class Acc>
int run_two_pass(mvs_srt* h, Acc acc) {
  mvs_ctx* ctx = h->ctx;
  int grid = grid_for(ctx, h->n, kSrtThreads, 4);
  MVS_TRY(h->partials.reserve((size_t)grid * 10));
  {
    ProfScope ps(ctx, "srt_centroids");
    k_srt_centroids<Acc><<<grid, kSrtThreads, 0, ctx->stream>>>(acc, h->n, h->partials.p,
                                                                h->ticket.p, h->res.p);
  }
  {
    ProfScope ps(ctx, "srt_scale_cov", 2);
    k_srt_scale_cov<Acc><<<grid, kSrtThreads, 0, ctx->stream>>>(acc, h->n, h->partials.p,
                                                                h->ticket.p, h->res.p);
    k_srt_finish<<<1, 32, 0, ctx->stream>>>(h->n, h->res.p);
  }
  MVS_CUDA(cudaGetLastError());
  return MVS_OK;
}

template <class Acc>
int run_score(mvs_srt* h, Acc acc, int nhyp, const double* rts_dev) {
  mvs_ctx* ctx = h->ctx;
  int chunks = div_up(nhyp, kHypChunk);
  int64_t tiles_needed = div_up(h->n, kSrtThreads);
  int64_t cap = chunks == 1 ? 4 * (int64_t)ctx->num_sms
                            : std::max<int64_t>(1, 8 * (int64_t)ctx->num_sms / chunks);
  int tiles = (int)std::max<int64_t>(1, std::min<int64_t>(tiles_needed, cap));
  MVS_TRY(h->partials.reserve((size_t)tiles * nhyp + 16));
  MVS_TRY(h->errs.reserve((size_t)nhyp));
  {
    ProfScope ps(ctx, "srt_residual");
    dim3 grid(tiles, chunks);
    k_residual<Acc><<<grid, kSrtThreads, 0, ctx->stream>>>(acc, h->n, h->cams, h->hyps.p, nhyp,
                                                           h->partials.p);
  }
  {
    ProfScope ps(ctx, "srt_residual_final");
    k_residual_final<<<1, 256, 0, ctx->stream>>>(h->partials.p, tiles, nhyp, h->n, h->errs.p,
                                                 rts_dev, h->res.p);
  }
  MVS_CUDA(cudaGetLastError());
  return MVS_OK;
}

template <class Acc>
int run_hyp(mvs_srt* h, Acc acc, int iters) {
  mvs_ctx* ctx = h->ctx;
  ProfScope ps(ctx, "srt_ransac_hyp");
  k_ransac_hyp<Acc><<<div_up(iters, 64), 64, 0, ctx->stream>>>(acc, h->triples.p, iters, h->res.p,
                                                               h->hyps.p, h->rts.p);
  MVS_CUDA(cudaGetLastError());
  return MVS_OK;
}

int check_ready(mvs_srt* h) {
  if (!h) {
    set_error("null SRT handle");
    return MVS_ERR_INVALID_ARG;
  }
  if (h->mode == 0) {
    set_error("SRT solver: estimate called before set_input");
    return MVS_ERR_STATE;
  }
  if (h->n <= 0) {
    set_error("SRT solver: no matches");
    return MVS_ERR_DEGENERATE;
  }
  return ctx_ensure_device(h->ctx);
}

int read_result(mvs_srt* h, double* host /*R_SIZE*/) {
  return d2h_sync(h->ctx, host, h->res.p, sizeof(double) * R_SIZE);
}

}  // namespace

extern "C" {

int mvs_srt_create(mvs_ctx* ctx, mvs_srt** out) {
  if (!out) return MVS_ERR_INVALID_ARG;
  *out = nullptr;
  if (!ctx) ctx = default_ctx();
  if (!ctx) return MVS_ERR_NO_DEVICE;
  mvs_srt* h = new mvs_srt();
  h->ctx = ctx;
  ctx_retain(ctx);
  *out = h;
  return MVS_OK;
}

int mvs_srt_destroy(mvs_srt* h) {
  if (!h) return MVS_OK;
  cudaSetDevice(h->ctx->device);
  cudaStreamSynchronize(h->ctx->stream);
  mvs_ctx* ctx = h->ctx;
  delete h;
  ctx_release(ctx);
  return MVS_OK;
}

int mvs_srt_set_input(mvs_srt* h, const double* p, const double* q, int64_t n,
                      const mvs_camera* cam1, const mvs_camera* cam2) {
  MVS_REQUIRE(h && n >= 0 && ((p && q) || n == 0), "null argument");
  mvs_ctx* ctx = h->ctx;
  MVS_TRY(ctx_ensure_device(ctx));
  MVS_TRY(srt_prepare(h));
  h->n = n;
  h->mode = 1;
  srt_set_cams(h, cam1, cam2);
  if (n == 0) return MVS_OK;
  MVS_TRY(h->planes.reserve((size_t)6 * n));
  DevBuf<double> staging_buf;  // stream-ordered (or the worker's arena): released behind the pack kernel
  MVS_TRY(staging_buf.reserve((size_t)6 * n));
  double* staging = staging_buf.p;
  int st = h2d(ctx, staging, p, sizeof(double) * 3 * n);
  if (st == MVS_OK) st = h2d(ctx, staging + 3 * n, q, sizeof(double) * 3 * n);
  if (st == MVS_OK) {
    ProfScope ps(ctx, "srt_pack");
    k_pack_planes<<<grid_for(ctx, n, 256, 8), 256, 0, ctx->stream>>>(staging, staging + 3 * n,
                                                                     h->planes.p, n);
  }
  MVS_TRY(st);
  MVS_CUDA(cudaGetLastError());
  // SetInput COPIES (Solver/SRTSolver.h:16-19): the caller may reuse or free p/q as soon as this
  // returns.  Pageable sources are staged by the driver before cudaMemcpyAsync returns; a pinned
  // source is read by the DMA engine asynchronously, so only then wait for the stream
  if (!host_is_pageable(p) || !host_is_pageable(q)) MVS_CUDA(cudaStreamSynchronize(ctx->stream));
  return MVS_OK;
}

int mvs_srt_set_input_f32(mvs_srt* h, const float* p, const float* q, int64_t n,
                          const mvs_camera* cam1, const mvs_camera* cam2) {
  MVS_REQUIRE(h && n >= 0 && ((p && q) || n == 0), "null argument");
  mvs_ctx* ctx = h->ctx;
  MVS_TRY(ctx_ensure_device(ctx));
  MVS_TRY(srt_prepare(h));
  h->n = n;
  h->mode = 2;
  srt_set_cams(h, cam1, cam2);
  MVS_TRY(h->pf.reserve((size_t)n));
  MVS_TRY(h->qf.reserve((size_t)n));
  MVS_TRY(h2d(ctx, h->pf.p, p, sizeof(float4) * n));
  MVS_TRY(h2d(ctx, h->qf.p, q, sizeof(float4) * n));
  h->p4 = h->pf.p;
  h->q4 = h->qf.p;
  if (!host_is_pageable(p) || !host_is_pageable(q))
    MVS_CUDA(cudaStreamSynchronize(ctx->stream));  // copy semantics, see mvs_srt_set_input
  return MVS_OK;
}

int mvs_srt_set_input_dev_f32(mvs_srt* h, const float* p_dev, const float* q_dev, int64_t n,
                              const mvs_camera* cam1, const mvs_camera* cam2) {
  MVS_REQUIRE(h && n >= 0 && ((p_dev && q_dev) || n == 0), "null argument");
  MVS_TRY(ctx_ensure_device(h->ctx));
  MVS_TRY(srt_prepare(h));
  h->n = n;
  h->mode = 2;
  srt_set_cams(h, cam1, cam2);
  h->p4 = reinterpret_cast<const float4*>(p_dev);
  h->q4 = reinterpret_cast<const float4*>(q_dev);
  return MVS_OK;
}

int mvs_srt_set_iteration_num(mvs_srt* h, int iter_num) {
  MVS_REQUIRE(h && iter_num >= 0, "bad argument");
  h->iter_num = iter_num;
  return MVS_OK;
}

int mvs_srt_set_print(mvs_srt* h, int on) {
  MVS_REQUIRE(h, "null handle");
  h->print = on != 0;
  return MVS_OK;
}

int mvs_srt_residual(mvs_srt* h, double scale, const double R[9], const double t[3], double* err) {
  MVS_TRY(check_ready(h));
  MVS_REQUIRE(R && t && err, "null argument");
  if (!h->have_cams) {
    set_error("SRT residual needs both cameras (set_input was given NULL)");
    return MVS_ERR_STATE;
  }
  double rec[kHypStride];
  make_hyp_record(scale, R, t, rec);
  MVS_TRY(h->hyps.reserve(kHypStride));
  MVS_TRY(pinned_reserve(h->ctx, sizeof(rec)));
  memcpy(h->ctx->pinned, rec, sizeof(rec));
  MVS_TRY(h2d(h->ctx, h->hyps.p, h->ctx->pinned, sizeof(rec)));
  if (h->mode == 1)
    MVS_TRY(run_score(h, AccF64{h->planes.p, h->n}, 1, nullptr));
  else
    MVS_TRY(run_score(h, AccF4{h->p4, h->q4}, 1, nullptr));
  double e = 0;
  MVS_TRY(d2h_sync(h->ctx, &e, h->errs.p, sizeof(double)));
  *err = e;
  return MVS_OK;
}

int mvs_srt_inliers(mvs_srt* h, double scale, const double R[9], const double t[3], double pixel_thr,
                    uint8_t* flags, double* err_sum, int64_t* n_inliers) {
  MVS_TRY(check_ready(h));
  MVS_REQUIRE(R && t && flags, "null argument");
  if (!h->have_cams) {
    set_error("SRT inlier test needs both cameras (set_input was given NULL)");
    return MVS_ERR_STATE;
  }
  mvs_ctx* ctx = h->ctx;
  MVS_TRY(ctx_ensure_device(ctx));
  MVS_TRY(srt_prepare(h));
  double rec[kHypStride];
  make_hyp_record(scale, R, t, rec);
  MVS_TRY(h->hyps.reserve(kHypStride));
  MVS_TRY(pinned_reserve(ctx, sizeof(rec)));
  memcpy(ctx->pinned, rec, sizeof(rec));
  MVS_TRY(h2d(ctx, h->hyps.p, ctx->pinned, sizeof(rec)));
  DevBuf<unsigned char> dflags;
  MVS_TRY(dflags.reserve((size_t)h->n));
  int grid = grid_for(ctx, h->n, kSrtThreads, 4);
  MVS_TRY(h->partials.reserve((size_t)grid * 2 + 16));
  {
    ProfScope ps(ctx, "srt_inliers");
    if (h->mode == 1)
      k_srt_inliers<AccF64><<<grid, kSrtThreads, 0, ctx->stream>>>(
          AccF64{h->planes.p, h->n}, h->n, h->cams, h->hyps.p, pixel_thr, dflags.p, h->partials.p,
          h->ticket.p, h->res.p + R_ACC);
    else
      k_srt_inliers<AccF4><<<grid, kSrtThreads, 0, ctx->stream>>>(
          AccF4{h->p4, h->q4}, h->n, h->cams, h->hyps.p, pixel_thr, dflags.p, h->partials.p,
          h->ticket.p, h->res.p + R_ACC);
  }
  MVS_CUDA(cudaGetLastError());
  double sums[2];
  MVS_TRY(d2h_async(ctx, flags, dflags.p, (size_t)h->n));
  MVS_TRY(d2h_sync(ctx, sums, h->res.p + R_ACC, sizeof(sums)));
  if (err_sum) *err_sum = sums[0];
  if (n_inliers) *n_inliers = (int64_t)sums[1];
  return MVS_OK;
}

int mvs_srt_estimate(mvs_srt* h, double* scale, double R[9], double t[3]) {
  MVS_TRY(check_ready(h));
  MVS_REQUIRE(scale && R && t, "null argument");
  if (h->mode == 1)
    MVS_TRY(run_two_pass(h, AccF64{h->planes.p, h->n}));
  else
    MVS_TRY(run_two_pass(h, AccF4{h->p4, h->q4}));
  double host[R_SIZE];
  MVS_TRY(read_result(h, host));
  *scale = host[R_SCALE];
  for (int i = 0; i < 9; ++i) R[i] = host[R_ROT + i];
  for (int i = 0; i < 3; ++i) t[i] = host[R_T + i];
  if (h->print) {
    // Solver/SRTSolver.cpp:121-128: the residual is only ever printed
    printf("det(R): %g\n", mat3_det(R));
    if (h->have_cams) {
      double e = 0;
      MVS_TRY(mvs_srt_residual(h, *scale, R, t, &e));
      printf("pixel reproject error = %g\n", e);
    }
  }
  return MVS_OK;
}

int mvs_srt_draw_triples(int64_t n, int iter_num, uint64_t seed, int32_t* triples) {
  if (n < 3 || iter_num < 0 || !triples) {
    set_error("mvs_srt_draw_triples: need n >= 3");
    return MVS_ERR_INVALID_ARG;
  }
  // Common/Utils.h:25-34 Shuffle(k, n, 3) with rand() replaced by a counter-based generator
  for (int it = 0; it < iter_num; ++it) {
    int32_t* k = triples + 3 * it;
    for (int i = 0; i < 3; i++) {
      uint64_t rnd = splitmix64(seed ^ splitmix64((uint64_t)it * 3u + (uint64_t)i + 1u));
      int64_t r = (int64_t)(rnd % (uint64_t)(n - i));
      int j;
      for (j = 0; j < i && r >= k[j]; j++) r++;
      int j0 = j;
      for (j = i; j > j0; j--) k[j] = k[j - 1];
      k[j0] = (int32_t)r;
    }
  }
  return MVS_OK;
}

int mvs_srt_estimate_ransac(mvs_srt* h, const int32_t* triples, uint64_t seed, double* scale,
                            double R[9], double t[3], double* best_err, int32_t* best_iter) {
  MVS_TRY(check_ready(h));
  MVS_REQUIRE(scale && R && t, "null argument");
  if (!h->have_cams) {
    set_error("SRT RANSAC needs both cameras (set_input was given NULL)");
    return MVS_ERR_STATE;
  }
  int iters = h->iter_num;
  if (iters <= 0 || h->n < 3) {
    set_error("SRT RANSAC: need iter_num > 0 and at least 3 matches");
    return MVS_ERR_DEGENERATE;
  }
  mvs_ctx* ctx = h->ctx;
  std::vector<int32_t> drawn;
  if (!triples) {
    drawn.resize((size_t)iters * 3);
    MVS_TRY(mvs_srt_draw_triples(h->n, iters, seed, drawn.data()));
    triples = drawn.data();
  } else {
    for (int i = 0; i < iters * 3; ++i)
      if (triples[i] < 0 || triples[i] >= h->n) {
        set_error("SRT RANSAC: triple index %d out of range", triples[i]);
        return MVS_ERR_INVALID_ARG;
      }
  }
  MVS_TRY(h->triples.reserve((size_t)iters * 3));
  MVS_TRY(h->hyps.reserve((size_t)iters * kHypStride));
  MVS_TRY(h->rts.reserve((size_t)iters * 12));
  MVS_TRY(pinned_reserve(ctx, sizeof(int32_t) * 3 * iters));
  memcpy(ctx->pinned, triples, sizeof(int32_t) * 3 * iters);
  MVS_TRY(h2d(ctx, h->triples.p, ctx->pinned, sizeof(int32_t) * 3 * iters));
  if (h->mode == 1) {
    AccF64 acc{h->planes.p, h->n};
    MVS_TRY(run_two_pass(h, acc));
    MVS_TRY(run_hyp(h, acc, iters));
    MVS_TRY(run_score(h, acc, iters, h->rts.p));
  } else {
    AccF4 acc{h->p4, h->q4};
    MVS_TRY(run_two_pass(h, acc));
    MVS_TRY(run_hyp(h, acc, iters));
    MVS_TRY(run_score(h, acc, iters, h->rts.p));
  }
  double host[R_SIZE];
  MVS_TRY(read_result(h, host));
  *scale = host[R_SCALE];
  for (int i = 0; i < 9; ++i) R[i] = host[R_ROT + i];
  for (int i = 0; i < 3; ++i) t[i] = host[R_T + i];
  if (best_err) *best_err = host[R_BEST_ERR];
  if (best_iter) *best_iter = (int32_t)host[R_BEST_IT];
  if (h->print) {
    std::vector<double> errs((size_t)iters);
    MVS_TRY(d2h_sync(ctx, errs.data(), h->errs.p, sizeof(double) * iters));
    double run = HUGE_VAL;
    for (int k = 0; k < iters; ++k) {
      if (errs[k] < run) run = errs[k];
      printf("Iteration#%d | pixel reproject error = %g\n", k, run);
    }
  }
  return MVS_OK;
}

int mvs_srt_estimate_array(mvs_srt* h, const int32_t* triples, uint64_t seed, double* scale,
                           double R[3][3], double t[3]) {
  double Rf[9];
  MVS_TRY(mvs_srt_estimate_ransac(h, triples, seed, scale, Rf, t, nullptr, nullptr));
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R[i][j] = Rf[3 * i + j];
  return MVS_OK;
}

}  // extern "C"
