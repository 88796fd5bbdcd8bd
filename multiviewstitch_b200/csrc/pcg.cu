// pcg.cu — block-Jacobi preconditioned conjugate gradients on (J^T J + lambda I) delta = -J^T r
// in the Kronecker-compressed block-CSR form (deform.cuh): per pattern entry (j,l) a 4x4 factor
// B_jl with H_jl = kron(B_jl, I3), plus a dense 9x9 rotation block and lambda on the diagonal.
//
// ONE persistent cooperative kernel runs the whole solve (all PCG iterations + x += delta):
//   * a CTA owns a contiguous range of node rows; its slice of the matrix (the 4x4 factors and
//     their column indices) is staged into shared memory once and reused by every iteration, so per
//     iteration only vectors move (L2-resident);
//   * 16 lanes per node row (12 active: one per unknown); r, z, p, delta of a row live in
//     registers for the whole solve (RS row slots per lane group);
//   * two grid-wide barriers per iteration (after p.q and after r.z); dot products are float64 with
//     a fixed reduction order (lane -> warp -> CTA partial in global memory -> every CTA adds the
//     partials in CTA order), so the solve is bit-reproducible run to run;
//   * the search direction of a neighbour row is rebuilt on the fly as z_l + beta p_prev_l, which
//     saves the third barrier a separate "p = z + beta p" phase would need.
// Standard PCG recurrences, identical to oracle/deform.cpp pcg(); vectors float32, scalars float64.
#include <cooperative_groups.h>

#include <algorithm>

#include "deform.cuh"

namespace cg = cooperative_groups;

namespace mvs {

constexpr int kRowLanes = 16;
constexpr int kPcgThreads = 256;
constexpr int kGroups = kPcgThreads / kRowLanes;  // row groups per CTA

struct PcgArgs {
  const int* row_ptr;
  const int* col;
  const float* B;
  const float* Rot;
  const float* Minv;
  const float* rhs;
  float* x;       // node state, updated in place at the end
  float* zbuf;    // m x 12
  float* pbuf;    // 2 x m x 12
  float* delta;   // m x 12 (exported for inspection)
  double* part;   // 3 x grid partial sums
  double* scal;   // S_* slots
  int64_t m;
  int iters;
  float lambda;
  double tol;
  int rows_per_cta;
  int stage_smem;  // 1: matrix slice staged in shared memory
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// CTA-wide sum of N doubles -> thread 0 writes them to out[0..N)
template <int N>
__device__ __forceinline__ void cta_partial(double (&v)[N], double* smem, double* out) {
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < N; ++j) v[j] = warp_sum(v[j]);
  if (lane == 0)
#pragma unroll
    for (int j = 0; j < N; ++j) smem[warp * N + j] = v[j];
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int j = 0; j < N; ++j) {
      double s = 0.0;
      for (int w = 0; w < kPcgThreads / 32; ++w) s += smem[w * N + j];
      out[j] = s;
    }
  }
  __syncthreads();
}

// Totals of N per-CTA partials (stride 3 doubles per CTA) after a grid barrier: every thread
// loads at most two partials (independent loads: one L2 latency instead of a serial chain), then a
// fixed-order warp / CTA reduction; all threads return the same totals.
template <int N>
__device__ __forceinline__ void grid_totals(const double* part, int ncta, double* smem,
                                            double (&out)[N]) {
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double v[N];
#pragma unroll
  for (int j = 0; j < N; ++j) v[j] = 0.0;
  int c0 = threadIdx.x, c1 = threadIdx.x + kPcgThreads;
  double a0[N], a1[N];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    a0[j] = c0 < ncta ? __ldcg(part + (size_t)c0 * 3 + j) : 0.0;
    a1[j] = c1 < ncta ? __ldcg(part + (size_t)c1 * 3 + j) : 0.0;
  }
#pragma unroll
  for (int j = 0; j < N; ++j) v[j] = a0[j] + a1[j];
  for (int c = threadIdx.x + 2 * kPcgThreads; c < ncta; c += kPcgThreads)
#pragma unroll
    for (int j = 0; j < N; ++j) v[j] += __ldcg(part + (size_t)c * 3 + j);
#pragma unroll
  for (int j = 0; j < N; ++j) v[j] = warp_sum(v[j]);
  __syncthreads();  // smem may still be read by the previous use
  if (lane == 0)
#pragma unroll
    for (int j = 0; j < N; ++j) smem[warp * N + j] = v[j];
  __syncthreads();
#pragma unroll
  for (int j = 0; j < N; ++j) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kPcgThreads / 32; ++w) t += smem[w * N + j];
    out[j] = t;
  }
}

// zp[buf][row][k] = (z, p): z of the row after the update of an iteration, p its search direction;
// iteration `it` reads buffer (it+1)&1 (written during iteration it-1) and writes buffer it&1.
template <int RS>
__global__ void __launch_bounds__(kPcgThreads, 2) k_pcg_persistent(PcgArgs a) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ float4 smem_dyn[];
  __shared__ double red[(kPcgThreads / 32) * 2];
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int k = lane & (kRowLanes - 1);
  const int gbase = lane & ~(kRowLanes - 1);
  const int gi = threadIdx.x / kRowLanes;
  const int ncta = gridDim.x;
  const int64_t m = a.m;
  const int64_t row0 = (int64_t)blockIdx.x * a.rows_per_cta;
  const int64_t row1 = min(m, row0 + a.rows_per_cta);
  const bool lane_on = k < 12;
  const int ba = k / 3, rr = k - 3 * ba;
  float2* zp = reinterpret_cast<float2*>(a.zbuf);  // 2 x m x 12 float2

  // ---- stage this CTA's matrix slice
  const int nz0 = row0 < m ? __ldg(a.row_ptr + row0) : 0;
  const int nz1 = row0 < m ? __ldg(a.row_ptr + row1) : 0;
  const float4* Bsrc = reinterpret_cast<const float4*>(a.B) + (size_t)nz0 * 4;
  const int* csrc = a.col + nz0;
  const float4* Bs = Bsrc;
  const int* cs = csrc;
  if (a.stage_smem) {
    int nb = (nz1 - nz0) * 4;
    int* scol = reinterpret_cast<int*>(smem_dyn + nb);
    for (int i = threadIdx.x; i < nb; i += kPcgThreads) smem_dyn[i] = __ldg(Bsrc + i);
    for (int i = threadIdx.x; i < nz1 - nz0; i += kPcgThreads) scol[i] = __ldg(csrc + i);
    Bs = smem_dyn;
    cs = scol;
    __syncthreads();
  }

  float r[RS], z[RS], p[RS], d[RS];
  int t0s[RS], t1s[RS];
  // with one row slot the rows of M^-1 and of the rotation block stay in registers for the solve
  float minv[RS == 1 ? 12 : 1], rot[RS == 1 ? 9 : 1];
  // ---- init: r = -g, z = M^-1 r, delta = 0, p_prev = 0
  double acc[2] = {0.0, 0.0};
#pragma unroll
  for (int s = 0; s < RS; ++s) {
    int64_t j = row0 + s * kGroups + gi;
    bool on = j < row1;
    bool act = on && lane_on;
    t0s[s] = act ? __ldg(a.row_ptr + j) - nz0 : 0;
    t1s[s] = act ? __ldg(a.row_ptr + j + 1) - nz0 : 0;
    if (RS == 1) {
#pragma unroll
      for (int q = 0; q < 12; ++q) minv[q] = act ? __ldg(a.Minv + j * 144 + k * 12 + q) : 0.f;
#pragma unroll
      for (int q = 0; q < 9; ++q) rot[q] = (act && k < 9) ? __ldg(a.Rot + j * 81 + k * 9 + q) : 0.f;
    }
    r[s] = act ? -__ldg(a.rhs + j * 12 + k) : 0.f;
    float zv = 0.f;
#pragma unroll
    for (int q = 0; q < 12; ++q) {
      float rq = __shfl_sync(full, r[s], gbase + q);
      if (RS == 1)
        zv += minv[q] * rq;
      else if (act)
        zv += __ldg(a.Minv + j * 144 + k * 12 + q) * rq;
    }
    z[s] = zv;
    p[s] = 0.f;
    d[s] = 0.f;
    if (act) zp[(size_t)m * 12 + j * 12 + k] = make_float2(zv, 0.f);  // buffer 1: read by iteration 0
    acc[0] += (double)r[s] * (double)zv;
    acc[1] += (double)r[s] * (double)r[s];
  }
  cta_partial<2>(acc, red, a.part + (size_t)blockIdx.x * 3 + 1);
  grid.sync();
  double rz, bn2;
  {
    double tot[2];
    grid_totals<2>(a.part + 1, ncta, red, tot);
    rz = tot[0];
    bn2 = tot[1];
  }
  double rn2 = bn2;
  double rz_old = 0.0;
  int it = 0;
  for (; it < a.iters; ++it) {
    if (a.tol > 0.0 && rn2 <= a.tol * a.tol * bn2) break;  // uniform over the grid
    // ---- phase A: p = z + beta p ; q = (H + lambda) p ; p.q
    float beta = (it > 0 && rz_old > 0.0) ? (float)(rz / rz_old) : 0.f;
    const float2* zprev = zp + (size_t)((it + 1) & 1) * m * 12;
    float2* zcur = zp + (size_t)(it & 1) * m * 12;
    float qv[RS];
    double pq[1] = {0.0};
#pragma unroll
    for (int s = 0; s < RS; ++s) {
      int64_t j = row0 + s * kGroups + gi;
      bool on = j < row1;
      bool act = on && lane_on;
      p[s] = z[s] + beta * p[s];
      if (act) zcur[j * 12 + k].y = p[s];
      float y = a.lambda * p[s];
#pragma unroll
      for (int q = 0; q < 9; ++q) {
        float pqv = __shfl_sync(full, p[s], gbase + q);
        if (RS == 1)
          y += rot[q] * pqv;
        else if (act && k < 9)
          y += __ldg(a.Rot + j * 81 + k * 9 + q) * pqv;
      }
      if (act) {
        int t = t0s[s];
        const int t1 = t1s[s];
        for (; t + 8 <= t1; t += 8) {  // 8 neighbours (32 loads) in flight
          float2 v[8][4];
          float4 Bv[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            int l = cs[t + u];
            Bv[u] = Bs[(size_t)(t + u) * 4 + ba];
            const float2* src = zprev + (size_t)l * 12 + rr;
#pragma unroll
            for (int b = 0; b < 4; ++b) v[u][b] = __ldcg(src + 3 * b);
          }
#pragma unroll
          for (int u = 0; u < 8; ++u)
            y += Bv[u].x * (v[u][0].x + beta * v[u][0].y) + Bv[u].y * (v[u][1].x + beta * v[u][1].y) +
                 Bv[u].z * (v[u][2].x + beta * v[u][2].y) + Bv[u].w * (v[u][3].x + beta * v[u][3].y);
        }
        for (; t + 2 <= t1; t += 2) {
          float2 v[2][4];
          float4 Bv[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            int l = cs[t + u];
            Bv[u] = Bs[(size_t)(t + u) * 4 + ba];
            const float2* src = zprev + (size_t)l * 12 + rr;
#pragma unroll
            for (int b = 0; b < 4; ++b) v[u][b] = __ldcg(src + 3 * b);
          }
#pragma unroll
          for (int u = 0; u < 2; ++u)
            y += Bv[u].x * (v[u][0].x + beta * v[u][0].y) + Bv[u].y * (v[u][1].x + beta * v[u][1].y) +
                 Bv[u].z * (v[u][2].x + beta * v[u][2].y) + Bv[u].w * (v[u][3].x + beta * v[u][3].y);
        }
        if (t < t1) {
          int l = cs[t];
          float4 Bv = Bs[(size_t)t * 4 + ba];
          const float2* src = zprev + (size_t)l * 12 + rr;
          float2 v0 = __ldcg(src), v1 = __ldcg(src + 3), v2 = __ldcg(src + 6), v3 = __ldcg(src + 9);
          y += Bv.x * (v0.x + beta * v0.y) + Bv.y * (v1.x + beta * v1.y) +
               Bv.z * (v2.x + beta * v2.y) + Bv.w * (v3.x + beta * v3.y);
        }
      } else {
        y = 0.f;
      }
      qv[s] = y;
      pq[0] += (double)p[s] * (double)y;
    }
    cta_partial<1>(pq, red, a.part + (size_t)blockIdx.x * 3);
    grid.sync();
    // ---- phase B: alpha ; delta += alpha p ; r -= alpha q ; z = M^-1 r ; r.z, |r|^2
    double pqa[1];
    grid_totals<1>(a.part, ncta, red, pqa);
    double pqt = pqa[0];
    float alpha = (pqt > 0.0 && rz > 0.0) ? (float)(rz / pqt) : 0.f;
    acc[0] = acc[1] = 0.0;
#pragma unroll
    for (int s = 0; s < RS; ++s) {
      int64_t j = row0 + s * kGroups + gi;
      bool act = (j < row1) && lane_on;
      d[s] += alpha * p[s];
      r[s] -= alpha * qv[s];
      float zv = 0.f;
#pragma unroll
      for (int q = 0; q < 12; ++q) {
        float rq = __shfl_sync(full, r[s], gbase + q);
        if (RS == 1)
          zv += minv[q] * rq;
        else if (act)
          zv += __ldg(a.Minv + j * 144 + k * 12 + q) * rq;
      }
      z[s] = zv;
      if (act) zcur[j * 12 + k].x = zv;
      acc[0] += (double)r[s] * (double)zv;
      acc[1] += (double)r[s] * (double)r[s];
    }
    cta_partial<2>(acc, red, a.part + (size_t)blockIdx.x * 3 + 1);
    grid.sync();
    rz_old = rz;
    {
      double tot[2];
      grid_totals<2>(a.part + 1, ncta, red, tot);
      rz = tot[0];
      rn2 = tot[1];
    }
  }
  // ---- x += delta
#pragma unroll
  for (int s = 0; s < RS; ++s) {
    int64_t j = row0 + s * kGroups + gi;
    if (j < row1 && lane_on) {
      a.x[j * 12 + k] += d[s];
      a.delta[j * 12 + k] = d[s];
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    a.scal[S_RZ] = rz;
    a.scal[S_BNORM2] = bn2;
    a.scal[S_RNORM2] = rn2;
    a.scal[S_PQ] = (double)it;  // iterations actually run
  }
}

template <int RS>
static int launch_pcg(mvs_deform* h, PcgArgs& a, int grid, size_t smem) {
  auto kern = k_pcg_persistent<RS>;
  if (smem > 48 * 1024)
    MVS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  void* args[] = {&a};
  MVS_CUDA(cudaLaunchCooperativeKernel((void*)kern, dim3(grid), dim3(kPcgThreads), args, smem,
                                       h->ctx->stream));
  return MVS_OK;
}

template <int RS>
static int max_coresident(size_t smem, int num_sms, int* out) {
  auto kern = k_pcg_persistent<RS>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      cudaGetLastError();
      *out = 0;
      return MVS_OK;
    }
  }
  int per_sm = 0;
  MVS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kPcgThreads, smem));
  *out = per_sm * num_sms;
  return MVS_OK;
}

// Picks the row-slot count RS, the grid and whether the matrix slice fits in shared memory.
static int plan_pcg(mvs_deform* h) {
  if (h->pcg_plan_valid) return MVS_OK;
  mvs_ctx* ctx = h->ctx;
  int64_t m = h->m;
  const int rs_opts[4] = {1, 2, 4, 8};
  for (int stage = 1; stage >= 0; --stage) {
    for (int oi = 0; oi < 4; ++oi) {
      int RS = rs_opts[oi];
      int rpc = RS * kGroups;
      int grid = (int)((m + rpc - 1) / rpc);
      size_t smem = 0;
      if (stage) {
        int worst = 0;
        for (int64_t r0 = 0; r0 < m; r0 += rpc) {
          int64_t r1 = std::min<int64_t>(m, r0 + rpc);
          worst = std::max(worst, h->h_row_ptr[r1] - h->h_row_ptr[r0]);
        }
        smem = (size_t)worst * (64 + 4) + 16;
        if (smem > 200 * 1024) continue;
      }
      int cap = 0;
      switch (RS) {
        case 1: MVS_TRY(max_coresident<1>(smem, ctx->num_sms, &cap)); break;
        case 2: MVS_TRY(max_coresident<2>(smem, ctx->num_sms, &cap)); break;
        case 4: MVS_TRY(max_coresident<4>(smem, ctx->num_sms, &cap)); break;
        default: MVS_TRY(max_coresident<8>(smem, ctx->num_sms, &cap)); break;
      }
      if (grid <= cap) {
        h->pcg_rs = RS;
        h->pcg_grid = grid;
        h->pcg_smem = smem;
        h->pcg_stage = stage;
        h->pcg_plan_valid = true;
        return MVS_OK;
      }
    }
  }
  set_error("pcg: %lld nodes exceed the co-resident capacity of the persistent solver (max %d rows)",
            (long long)m, 8 * kGroups * ctx->num_sms * 2);
  return MVS_ERR_CAPACITY;
}

int pcg_solve(mvs_deform* h) {
  mvs_ctx* ctx = h->ctx;
  MVS_TRY(plan_pcg(h));
  int64_t m = h->m;
  size_t N = (size_t)m * 12;
  PcgArgs a;
  a.row_ptr = h->row_ptr.p;
  a.col = h->col.p;
  a.B = h->B();
  a.Rot = h->Rot();
  a.Minv = h->Minv.p;
  a.rhs = h->rhs();
  a.x = h->x.p;
  a.zbuf = h->pcg.p;          // 2 x N float2 = 4N floats
  a.pbuf = nullptr;
  a.delta = h->pcg.p + 4 * N;
  MVS_TRY(h->partials.reserve((size_t)h->pcg_grid * 3 + 16));
  a.part = h->partials.p;
  a.scal = h->scal.p;
  a.m = m;
  a.iters = h->prm.pcg_iters;
  a.lambda = (float)h->prm.lambda;
  a.tol = h->prm.pcg_tol;
  a.rows_per_cta = h->pcg_rs * kGroups;
  a.stage_smem = h->pcg_stage;
  ProfScope ps(ctx, "pcg_solve");
  switch (h->pcg_rs) {
    case 1: MVS_TRY(launch_pcg<1>(h, a, h->pcg_grid, h->pcg_smem)); break;
    case 2: MVS_TRY(launch_pcg<2>(h, a, h->pcg_grid, h->pcg_smem)); break;
    case 4: MVS_TRY(launch_pcg<4>(h, a, h->pcg_grid, h->pcg_smem)); break;
    default: MVS_TRY(launch_pcg<8>(h, a, h->pcg_grid, h->pcg_smem)); break;
  }
  MVS_CUDA(cudaGetLastError());
  return MVS_OK;
}

}  // namespace mvs
