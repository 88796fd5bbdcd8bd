// pcg.cu — block-Jacobi preconditioned conjugate gradients on (J^T J + lambda I) delta = -J^T r
// in the Kronecker-compressed block-CSR form (deform.cuh): per pattern entry (j,l) a 4x4 factor
// B_jl with H_jl = kron(B_jl, I3), plus a dense 9x9 rotation block and lambda on the diagonal.
//
// Two kernels per iteration, all scalars stay on the device (no host round trip):
//   k_pcg_spmv    p = z + beta p_prev (recomputed on the fly for gathered rows, written for the own
//                 row), q = H p, pq = p.q (grid reduction)
//   k_pcg_update  alpha = rz/pq; delta += alpha p; r -= alpha q; z = M^-1 r; rz' = r.z, |r|^2
// 16 lanes per node row (12 active: one per unknown), 2 rows per warp; the 12 entries of a
// neighbour's vector are exchanged with warp shuffles inside the 16-lane group.
// Vectors are float32, every dot product is float64 with a fixed reduction tree.
#include "deform.cuh"
#include "reduce.cuh"

namespace mvs {

constexpr int kRowLanes = 16;
constexpr int kPcgThreads = 256;
constexpr int kRowsPerBlock = kPcgThreads / kRowLanes;

enum { S_RZ0 = 0, S_RZ1 = 1, S_PQV = 2, S_BN2 = 3, S_RN2 = 4, S_DONE = 5 };

// r = -g, delta = 0, z = M^-1 r, p = z, rz = r.z, |b|^2
__global__ void __launch_bounds__(kPcgThreads)
    k_pcg_init(const float* __restrict__ rhs, const float* __restrict__ Minv, int64_t m,
               float* __restrict__ r, float* __restrict__ z, float* __restrict__ p,
               float* __restrict__ delta, double* partials, unsigned* ticket, double* scal) {
  __shared__ double smem[32 * 2];
  const unsigned full = 0xffffffffu;
  int lane = threadIdx.x & 31;
  int k = lane & (kRowLanes - 1);
  int gbase = lane & ~(kRowLanes - 1);
  int64_t row = (int64_t)blockIdx.x * kRowsPerBlock + (threadIdx.x / kRowLanes);
  bool act = row < m && k < 12;
  float rv = act ? -rhs[row * 12 + k] : 0.f;
  float zv = 0.f;
#pragma unroll
  for (int q = 0; q < 12; ++q) {
    float rq = __shfl_sync(full, rv, gbase + q);
    if (act) zv += __ldg(Minv + row * 144 + k * 12 + q) * rq;
  }
  if (act) {
    r[row * 12 + k] = rv;
    z[row * 12 + k] = zv;
    p[row * 12 + k] = zv;
    delta[row * 12 + k] = 0.f;
  }
  double v[2] = {(double)rv * (double)zv, (double)rv * (double)rv};
  if (grid_reduce_sum<2>(v, smem, partials, ticket, scal + 6)) {
    if (threadIdx.x == 0) {
      scal[S_RZ0] = scal[6];
      scal[S_BN2] = scal[7];
      scal[S_RN2] = scal[7];
      scal[S_DONE] = 0.0;
    }
  }
}

__global__ void __launch_bounds__(kPcgThreads)
    k_pcg_spmv(const int* __restrict__ row_ptr, const int* __restrict__ col,
               const float* __restrict__ B, const float* __restrict__ Rot, float lambda, int64_t m,
               int it, const float* __restrict__ z, const float* __restrict__ p_prev,
               float* __restrict__ p_cur, float* __restrict__ qv, double* partials,
               unsigned* ticket, double* scal) {
  __shared__ double smem[32];
  const unsigned full = 0xffffffffu;
  int lane = threadIdx.x & 31;
  int k = lane & (kRowLanes - 1);
  int gbase = lane & ~(kRowLanes - 1);
  int64_t row = (int64_t)blockIdx.x * kRowsPerBlock + (threadIdx.x / kRowLanes);
  bool rowok = row < m;
  bool act = rowok && k < 12;
  float beta = 0.f;
  if (it > 0) {
    double rz_old = scal[(it - 1) & 1], rz_new = scal[it & 1];
    beta = rz_old > 0.0 ? (float)(rz_new / rz_old) : 0.f;
  }
  // own row's search direction
  float pj = 0.f;
  if (act) {
    pj = (it == 0) ? p_prev[row * 12 + k] : z[row * 12 + k] + beta * p_prev[row * 12 + k];
    p_cur[row * 12 + k] = pj;
  }
  float y = lambda * pj;
  // dense rotation block on the first 9 unknowns
#pragma unroll
  for (int q = 0; q < 9; ++q) {
    float pq = __shfl_sync(full, pj, gbase + q);
    if (act && k < 9) y += __ldg(Rot + row * 81 + k * 9 + q) * pq;
  }
  int s = rowok ? __ldg(row_ptr + row) : 0;
  int e = rowok ? __ldg(row_ptr + row + 1) : 0;
  int len = e - s;
  int maxlen = max(len, __shfl_xor_sync(full, len, kRowLanes));
  int a = k / 3, rr = k - 3 * a;
  for (int t = 0; t < maxlen; ++t) {
    bool have = t < len;
    int l = have ? __ldg(col + s + t) : 0;
    float pl = 0.f;
    if (have && k < 12) {
      int64_t o = (int64_t)l * 12 + k;
      pl = (it == 0) ? __ldg(p_prev + o) : __ldg(z + o) + beta * __ldg(p_prev + o);
    }
    float v0 = __shfl_sync(full, pl, gbase + rr);
    float v1 = __shfl_sync(full, pl, gbase + 3 + rr);
    float v2 = __shfl_sync(full, pl, gbase + 6 + rr);
    float v3 = __shfl_sync(full, pl, gbase + 9 + rr);
    if (have && act) {
      float4 Ba = __ldg(reinterpret_cast<const float4*>(B + (size_t)(s + t) * 16 + 4 * a));
      y += Ba.x * v0 + Ba.y * v1 + Ba.z * v2 + Ba.w * v3;
    }
  }
  if (act) qv[row * 12 + k] = y;
  double v[1] = {(double)pj * (double)y};
  grid_reduce_sum<1>(v, smem, partials, ticket, scal + S_PQV);
}

__global__ void __launch_bounds__(kPcgThreads)
    k_pcg_update(const float* __restrict__ Minv, int64_t m, int it, double tol,
                 const float* __restrict__ p, const float* __restrict__ qv, float* __restrict__ r,
                 float* __restrict__ z, float* __restrict__ delta, double* partials,
                 unsigned* ticket, double* scal) {
  __shared__ double smem[32 * 2];
  const unsigned full = 0xffffffffu;
  int lane = threadIdx.x & 31;
  int k = lane & (kRowLanes - 1);
  int gbase = lane & ~(kRowLanes - 1);
  int64_t row = (int64_t)blockIdx.x * kRowsPerBlock + (threadIdx.x / kRowLanes);
  bool act = row < m && k < 12;
  double rz = scal[it & 1], pq = scal[S_PQV];
  float alpha = (pq > 0.0 && rz > 0.0 && scal[S_DONE] == 0.0) ? (float)(rz / pq) : 0.f;
  float rv = 0.f;
  if (act) {
    int64_t o = row * 12 + k;
    delta[o] += alpha * p[o];
    rv = r[o] - alpha * qv[o];
    r[o] = rv;
  }
  float zv = 0.f;
#pragma unroll
  for (int q = 0; q < 12; ++q) {
    float rq = __shfl_sync(full, rv, gbase + q);
    if (act) zv += __ldg(Minv + row * 144 + k * 12 + q) * rq;
  }
  if (act) z[row * 12 + k] = zv;
  double v[2] = {(double)rv * (double)zv, (double)rv * (double)rv};
  if (grid_reduce_sum<2>(v, smem, partials, ticket, scal + 6)) {
    if (threadIdx.x == 0) {
      scal[(it + 1) & 1] = scal[6];
      scal[S_RN2] = scal[7];
      if (tol > 0.0 && scal[7] <= tol * tol * scal[S_BN2]) scal[S_DONE] = 1.0;
    }
  }
}

int pcg_solve(mvs_deform* h) {
  mvs_ctx* ctx = h->ctx;
  cudaStream_t st = ctx->stream;
  int64_t m = h->m;
  size_t N = (size_t)m * 12;
  float* r = h->pcg.p;
  float* z = r + N;
  float* p0 = z + N;
  float* p1 = p0 + N;
  float* q = p1 + N;
  float* delta = q + N;
  int grid = div_up(m, kRowsPerBlock);
  MVS_TRY(h->partials.reserve((size_t)grid * 2 + 16));
  {
    ProfScope ps(ctx, "pcg_init");
    k_pcg_init<<<grid, kPcgThreads, 0, st>>>(h->rhs(), h->Minv.p, m, r, z, p0, delta,
                                             h->partials.p, h->ticket.p, h->scal.p);
  }
  int iters = h->prm.pcg_iters;
  {
    ProfScope ps(ctx, "pcg_iterations", 2 * (int64_t)iters);
    for (int it = 0; it < iters; ++it) {
      float* pprev = (it & 1) ? p1 : p0;
      float* pcur = (it & 1) ? p0 : p1;
      if (it == 0) pprev = p0, pcur = p1;
      k_pcg_spmv<<<grid, kPcgThreads, 0, st>>>(h->row_ptr.p, h->col.p, h->B(), h->Rot(),
                                               (float)h->prm.lambda, m, it, z, pprev, pcur, q,
                                               h->partials.p, h->ticket.p, h->scal.p);
      k_pcg_update<<<grid, kPcgThreads, 0, st>>>(h->Minv.p, m, it, h->prm.pcg_tol, pcur, q, r, z,
                                                 delta, h->partials.p, h->ticket.p, h->scal.p);
    }
  }
  MVS_CUDA(cudaGetLastError());
  return MVS_OK;
}

}  // namespace mvs
