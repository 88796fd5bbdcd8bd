// arap.cu — as-rigid-as-possible surface deformation with the sampled nodes as hard constraints:
// the solver the reference runs through CGAL `Surface_mesh_deformation`
// (Deformation/Deformation.cpp:256-260 insert_roi_vertices(all), :383-392 control vertices +
// target positions, :393 preprocess(), :398 deform(5, 1e-4), :400 overwrite_initial_geometry()).
// CGAL 4.6 is not under /root/reference: the published algorithm (Sorkine & Alexa 2007, CGAL's
// ORIGINAL_ARAP mode) is restated at that call site, specification in oracle/arap.cpp; parity with
// CGAL's own floating-point results is unpinned.
//
// Device formulation (all arrays in ORIGINAL vertex order, float4 records):
//   build (once per rest geometry):  half-edge (i, j, cot/2) triples from the facets -> radix sort by
//     (i, j) -> run heads -> CSR {row_ptr, col, w = max(0, sum)} + diag.  6 x 12 B per facet.
//   global step: Jacobi-preconditioned CG on the cotangent Laplacian, the three coordinate columns
//     in lockstep with their own alpha/beta; per iteration three streaming kernels
//       k_arap_spmv   q = L p on the free rows, p.q          8 B/nnz + 32 B/vertex
//       k_arap_update x += a p, r -= a q, z = r / d, r.z     112 B/vertex
//       k_arap_dir    p = z + b p                            48 B/vertex
//     scalars stay on the device (deterministic last-CTA reductions, reduce.cuh); HBM bound.
//   local step: per vertex S = sum w e e'^T, 3x3 one-sided Jacobi SVD in fp64, R = V U^T, energy.
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cmath>

#include "deform.cuh"
#include "linalg3.h"
#include "reduce.cuh"

namespace mvs {

struct ArapState {
  int64_t n = 0, nnz = 0;
  bool built = false;
  DevBuf<int> row_ptr, col;
  DevBuf<float> w, diag;
  DevBuf<float4> P, X, B, r, z, p, q;
  DevBuf<float> R;  // n x 12 (9 used, 16-byte aligned rows)
  DevBuf<int> fixed;
  DevBuf<double> scal, partials;
  DevBuf<unsigned> ticket;
};

void arap_free(ArapState* a) { delete a; }

// ------------------------------------------------------------------ cotangent Laplacian
__global__ void k_arap_halfedges(const float4* __restrict__ p, const int* __restrict__ facets,
                                 int64_t nf, unsigned long long* __restrict__ keys,
                                 float* __restrict__ vals) {
  int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (f >= nf) return;
  int v[3] = {facets[3 * f], facets[3 * f + 1], facets[3 * f + 2]};
  double P[3][3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float4 q = __ldg(p + v[c]);
    P[c][0] = q.x; P[c][1] = q.y; P[c][2] = q.z;
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int a = (c + 1) % 3, b = (c + 2) % 3;
    double u[3], w[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      u[k] = __dsub_rn(P[a][k], P[c][k]);
      w[k] = __dsub_rn(P[b][k], P[c][k]);
    }
    // FMA-free so the float32 weights are bit-identical to the host restatement
    double dot = __dadd_rn(__dadd_rn(__dmul_rn(u[0], w[0]), __dmul_rn(u[1], w[1])), __dmul_rn(u[2], w[2]));
    double cx = __dsub_rn(__dmul_rn(u[1], w[2]), __dmul_rn(u[2], w[1]));
    double cy = __dsub_rn(__dmul_rn(u[2], w[0]), __dmul_rn(u[0], w[2]));
    double cz = __dsub_rn(__dmul_rn(u[0], w[1]), __dmul_rn(u[1], w[0]));
    double cr = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(cx, cx), __dmul_rn(cy, cy)), __dmul_rn(cz, cz)));
    float hc = cr > 0.0 ? (float)__ddiv_rn(__dmul_rn(0.5, dot), cr) : 0.0f;
    size_t o = (size_t)f * 6 + 2 * c;
    keys[o] = ((unsigned long long)(unsigned)v[a] << 32) | (unsigned)v[b];
    vals[o] = hc;
    keys[o + 1] = ((unsigned long long)(unsigned)v[b] << 32) | (unsigned)v[a];
    vals[o + 1] = hc;
  }
}

__global__ void k_arap_heads(const unsigned long long* __restrict__ keys, int64_t n,
                             int* __restrict__ flag) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  flag[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// heads write their edge: column, clamped sum of the run (left to right: the sort is stable, so
// equal keys keep their facet-major emission order)
__global__ void k_arap_compact(const unsigned long long* __restrict__ keys,
                               const float* __restrict__ vals, const int* __restrict__ scan,
                               int64_t n, unsigned long long* __restrict__ ukeys,
                               int* __restrict__ col, float* __restrict__ w) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long k = keys[i];
  if (i > 0 && keys[i - 1] == k) return;
  float s = 0.f;
  for (int64_t t = i; t < n && keys[t] == k; ++t) s = __fadd_rn(s, vals[t]);
  int r = scan[i];
  ukeys[r] = k;
  col[r] = (int)(k & 0xffffffffull);
  w[r] = s > 0.f ? s : 0.f;
}

__global__ void k_arap_rows(const unsigned long long* __restrict__ ukeys, int nnz, int64_t n,
                            int* __restrict__ row_ptr) {
  int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (v > n) return;
  unsigned long long want = (unsigned long long)v << 32;
  int lo = 0, hi = nnz;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (ukeys[mid] < want) lo = mid + 1; else hi = mid;
  }
  row_ptr[v] = lo;
}

__global__ void k_arap_diag(const int* __restrict__ row_ptr, const float* __restrict__ w, int64_t n,
                            float* __restrict__ diag) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float d = 0.f;
  for (int t = row_ptr[i]; t < row_ptr[i + 1]; ++t) d = __fadd_rn(d, w[t]);
  diag[i] = d;
}

static int arap_build(mvs_deform* h) {
  mvs_ctx* ctx = h->ctx;
  cudaStream_t st = ctx->stream;
  if (!h->arap) h->arap = new ArapState();
  ArapState& A = *h->arap;
  int64_t n = h->n, nf = h->nf;
  if (nf <= 0) {
    set_error("arap: the mesh has no facets");
    return MVS_ERR_STATE;
  }
  if (nf * 6 >= ((int64_t)1 << 31)) {
    set_error("arap: too many facets");
    return MVS_ERR_CAPACITY;
  }
  int64_t ne = nf * 6;
  DevBuf<unsigned long long> keys, keys2, ukeys;
  DevBuf<float> vals, vals2;
  DevBuf<int> flag, scan_tmp;
  DevBuf<char> tmp;
  MVS_TRY(keys.reserve((size_t)ne));
  MVS_TRY(keys2.reserve((size_t)ne));
  MVS_TRY(vals.reserve((size_t)ne));
  MVS_TRY(vals2.reserve((size_t)ne));
  MVS_TRY(flag.reserve((size_t)ne + 2));
  {
    ProfScope ps(ctx, "arap_halfedges");
    k_arap_halfedges<<<div_up(nf, 128), 128, 0, st>>>(h->v0.p, h->facets.p, nf, keys.p, vals.p);
  }
  size_t tmp_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys.p, keys2.p, vals.p, vals2.p, (int)ne, 0,
                                  64, st);
  MVS_TRY(tmp.reserve(tmp_bytes));
  {
    ProfScope ps(ctx, "arap_sort", 8);
    MVS_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, keys.p, keys2.p, vals.p, vals2.p,
                                             (int)ne, 0, 64, st));
  }
  {
    ProfScope ps(ctx, "arap_heads");
    k_arap_heads<<<div_up(ne, 256), 256, 0, st>>>(keys2.p, ne, flag.p);
  }
  MVS_CUDA(cudaMemsetAsync(flag.p + ne, 0, sizeof(int) * 2, st));
  MVS_TRY(exclusive_scan_inplace(ctx, flag.p, ne, scan_tmp));
  int nnz = 0;
  MVS_TRY(d2h_sync(ctx, &nnz, flag.p + ne, sizeof(int)));
  A.n = n;
  A.nnz = nnz;
  MVS_TRY(ukeys.reserve((size_t)nnz + 1));
  MVS_TRY(A.col.reserve((size_t)nnz + 1));
  MVS_TRY(A.w.reserve((size_t)nnz + 1));
  MVS_TRY(A.row_ptr.reserve((size_t)n + 2));
  MVS_TRY(A.diag.reserve((size_t)n));
  {
    ProfScope ps(ctx, "arap_csr", 3);
    k_arap_compact<<<div_up(ne, 256), 256, 0, st>>>(keys2.p, vals2.p, flag.p, ne, ukeys.p, A.col.p,
                                                    A.w.p);
    k_arap_rows<<<div_up(n + 1, 256), 256, 0, st>>>(ukeys.p, nnz, n, A.row_ptr.p);
    k_arap_diag<<<div_up(n, 256), 256, 0, st>>>(A.row_ptr.p, A.w.p, n, A.diag.p);
  }
  MVS_CUDA(cudaGetLastError());
  for (DevBuf<float4>* b : {&A.P, &A.X, &A.B, &A.r, &A.z, &A.p, &A.q}) MVS_TRY(b->reserve((size_t)n));
  MVS_TRY(A.R.reserve((size_t)n * 12));
  MVS_TRY(A.fixed.reserve((size_t)n));
  MVS_TRY(A.scal.reserve(32));
  MVS_TRY(A.ticket.reserve(1));
  MVS_CUDA(cudaMemsetAsync(A.ticket.p, 0, sizeof(unsigned), st));
  MVS_TRY(A.partials.reserve((size_t)ctx->num_sms * 8 * 6 + 64));
  MVS_CUDA(cudaStreamSynchronize(st));  // the temporaries above are released in stream order
  A.built = true;
  return MVS_OK;
}

// ------------------------------------------------------------------ solve
// scal[]: 0..2 rz, 3..5 pq, 6..8 rz_new, 9..11 rr, 12..14 bb, 15 energy
enum { A_RZ = 0, A_PQ = 3, A_RZN = 6, A_RR = 9, A_BB = 12, A_E = 15 };

__global__ void k_arap_init(const float4* __restrict__ v0, int64_t n, float4* __restrict__ P,
                            float4* __restrict__ X, float* __restrict__ R, int* __restrict__ fixed) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = v0[i];
  p.w = 0.f;
  P[i] = p;
  X[i] = p;
  fixed[i] = 0;
  R[(size_t)i * 12 + 0] = 1.f; R[(size_t)i * 12 + 1] = 0.f; R[(size_t)i * 12 + 2] = 0.f;
  R[(size_t)i * 12 + 3] = 0.f; R[(size_t)i * 12 + 4] = 1.f; R[(size_t)i * 12 + 5] = 0.f;
  R[(size_t)i * 12 + 6] = 0.f; R[(size_t)i * 12 + 7] = 0.f; R[(size_t)i * 12 + 8] = 1.f;
  R[(size_t)i * 12 + 9] = 0.f; R[(size_t)i * 12 + 10] = 0.f; R[(size_t)i * 12 + 11] = 0.f;
}

__global__ void k_arap_constraints(const int* __restrict__ ids, const double* __restrict__ xyz,
                                   int64_t m, int64_t n, float4* __restrict__ X,
                                   int* __restrict__ fixed) {
  int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= m) return;
  int v = ids[c];
  if (v < 0 || v >= n) return;
  fixed[v] = 1;
  X[v] = make_float4((float)xyz[3 * c], (float)xyz[3 * c + 1], (float)xyz[3 * c + 2], 0.f);
}

// b_i = sum_j w_ij / 2 (R_i + R_j)(p_i - p_j)   (free rows; zero on the fixed ones)
__global__ void __launch_bounds__(256)
    k_arap_rhs(const int* __restrict__ row_ptr, const int* __restrict__ col,
               const float* __restrict__ w, const float4* __restrict__ P,
               const float* __restrict__ R, const int* __restrict__ fixed, int64_t n,
               float4* __restrict__ B) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float bx = 0.f, by = 0.f, bz = 0.f;
  if (!fixed[i]) {
    float4 pi = __ldg(P + i);
    const float4* ri4 = reinterpret_cast<const float4*>(R + (size_t)i * 12);
    float4 a0 = __ldg(ri4), a1 = __ldg(ri4 + 1), a2 = __ldg(ri4 + 2);
    float Ri[9] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, a2.x};
    for (int t = __ldg(row_ptr + i); t < __ldg(row_ptr + i + 1); ++t) {
      int j = __ldg(col + t);
      float hw = 0.5f * __ldg(w + t);
      float4 pj = __ldg(P + j);
      const float4* rj4 = reinterpret_cast<const float4*>(R + (size_t)j * 12);
      float4 c0 = __ldg(rj4), c1 = __ldg(rj4 + 1), c2 = __ldg(rj4 + 2);
      float Rj[9] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w, c2.x};
      float e0 = pi.x - pj.x, e1 = pi.y - pj.y, e2 = pi.z - pj.z;
      bx += hw * ((Ri[0] + Rj[0]) * e0 + (Ri[1] + Rj[1]) * e1 + (Ri[2] + Rj[2]) * e2);
      by += hw * ((Ri[3] + Rj[3]) * e0 + (Ri[4] + Rj[4]) * e1 + (Ri[5] + Rj[5]) * e2);
      bz += hw * ((Ri[6] + Rj[6]) * e0 + (Ri[7] + Rj[7]) * e1 + (Ri[8] + Rj[8]) * e2);
    }
  }
  B[i] = make_float4(bx, by, bz, 0.f);
}

// y = L x on the free rows (0 on fixed rows); optional per-column dot x.y -> scal[dot_slot..+3)
template <bool kResidual>
__global__ void __launch_bounds__(256)
    k_arap_spmv(const int* __restrict__ row_ptr, const int* __restrict__ col,
                const float* __restrict__ w, const float4* __restrict__ x,
                const int* __restrict__ fixed, int64_t n, const float4* __restrict__ B,
                const float* __restrict__ diag, float4* __restrict__ y, float4* __restrict__ zout,
                float4* __restrict__ pout, double* partials, unsigned* ticket, double* scal) {
  __shared__ double smem[32 * 6];
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  double acc[6] = {0, 0, 0, 0, 0, 0};
  if (i < n) {
    float sx = 0.f, sy = 0.f, sz = 0.f;
    float4 xi = __ldg(x + i);
    const bool fx = fixed[i] != 0;
    if (!fx) {
      for (int t = __ldg(row_ptr + i); t < __ldg(row_ptr + i + 1); ++t) {
        float4 xj = __ldg(x + __ldg(col + t));
        float wt = __ldg(w + t);
        sx += wt * (xi.x - xj.x);
        sy += wt * (xi.y - xj.y);
        sz += wt * (xi.z - xj.z);
      }
    }
    if (kResidual) {
      // r = b - L x, z = r / d, p = z; rz and bb = |r|^2
      float4 b = __ldg(B + i);
      float d = __ldg(diag + i);
      float rx = fx ? 0.f : b.x - sx, ry = fx ? 0.f : b.y - sy, rz = fx ? 0.f : b.z - sz;
      float id = d > 0.f ? 1.0f / d : 0.f;
      float4 zz = make_float4(rx * id, ry * id, rz * id, 0.f);
      y[i] = make_float4(rx, ry, rz, 0.f);
      zout[i] = zz;
      pout[i] = zz;
      acc[0] = (double)rx * zz.x; acc[1] = (double)ry * zz.y; acc[2] = (double)rz * zz.z;
      acc[3] = (double)rx * rx; acc[4] = (double)ry * ry; acc[5] = (double)rz * rz;
    } else {
      y[i] = make_float4(sx, sy, sz, 0.f);
      acc[0] = (double)xi.x * sx; acc[1] = (double)xi.y * sy; acc[2] = (double)xi.z * sz;
    }
  }
  if (kResidual) {
    if (grid_reduce_sum<6>(acc, smem, partials, ticket, scal + 16) && threadIdx.x == 0) {
      for (int a = 0; a < 3; ++a) {
        scal[A_RZ + a] = scal[16 + a];
        scal[A_BB + a] = scal[19 + a];
        scal[A_RR + a] = scal[19 + a];
      }
    }
  } else {
    double a3[3] = {acc[0], acc[1], acc[2]};
    grid_reduce_sum<3>(a3, smem, partials, ticket, scal + A_PQ);
  }
}

__global__ void __launch_bounds__(256)
    k_arap_update(const float4* __restrict__ p, const float4* __restrict__ q,
                  const float* __restrict__ diag, int64_t n, float4* __restrict__ X,
                  float4* __restrict__ r, float4* __restrict__ z, double* partials,
                  unsigned* ticket, double* scal) {
  __shared__ double smem[32 * 6];
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  float al[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    double pq = scal[A_PQ + a], rz = scal[A_RZ + a];
    al[a] = (pq > 0.0 && rz > 0.0) ? (float)(rz / pq) : 0.f;
  }
  double acc[6] = {0, 0, 0, 0, 0, 0};
  if (i < n) {
    float4 pi = __ldg(p + i), qi = __ldg(q + i), xi = X[i], ri = r[i];
    float d = __ldg(diag + i);
    float id = d > 0.f ? 1.0f / d : 0.f;
    xi.x += al[0] * pi.x; xi.y += al[1] * pi.y; xi.z += al[2] * pi.z;
    ri.x -= al[0] * qi.x; ri.y -= al[1] * qi.y; ri.z -= al[2] * qi.z;
    float4 zi = make_float4(ri.x * id, ri.y * id, ri.z * id, 0.f);
    X[i] = xi;
    r[i] = ri;
    z[i] = zi;
    acc[0] = (double)ri.x * zi.x; acc[1] = (double)ri.y * zi.y; acc[2] = (double)ri.z * zi.z;
    acc[3] = (double)ri.x * ri.x; acc[4] = (double)ri.y * ri.y; acc[5] = (double)ri.z * ri.z;
  }
  if (grid_reduce_sum<6>(acc, smem, partials, ticket, scal + 16) && threadIdx.x == 0) {
    for (int a = 0; a < 3; ++a) {
      scal[A_RZN + a] = scal[16 + a];
      scal[A_RR + a] = scal[19 + a];
    }
  }
}

// p = z + beta p; the last thread rolls rz <- rz_new for the next iteration
__global__ void __launch_bounds__(256)
    k_arap_dir(const float4* __restrict__ z, int64_t n, float4* __restrict__ p,
               const double* __restrict__ scal) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float be[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    double rz = scal[A_RZ + a];
    be[a] = rz > 0.0 ? (float)(scal[A_RZN + a] / rz) : 0.f;
  }
  float4 zi = __ldg(z + i), pi = p[i];
  p[i] = make_float4(zi.x + be[0] * pi.x, zi.y + be[1] * pi.y, zi.z + be[2] * pi.z, 0.f);
}

__global__ void k_arap_roll(double* scal) {
  if (threadIdx.x < 3) scal[A_RZ + threadIdx.x] = scal[A_RZN + threadIdx.x];
}

// local step: rotation of every vertex from its one-ring + energy at the new rotation
__global__ void __launch_bounds__(128)
    k_arap_local(const int* __restrict__ row_ptr, const int* __restrict__ col,
                 const float* __restrict__ w, const float4* __restrict__ P,
                 const float4* __restrict__ X, int64_t n, float* __restrict__ R, double* partials,
                 unsigned* ticket, double* scal) {
  __shared__ double smem[32];
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  double e[1] = {0.0};
  if (i < n) {
    float4 pi = __ldg(P + i), xi = __ldg(X + i);
    double S[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    int t0 = __ldg(row_ptr + i), t1 = __ldg(row_ptr + i + 1);
    for (int t = t0; t < t1; ++t) {
      int j = __ldg(col + t);
      double wt = (double)__ldg(w + t);
      float4 pj = __ldg(P + j), xj = __ldg(X + j);
      double ep[3] = {(double)pi.x - pj.x, (double)pi.y - pj.y, (double)pi.z - pj.z};
      double ex[3] = {(double)xi.x - xj.x, (double)xi.y - xj.y, (double)xi.z - xj.z};
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) S[3 * a + b] += wt * ep[a] * ex[b];
    }
    double U[9], V[9], sig[3], Ri[9];
    svd3(S, U, sig, V);
    mat3_mul_bt(V, U, Ri);
    if (mat3_det(Ri) < 0.0) {
      V[2] = -V[2]; V[5] = -V[5]; V[8] = -V[8];
      mat3_mul_bt(V, U, Ri);
    }
    float* out = R + (size_t)i * 12;
#pragma unroll
    for (int k = 0; k < 9; ++k) out[k] = (float)Ri[k];
    for (int t = t0; t < t1; ++t) {
      int j = __ldg(col + t);
      double wt = (double)__ldg(w + t);
      float4 pj = __ldg(P + j), xj = __ldg(X + j);
      double ep[3] = {(double)pi.x - pj.x, (double)pi.y - pj.y, (double)pi.z - pj.z};
      double d2 = 0.0;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        double re = Ri[3 * a] * ep[0] + Ri[3 * a + 1] * ep[1] + Ri[3 * a + 2] * ep[2];
        double xa = a == 0 ? (double)xi.x - xj.x : (a == 1 ? (double)xi.y - xj.y : (double)xi.z - xj.z);
        double d = xa - re;
        d2 += d * d;
      }
      e[0] += wt * d2;
    }
  }
  grid_reduce_sum<1>(e, smem, partials, ticket, scal + A_E);
}

int arap_solve(mvs_deform* h, const double* ctrl_xyz_host, const int32_t* ctrl_ids_host, int64_t m,
               const mvs_arap_params* prm, double* energy_out, int* iters_out) {
  mvs_ctx* ctx = h->ctx;
  cudaStream_t st = ctx->stream;
  if (!h->arap || !h->arap->built || h->arap->n != h->n) MVS_TRY(arap_build(h));
  ArapState& A = *h->arap;
  int64_t n = h->n;
  int grid = (int)div_up(n, 256);
  MVS_TRY(h->cur0.reserve((size_t)n));
  MVS_TRY(h->curn0.reserve((size_t)n));
  MVS_TRY(A.partials.reserve((size_t)std::max(grid, (int)div_up(n, 128)) * 6 + 64));
  DevBuf<int> ids_d;
  DevBuf<double> xyz_d;
  MVS_TRY(ids_d.reserve((size_t)std::max<int64_t>(m, 1)));
  MVS_TRY(xyz_d.reserve((size_t)std::max<int64_t>(m, 1) * 3));
  if (m > 0) {
    MVS_CUDA(cudaMemcpyAsync(ids_d.p, ctrl_ids_host, sizeof(int) * m, cudaMemcpyHostToDevice, st));
    MVS_CUDA(cudaMemcpyAsync(xyz_d.p, ctrl_xyz_host, sizeof(double) * 3 * m, cudaMemcpyHostToDevice, st));
  }
  {
    ProfScope ps(ctx, "arap_init", 2);
    k_arap_init<<<grid, 256, 0, st>>>(h->v0.p, n, A.P.p, A.X.p, A.R.p, A.fixed.p);
    if (m > 0)
      k_arap_constraints<<<div_up(m, 128), 128, 0, st>>>(ids_d.p, xyz_d.p, m, n, A.X.p, A.fixed.p);
  }
  MVS_CUDA(cudaGetLastError());
  double E_prev = 0.0, E = 0.0;
  int it = 0;
  const int check_every = 16;
  for (; it < prm->iters; ++it) {
    {
      ProfScope ps(ctx, "arap_rhs");
      k_arap_rhs<<<grid, 256, 0, st>>>(A.row_ptr.p, A.col.p, A.w.p, A.P.p, A.R.p, A.fixed.p, n, A.B.p);
    }
    {
      ProfScope ps(ctx, "arap_cg_init");
      k_arap_spmv<true><<<grid, 256, 0, st>>>(A.row_ptr.p, A.col.p, A.w.p, A.X.p, A.fixed.p, n, A.B.p,
                                              A.diag.p, A.r.p, A.z.p, A.p.p, A.partials.p,
                                              A.ticket.p, A.scal.p);
    }
    for (int k = 0; k < prm->cg_iters; ++k) {
      if (prm->cg_tol > 0.0 && k % check_every == 0) {
        double s[16];
        MVS_TRY(d2h_sync(ctx, s, A.scal.p, sizeof(double) * 16));
        if (s[A_RR] + s[A_RR + 1] + s[A_RR + 2] <=
            prm->cg_tol * prm->cg_tol * (s[A_BB] + s[A_BB + 1] + s[A_BB + 2]))
          break;
      }
      ProfScope ps(ctx, "arap_cg_iter", 4);
      k_arap_spmv<false><<<grid, 256, 0, st>>>(A.row_ptr.p, A.col.p, A.w.p, A.p.p, A.fixed.p, n,
                                               nullptr, nullptr, A.q.p, nullptr, nullptr,
                                               A.partials.p, A.ticket.p, A.scal.p);
      k_arap_update<<<grid, 256, 0, st>>>(A.p.p, A.q.p, A.diag.p, n, A.X.p, A.r.p, A.z.p,
                                          A.partials.p, A.ticket.p, A.scal.p);
      k_arap_dir<<<grid, 256, 0, st>>>(A.z.p, n, A.p.p, A.scal.p);
      k_arap_roll<<<1, 32, 0, st>>>(A.scal.p);
    }
    {
      ProfScope ps(ctx, "arap_local");
      k_arap_local<<<div_up(n, 128), 128, 0, st>>>(A.row_ptr.p, A.col.p, A.w.p, A.P.p, A.X.p, n,
                                                   A.R.p, A.partials.p, A.ticket.p, A.scal.p);
    }
    MVS_CUDA(cudaGetLastError());
    MVS_TRY(d2h_sync(ctx, &E, A.scal.p + A_E, sizeof(double)));
    if (prm->tol > 0.0 && it > 0 && std::fabs(E_prev - E) <= prm->tol * E) {
      ++it;
      break;
    }
    E_prev = E;
  }
  // deformed geometry -> cur0 (original order) + normals
  MVS_CUDA(cudaMemcpyAsync(h->cur0.p, A.X.p, sizeof(float4) * n, cudaMemcpyDeviceToDevice, st));
  MVS_TRY(gn_recompute_normals(h));
  if (energy_out) *energy_out = E;
  if (iters_out) *iters_out = it;
  return MVS_OK;
}

}  // namespace mvs
