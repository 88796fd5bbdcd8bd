// nn_grid.cu — build of the uniform grid (bbox -> cell histogram -> scan -> scatter) and the exact
// kNN query kernels.  HBM-bound integer/float streaming work: float4 loads, one thread per query,
// queries binned by coarse cell first so that a warp's lanes walk neighbouring cells (L1/L2 reuse).
#include <float.h>

#include <algorithm>
#include <cmath>

#include "nn_grid.cuh"

namespace mvs {

// ------------------------------------------------------------------ conversions / uploads
__global__ void k_xyz_f32_to_f4(const float* __restrict__ in, float4* __restrict__ out, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    out[i] = make_float4(in[3 * i], in[3 * i + 1], in[3 * i + 2], 0.f);
  }
}
__global__ void k_xyz_f64_to_f4(const double* __restrict__ in, float4* __restrict__ out,
                                int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    out[i] = make_float4((float)in[3 * i], (float)in[3 * i + 1], (float)in[3 * i + 2], 0.f);
  }
}

template <typename T>
static int upload_xyz(mvs_ctx* ctx, const T* xyz_host, int64_t n, DevBuf<float4>& out) {
  MVS_TRY(out.reserve((size_t)n));
  if (n == 0) return MVS_OK;
  T* staging = nullptr;
  MVS_CUDA(cudaMallocAsync((void**)&staging, sizeof(T) * 3 * n, ctx->stream));
  int st = h2d(ctx, staging, xyz_host, sizeof(T) * 3 * n);
  if (st == MVS_OK) {
    ProfScope ps(ctx, "convert_xyz");
    int grid = grid_for(ctx, n, 256, 8);
    if constexpr (sizeof(T) == 4)
      k_xyz_f32_to_f4<<<grid, 256, 0, ctx->stream>>>((const float*)staging, out.p, n);
    else
      k_xyz_f64_to_f4<<<grid, 256, 0, ctx->stream>>>((const double*)staging, out.p, n);
  }
  cudaFreeAsync(staging, ctx->stream);
  MVS_TRY(st);
  MVS_CUDA(cudaGetLastError());
  return MVS_OK;
}
int upload_xyz_f32(mvs_ctx* ctx, const float* xyz_host, int64_t n, DevBuf<float4>& out) {
  return upload_xyz<float>(ctx, xyz_host, n, out);
}
int upload_xyz_f64(mvs_ctx* ctx, const double* xyz_host, int64_t n, DevBuf<float4>& out) {
  return upload_xyz<double>(ctx, xyz_host, n, out);
}

// ------------------------------------------------------------------ exclusive scan (int32)
constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ int block_exclusive_scan(int v, int* smem /*32 ints*/, int* total) {
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) smem[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int nw = blockDim.x >> 5;
    int w = lane < nw ? smem[lane] : 0;
    int wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    smem[lane] = wi - w;  // exclusive warp offsets
    if (lane == nw - 1) smem[32] = wi;
  }
  __syncthreads();
  int res = incl - v + smem[warp];
  if (total) *total = smem[32];
  __syncthreads();
  return res;
}

__global__ void k_scan_tile_sums(const int* __restrict__ data, int64_t n, int* __restrict__ sums) {
  __shared__ int sm[33];
  int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int s = 0;
  if (base + kScanItems <= n) {
    const int4* p = reinterpret_cast<const int4*>(data + base);
#pragma unroll
    for (int j = 0; j < kScanItems / 4; ++j) {
      int4 v = p[j];
      s += v.x + v.y + v.z + v.w;
    }
  } else {
    for (int j = 0; j < kScanItems; ++j)
      if (base + j < n) s += data[base + j];
  }
  int total;
  block_exclusive_scan(s, sm, &total);
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void k_scan_sums(int* __restrict__ sums, int nb, int* __restrict__ total_out) {
  __shared__ int sm[33];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += blockDim.x) {
    int i = base + threadIdx.x;
    int v = i < nb ? sums[i] : 0;
    int total;
    int ex = block_exclusive_scan(v, sm, &total);
    int c = carry;
    if (i < nb) sums[i] = ex + c;
    __syncthreads();
    if (threadIdx.x == 0) carry = c + total;
    __syncthreads();
  }
  if (threadIdx.x == 0 && total_out) *total_out = carry;
}

__global__ void k_scan_apply(int* __restrict__ data, int64_t n, const int* __restrict__ sums) {
  __shared__ int sm[33];
  int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int v[kScanItems];
  bool full = base + kScanItems <= n;
  if (full) {
    const int4* p = reinterpret_cast<const int4*>(data + base);
#pragma unroll
    for (int j = 0; j < kScanItems / 4; ++j) {
      int4 t = p[j];
      v[4 * j] = t.x;
      v[4 * j + 1] = t.y;
      v[4 * j + 2] = t.z;
      v[4 * j + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) v[j] = (base + j < n) ? data[base + j] : 0;
  }
  int s = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) s += v[j];
  int off = block_exclusive_scan(s, sm, nullptr) + sums[blockIdx.x];
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    int t = v[j];
    v[j] = off;
    off += t;
  }
  if (full) {
    int4* p = reinterpret_cast<int4*>(data + base);
#pragma unroll
    for (int j = 0; j < kScanItems / 4; ++j)
      p[j] = make_int4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  } else {
#pragma unroll
    for (int j = 0; j < kScanItems; ++j)
      if (base + j < n) data[base + j] = v[j];
  }
}

// data[0..n) -> exclusive prefix sums in place; data[n] (must be allocated) receives the total.
int exclusive_scan_inplace(mvs_ctx* ctx, int* data, int64_t n, DevBuf<int>& tmp) {
  if (n <= 0) return MVS_OK;
  int nb = div_up(n, kScanTile);
  MVS_TRY(tmp.reserve((size_t)nb + 1));
  ProfScope ps(ctx, "scan", 3);
  k_scan_tile_sums<<<nb, kScanThreads, 0, ctx->stream>>>(data, n, tmp.p);
  k_scan_sums<<<1, 1024, 0, ctx->stream>>>(tmp.p, nb, data + n);
  k_scan_apply<<<nb, kScanThreads, 0, ctx->stream>>>(data, n, tmp.p);
  MVS_CUDA(cudaGetLastError());
  return MVS_OK;
}

// ------------------------------------------------------------------ bbox
__global__ void k_bbox_partial(const float4* __restrict__ pts, int64_t n, float* __restrict__ out) {
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float4 p = __ldg(pts + i);
    mn[0] = fminf(mn[0], p.x);
    mn[1] = fminf(mn[1], p.y);
    mn[2] = fminf(mn[2], p.z);
    mx[0] = fmaxf(mx[0], p.x);
    mx[1] = fmaxf(mx[1], p.y);
    mx[2] = fmaxf(mx[2], p.z);
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
      mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
    }
  }
  __shared__ float sm[32][6];
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    for (int a = 0; a < 3; ++a) {
      sm[warp][a] = mn[a];
      sm[warp][3 + a] = mx[a];
    }
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    int nw = blockDim.x >> 5;
    float v = sm[0][threadIdx.x];
    for (int w = 1; w < nw; ++w)
      v = threadIdx.x < 3 ? fminf(v, sm[w][threadIdx.x]) : fmaxf(v, sm[w][threadIdx.x]);
    out[blockIdx.x * 6 + threadIdx.x] = v;
  }
}

int bbox_host(mvs_ctx* ctx, const float4* pts, int64_t n, DevBuf<float>& red, float mn[3],
              float mx[3]) {
  int grid = grid_for(ctx, n, 256, 4);
  MVS_TRY(red.reserve((size_t)grid * 6));
  {
    ProfScope ps(ctx, "bbox");
    k_bbox_partial<<<grid, 256, 0, ctx->stream>>>(pts, n, red.p);
  }
  MVS_CUDA(cudaGetLastError());
  std::vector<float> h((size_t)grid * 6);
  MVS_TRY(d2h_sync(ctx, h.data(), red.p, h.size() * sizeof(float)));
  for (int a = 0; a < 3; ++a) {
    mn[a] = FLT_MAX;
    mx[a] = -FLT_MAX;
  }
  for (int b = 0; b < grid; ++b)
    for (int a = 0; a < 3; ++a) {
      mn[a] = std::min(mn[a], h[b * 6 + a]);
      mx[a] = std::max(mx[a], h[b * 6 + 3 + a]);
    }
  return MVS_OK;
}

// ------------------------------------------------------------------ grid build
__global__ void k_cell_hist(GridParams g, const float4* __restrict__ pts, int* __restrict__ hist) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < g.n; i += stride) {
    float4 p = __ldg(pts + i);
    atomicAdd(hist + grid_cell(g, p.x, p.y, p.z), 1);
  }
}

__global__ void k_count_occupied(const int* __restrict__ hist, int64_t ncells,
                                 unsigned long long* __restrict__ out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int c = 0;
  for (; i < ncells; i += stride) c += hist[i] > 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, (unsigned long long)c);
}

__global__ void k_cell_scatter(GridParams g, const float4* __restrict__ pts,
                               const int* __restrict__ cell_start, int* __restrict__ fill,
                               float4* __restrict__ sorted) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < g.n; i += stride) {
    float4 p = __ldg(pts + i);
    int64_t c = grid_cell(g, p.x, p.y, p.z);
    int pos = cell_start[c] + atomicAdd(fill + c, 1);
    p.w = __int_as_float((int)i);
    sorted[pos] = p;
  }
}

static void choose_dims(const float mn[3], const float mx[3], float h, GridParams& g) {
  g.h = h;
  g.inv_h = 1.0f / h;
  g.ncells = 1;
  for (int a = 0; a < 3; ++a) {
    g.bmin[a] = mn[a];
    double e = (double)mx[a] - (double)mn[a];
    int d = (int)std::floor(e / h) + 1;
    g.dims[a] = std::max(1, d);
    g.ncells *= g.dims[a];
  }
}

constexpr int64_t kMaxCells = (int64_t)1 << 27;
constexpr int kMaxDim = 2048;

static float clamp_h(const float mn[3], const float mx[3], float h) {
  // enforce per-axis and total cell caps by growing h
  for (int it = 0; it < 64; ++it) {
    GridParams g{};
    choose_dims(mn, mx, h, g);
    bool ok = g.ncells <= kMaxCells;
    for (int a = 0; a < 3; ++a) ok = ok && g.dims[a] <= kMaxDim;
    if (ok) return h;
    h *= 1.26f;
  }
  return h;
}

int nn_build(mvs_nn_index* idx, const float4* pts, int64_t n, float target_ppc) {
  mvs_ctx* ctx = idx->ctx;
  MVS_TRY(ctx_ensure_device(ctx));
  idx->built = false;
  if (n <= 0) {
    set_error("nn_build: empty point set");
    return MVS_ERR_DEGENERATE;
  }
  if (n >= (int64_t)INT32_MAX) {
    set_error("nn_build: more than 2^31-1 points");
    return MVS_ERR_INVALID_ARG;
  }
  float mn[3], mx[3];
  MVS_TRY(bbox_host(ctx, pts, n, idx->red, mn, mx));
  double ext[3], emax = 0;
  for (int a = 0; a < 3; ++a) {
    ext[a] = (double)mx[a] - (double)mn[a];
    emax = std::max(emax, ext[a]);
  }
  GridParams g{};
  g.n = n;
  if (!(emax > 0) || !std::isfinite(emax)) {
    // all points coincide (or non-finite data): a single cell
    float h1 = 1.0f;
    choose_dims(mn, mx, h1, g);
    g.dims[0] = g.dims[1] = g.dims[2] = 1;
    g.ncells = 1;
  } else {
    double vol = 1.0;
    for (int a = 0; a < 3; ++a) vol *= std::max(ext[a], 1e-3 * emax);
    float h = (float)std::cbrt(vol * target_ppc / (double)n);
    h = clamp_h(mn, mx, h);
    choose_dims(mn, mx, h, g);
  }
  g.n = n;

  DevBuf<unsigned long long> occ;
  MVS_TRY(occ.reserve(1));
  for (int attempt = 0; attempt < 3; ++attempt) {
    MVS_TRY(idx->cell_start.reserve((size_t)g.ncells + 2));
    MVS_CUDA(cudaMemsetAsync(idx->cell_start.p, 0, sizeof(int) * (g.ncells + 2), ctx->stream));
    {
      ProfScope ps(ctx, "grid_hist");
      k_cell_hist<<<grid_for(ctx, n, 256, 8), 256, 0, ctx->stream>>>(g, pts, idx->cell_start.p);
    }
    MVS_CUDA(cudaGetLastError());
    if (attempt == 2 || g.ncells == 1) break;
    // occupancy check: surfaces occupy O(1/h^2) cells, so the volumetric guess leaves too many
    // points per occupied cell; shrink h once or twice using the measured occupancy.
    MVS_CUDA(cudaMemsetAsync(occ.p, 0, sizeof(unsigned long long), ctx->stream));
    {
      ProfScope ps(ctx, "grid_occupancy");
      k_count_occupied<<<grid_for(ctx, g.ncells, 256, 8), 256, 0, ctx->stream>>>(
          idx->cell_start.p, g.ncells, occ.p);
    }
    unsigned long long hocc = 0;
    MVS_TRY(d2h_sync(ctx, &hocc, occ.p, sizeof(hocc)));
    double ppc = (double)n / (double)std::max<unsigned long long>(hocc, 1);
    if (ppc <= 2.0 * target_ppc) break;
    float h_new = g.h * (float)std::sqrt(target_ppc / ppc);
    h_new = clamp_h(mn, mx, h_new);
    if (h_new >= g.h * 0.95f) break;
    choose_dims(mn, mx, h_new, g);
    g.n = n;
  }
  MVS_TRY(exclusive_scan_inplace(ctx, idx->cell_start.p, g.ncells, idx->scan_tmp));
  MVS_TRY(idx->cell_fill.reserve((size_t)g.ncells));
  MVS_CUDA(cudaMemsetAsync(idx->cell_fill.p, 0, sizeof(int) * g.ncells, ctx->stream));
  MVS_TRY(idx->pts_sorted.reserve((size_t)n));
  {
    ProfScope ps(ctx, "grid_scatter");
    k_cell_scatter<<<grid_for(ctx, n, 256, 8), 256, 0, ctx->stream>>>(
        g, pts, idx->cell_start.p, idx->cell_fill.p, idx->pts_sorted.p);
  }
  MVS_CUDA(cudaGetLastError());
  idx->g = g;
  idx->built = true;
  idx->pyr_built = false;
  return MVS_OK;
}

// ------------------------------------------------------------------ query binning
struct CoarseParams {
  int shift;
  int dims[3];
  int ncells;
};

__device__ __forceinline__ int coarse_cell(const GridParams& g, const CoarseParams& cp, float4 q) {
  int cx = grid_coord(q.x, g.bmin[0], g.inv_h, g.dims[0]) >> cp.shift;
  int cy = grid_coord(q.y, g.bmin[1], g.inv_h, g.dims[1]) >> cp.shift;
  int cz = grid_coord(q.z, g.bmin[2], g.inv_h, g.dims[2]) >> cp.shift;
  return (cz * cp.dims[1] + cy) * cp.dims[0] + cx;
}

__global__ void k_query_hist(GridParams g, CoarseParams cp, const float4* __restrict__ q,
                             int64_t nq, int* __restrict__ qcell, int* __restrict__ hist) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < nq; i += stride) {
    int c = coarse_cell(g, cp, __ldg(q + i));
    qcell[i] = c;
    atomicAdd(hist + c, 1);
  }
}

__global__ void k_query_scatter(const int* __restrict__ qcell, int64_t nq,
                                int* __restrict__ start /*consumed as cursor*/,
                                int* __restrict__ perm) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < nq; i += stride) {
    int pos = atomicAdd(start + qcell[i], 1);
    perm[pos] = (int)i;
  }
}

// ------------------------------------------------------------------ kNN kernel
template <int K>
__global__ void __launch_bounds__(128)
    k_knn(GridParams g, const float4* __restrict__ pts, const int* __restrict__ cell_start,
          const float4* __restrict__ q, const int* __restrict__ perm, int64_t nq, int k,
          int* __restrict__ out_idx, float* __restrict__ out_d2) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= nq) return;
  int64_t qi = perm ? perm[t] : t;
  float4 qq = __ldg(q + qi);
  KnnList<K> list;
  list.init(qq.x, qq.y, qq.z);
  grid_search(g, cell_start, pts, qq.x, qq.y, qq.z, list);
#pragma unroll
  for (int j = 0; j < K; ++j) {
    if (j < k) {
      int id = list.idx[j];
      if (out_idx) out_idx[qi * k + j] = (id == 0x7fffffff) ? -1 : id;
      if (out_d2) out_d2[qi * k + j] = list.d2[j];
    }
  }
}

template <int K>
static void launch_knn(mvs_nn_index* idx, const float4* q, const int* perm, int64_t nq, int k,
                       int* out_idx, float* out_d2) {
  mvs_ctx* ctx = idx->ctx;
  int grid = div_up(nq, 128);
  k_knn<K><<<grid, 128, 0, ctx->stream>>>(idx->g, idx->pts_sorted.p, idx->cell_start.p, q, perm,
                                          nq, k, out_idx, out_d2);
}

int nn_knn(mvs_nn_index* idx, const float4* q, int64_t nq, int k, int* out_idx, float* out_d2) {
  mvs_ctx* ctx = idx->ctx;
  MVS_TRY(ctx_ensure_device(ctx));
  if (!idx->built) {
    set_error("nn_knn: index not built");
    return MVS_ERR_STATE;
  }
  if (k < 1 || k > 32) {
    set_error("nn_knn: k must be in 1..32 (got %d)", k);
    return MVS_ERR_INVALID_ARG;
  }
  if (nq <= 0) return MVS_OK;
  if (nq >= (int64_t)INT32_MAX / 32) {
    set_error("nn_knn: too many queries for one call");
    return MVS_ERR_INVALID_ARG;
  }
  const int* perm = nullptr;
  if (nq >= 16384) {
    // bin queries by coarse cell (<= 2^18 bins) so neighbouring threads walk neighbouring cells
    CoarseParams cp{};
    cp.shift = 0;
    for (;;) {
      int64_t nc = 1;
      for (int a = 0; a < 3; ++a) {
        cp.dims[a] = ((idx->g.dims[a] - 1) >> cp.shift) + 1;
        nc *= cp.dims[a];
      }
      if (nc <= (1 << 18)) {
        cp.ncells = (int)nc;
        break;
      }
      cp.shift++;
    }
    MVS_TRY(idx->q_cell.reserve((size_t)nq));
    MVS_TRY(idx->q_perm.reserve((size_t)nq));
    MVS_TRY(idx->q_hist.reserve((size_t)cp.ncells + 2));
    MVS_CUDA(cudaMemsetAsync(idx->q_hist.p, 0, sizeof(int) * (cp.ncells + 2), ctx->stream));
    {
      ProfScope ps(ctx, "query_bin", 2);
      k_query_hist<<<grid_for(ctx, nq, 256, 8), 256, 0, ctx->stream>>>(idx->g, cp, q, nq,
                                                                       idx->q_cell.p, idx->q_hist.p);
    }
    MVS_TRY(exclusive_scan_inplace(ctx, idx->q_hist.p, cp.ncells, idx->scan_tmp));
    k_query_scatter<<<grid_for(ctx, nq, 256, 8), 256, 0, ctx->stream>>>(idx->q_cell.p, nq,
                                                                        idx->q_hist.p, idx->q_perm.p);
    MVS_CUDA(cudaGetLastError());
    perm = idx->q_perm.p;
  }
  if (k == 1) {
    // 1-NN: sparse occupancy pyramid (built lazily from the cell-sorted copy, .w = original index)
    if (!idx->pyr_built)
      MVS_TRY(nn_build_pyramid(idx, idx->pts_sorted.p, idx->g.n, kPyrPpcNear, true));
    return nn_nn1_pyramid(idx, q, perm, nq, out_idx, out_d2);
  }
  {
    ProfScope ps(ctx, "knn");
    if (k <= 4)
      launch_knn<4>(idx, q, perm, nq, k, out_idx, out_d2);
    else if (k <= 5)
      launch_knn<5>(idx, q, perm, nq, k, out_idx, out_d2);
    else if (k <= 9)
      launch_knn<9>(idx, q, perm, nq, k, out_idx, out_d2);
    else if (k <= 16)
      launch_knn<16>(idx, q, perm, nq, k, out_idx, out_d2);
    else
      launch_knn<32>(idx, q, perm, nq, k, out_idx, out_d2);
  }
  MVS_CUDA(cudaGetLastError());
  return MVS_OK;
}

}  // namespace mvs

// ==================================================================== C ABI
using namespace mvs;

extern "C" {

int mvs_nn_index_create(mvs_ctx* ctx, mvs_nn_index** out) {
  if (!out) return MVS_ERR_INVALID_ARG;
  *out = nullptr;
  if (!ctx) ctx = default_ctx();
  if (!ctx) return MVS_ERR_NO_DEVICE;
  mvs_nn_index* i = new mvs_nn_index();
  i->ctx = ctx;
  ctx_retain(ctx);
  *out = i;
  return MVS_OK;
}

int mvs_nn_index_destroy(mvs_nn_index* idx) {
  if (!idx) return MVS_OK;
  cudaSetDevice(idx->ctx->device);
  cudaStreamSynchronize(idx->ctx->stream);
  mvs_ctx* ctx = idx->ctx;
  delete idx;
  ctx_release(ctx);
  return MVS_OK;
}

int mvs_nn_index_build_host(mvs_nn_index* idx, const float* pts_xyz, int64_t n) {
  MVS_REQUIRE(idx && (pts_xyz || n == 0), "null argument");
  MVS_TRY(ctx_ensure_device(idx->ctx));
  MVS_TRY(upload_xyz_f32(idx->ctx, pts_xyz, n, idx->pts_in));
  return nn_build(idx, idx->pts_in.p, n);
}

int mvs_nn_index_build_host_f64(mvs_nn_index* idx, const double* pts_xyz, int64_t n) {
  MVS_REQUIRE(idx && (pts_xyz || n == 0), "null argument");
  MVS_TRY(ctx_ensure_device(idx->ctx));
  MVS_TRY(upload_xyz_f64(idx->ctx, pts_xyz, n, idx->pts_in));
  return nn_build(idx, idx->pts_in.p, n);
}

int mvs_nn_index_build_dev(mvs_nn_index* idx, const float* pts_xyzw_dev, int64_t n) {
  MVS_REQUIRE(idx && (pts_xyzw_dev || n == 0), "null argument");
  return nn_build(idx, reinterpret_cast<const float4*>(pts_xyzw_dev), n);
}

static int knn_host_common(mvs_nn_index* idx, int64_t nq, int k, int32_t* out_idx, float* out_d2) {
  mvs_ctx* ctx = idx->ctx;
  MVS_TRY(idx->o_idx.reserve((size_t)nq * k));
  MVS_TRY(idx->o_d2.reserve((size_t)nq * k));
  MVS_TRY(nn_knn(idx, idx->q_in.p, nq, k, idx->o_idx.p, idx->o_d2.p));
  if (out_idx)
    MVS_TRY(d2h_async(ctx, out_idx, idx->o_idx.p, sizeof(int) * nq * k));
  if (out_d2)
    MVS_TRY(d2h_async(ctx, out_d2, idx->o_d2.p, sizeof(float) * nq * k));
  MVS_CUDA(cudaStreamSynchronize(ctx->stream));
  return MVS_OK;
}

int mvs_nn_knn_host(mvs_nn_index* idx, const float* q_xyz, int64_t nq, int k, int32_t* out_idx,
                    float* out_d2) {
  MVS_REQUIRE(idx && (q_xyz || nq == 0), "null argument");
  MVS_TRY(ctx_ensure_device(idx->ctx));
  if (nq == 0) return MVS_OK;
  MVS_TRY(upload_xyz_f32(idx->ctx, q_xyz, nq, idx->q_in));
  return knn_host_common(idx, nq, k, out_idx, out_d2);
}

int mvs_nn_knn_host_f64(mvs_nn_index* idx, const double* q_xyz, int64_t nq, int k,
                        int32_t* out_idx, float* out_d2) {
  MVS_REQUIRE(idx && (q_xyz || nq == 0), "null argument");
  MVS_TRY(ctx_ensure_device(idx->ctx));
  if (nq == 0) return MVS_OK;
  MVS_TRY(upload_xyz_f64(idx->ctx, q_xyz, nq, idx->q_in));
  return knn_host_common(idx, nq, k, out_idx, out_d2);
}

int mvs_nn_knn_dev(mvs_nn_index* idx, const float* q_xyzw_dev, int64_t nq, int k,
                   int32_t* out_idx_dev, float* out_d2_dev) {
  MVS_REQUIRE(idx && (q_xyzw_dev || nq == 0), "null argument");
  return nn_knn(idx, reinterpret_cast<const float4*>(q_xyzw_dev), nq, k, out_idx_dev, out_d2_dev);
}

int mvs_nn_search(mvs_ctx* ctx, const float* q_xyz, int64_t nq, const float* t_xyz, int64_t nt,
                  int k, int32_t* out_idx, float* out_d2) {
  mvs_nn_index* idx = nullptr;
  MVS_TRY(mvs_nn_index_create(ctx, &idx));
  int st = mvs_nn_index_build_host(idx, t_xyz, nt);
  if (st == MVS_OK) st = mvs_nn_knn_host(idx, q_xyz, nq, k, out_idx, out_d2);
  mvs_nn_index_destroy(idx);
  return st;
}

int mvs_part_recog(mvs_ctx* ctx, const double* template_xyz, int64_t n_template,
                   const int32_t* template_labels, const double* query_xyz, int64_t n_query,
                   int32_t* out_labels) {
  MVS_REQUIRE(template_xyz && template_labels && (query_xyz || n_query == 0) &&
                  (out_labels || n_query == 0),
              "null argument");
  mvs_nn_index* idx = nullptr;
  MVS_TRY(mvs_nn_index_create(ctx, &idx));
  std::vector<int32_t> nn((size_t)n_query);
  int st = mvs_nn_index_build_host_f64(idx, template_xyz, n_template);
  if (st == MVS_OK) st = mvs_nn_knn_host_f64(idx, query_xyz, n_query, 1, nn.data(), nullptr);
  mvs_nn_index_destroy(idx);
  MVS_TRY(st);
  for (int64_t i = 0; i < n_query; ++i) out_labels[i] = template_labels[nn[i]];
  return MVS_OK;
}

}  // extern "C"
