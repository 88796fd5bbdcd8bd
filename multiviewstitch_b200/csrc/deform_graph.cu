// deform_graph.cu — deformation-graph construction on the device:
//   node sampling (Deformation.cpp:63-106 greedy rule; farthest-point sampling for a prescribed
//   node count), per-vertex k-nearest-node weights, vertex reordering by node tuple, node graph
//   (Deformation.cpp:108-153), block-sparse pattern of the normal equations, and the node
//   correspondence + smoothing stage (Deformation.cpp:266-381).
#include <cooperative_groups.h>
#include <cub/device/device_radix_sort.cuh>

#include <stdlib.h>

#include <algorithm>
#include <cmath>

#include "comm.cuh"
#include "deform.cuh"
#include "nn_pyramid.cuh"

namespace mvs {

int bbox_host(mvs_ctx* ctx, const float4* pts, int64_t n, DevBuf<float>& red, float mn[3],
              float mx[3]);

// ------------------------------------------------------------------ small gathers
__global__ void k_gather_f4(const float4* __restrict__ src, const int* __restrict__ ids, int64_t m,
                            float4* __restrict__ dst) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < m) dst[i] = src[ids[i]];
}

// ------------------------------------------------------------------ farthest point sampling
struct FpsState {
  int cur;
};

__global__ void __launch_bounds__(256)
    k_fps_step(const float4* __restrict__ pts, int64_t n, float* __restrict__ mind,
               int* __restrict__ cur_io, int* __restrict__ out_ids, int step,
               float* __restrict__ pval, int* __restrict__ pidx, unsigned* __restrict__ ticket) {
  __shared__ float sv[8];
  __shared__ int si[8];
  __shared__ bool is_last;
  int cur = *cur_io;
  float4 c = __ldg(pts + cur);
  float best = -1.0f;
  int bi = 0x7fffffff;
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float4 p = __ldg(pts + i);
    float d = dist2_rn(p.x, p.y, p.z, c.x, c.y, c.z);
    float md = mind[i];
    if (d < md) {
      md = d;
      mind[i] = d;
    }
    if (md > best || (md == best && (int)i < bi)) {
      best = md;
      bi = (int)i;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, best, o);
    int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) {
      best = ov;
      bi = oi;
    }
  }
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    sv[warp] = best;
    si[warp] = bi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
      if (sv[w] > best || (sv[w] == best && si[w] < bi)) {
        best = sv[w];
        bi = si[w];
      }
    pval[blockIdx.x] = best;
    pidx[blockIdx.x] = bi;
    __threadfence();
    unsigned t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  // the last CTA picks the winner among the per-CTA candidates: every thread takes a strided share
  // (independent loads), then the same warp / CTA argmax tree as above — (value desc, index asc)
  __threadfence();
  best = -1.0f;
  bi = 0x7fffffff;
  for (unsigned k = threadIdx.x; k < gridDim.x; k += blockDim.x) {
    float v = __ldcg(pval + k);
    int ix = __ldcg(pidx + k);
    if (v > best || (v == best && ix < bi)) {
      best = v;
      bi = ix;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, best, o);
    int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) {
      best = ov;
      bi = oi;
    }
  }
  __syncthreads();
  if (lane == 0) {
    sv[warp] = best;
    si[warp] = bi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
      if (sv[w] > best || (sv[w] == best && si[w] < bi)) {
        best = sv[w];
        bi = si[w];
      }
    out_ids[step] = cur;
    *cur_io = bi;
    *ticket = 0u;
  }
}

// Farthest point sampling of a small point set (n <= 8 x kFpsCap: the C1/C4 view meshes) in ONE
// launch of one thread-block cluster.  The per-step launch above costs ~7 us per node whatever n is;
// here every thread keeps its PPT points AND their running minimum distances in registers (the loop
// over them is fully unrolled: 12 instructions per point, no shared-memory traffic), the argmax of a
// warp is two REDUX instructions (maximum of the float bits — distances are non-negative — then the
// minimum index among the lanes that hold it), the CTAs publish one candidate each (value, index,
// position) in a double-buffered slot, meet at the cluster barrier, and every warp reads the 8
// candidates through distributed shared memory and picks the winner — one cluster barrier per node,
// no global traffic.  Same result as the oracle (orc_sample_fps): FMA-free float32 distances,
// farthest point with the lowest index on ties.
namespace cg = cooperative_groups;
constexpr int kFpsCluster = 8;
constexpr int kFpsThreads = 512;
constexpr int kFpsMaxPpt = 24;                        // points per thread
constexpr int kFpsCap = kFpsThreads * kFpsMaxPpt;     // points per CTA (12 B each in shared memory)

struct FpsCand {
  unsigned v;  // float bits of the running minimum distance (>= 0: ordered like unsigned)
  int i;
  float x, y, z;
};

// (max value, lowest index among the lanes holding it) over the lanes of `mask`
__device__ __forceinline__ void warp_argmax(unsigned mask, unsigned& vb, int& bi) {
  unsigned vmax = __reduce_max_sync(mask, vb);
  int cand = vb == vmax ? bi : 0x7fffffff;
  bi = __reduce_min_sync(mask, cand);
  vb = vmax;
}

template <int PPT>
__global__ void __launch_bounds__(kFpsThreads, 1)
    k_fps_cluster(const float4* __restrict__ pts, int n, int n_nodes, int first,
                  int* __restrict__ out_ids) {
  extern __shared__ float fps_smem[];
  __shared__ unsigned sv[kFpsThreads / 32];
  __shared__ int si[kFpsThreads / 32];
  __shared__ FpsCand cand[2];
  const unsigned full = 0xffffffffu;
  cg::cluster_group cl = cg::this_cluster();
  const int rank = (int)cl.block_rank();
  const int nblk = (int)cl.num_blocks();  // cluster size is a launch attribute (<= kFpsCluster)
  const int per = (n + nblk - 1) / nblk;
  const int b0 = rank * per;
  const int cnt = max(0, min(per, n - b0));
  // the CTA's coordinates once more in shared memory: the winner's position is read from there
  float* xs = fps_smem;
  float* ys = xs + per;
  float* zs = ys + per;
  float x[PPT], y[PPT], z[PPT], md[PPT];
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    const int k = threadIdx.x + j * kFpsThreads;
    if (k < cnt) {
      float4 p = __ldg(pts + b0 + k);
      x[j] = p.x; y[j] = p.y; z[j] = p.z;
      md[j] = INFINITY;
      xs[k] = p.x; ys[k] = p.y; zs[k] = p.z;
    } else {
      x[j] = y[j] = z[j] = 0.f;
      md[j] = -1.0f;  // never updated (d >= 0), never selected (best starts at -1, strict >)
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 c = __ldg(pts + first);
  int cur = first;
  __syncthreads();
  for (int s = 0; s < n_nodes; ++s) {
    if (rank == 0 && threadIdx.x == 0) out_ids[s] = cur;
    float best = -1.0f;
    int bi = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < PPT; ++j) {  // ascending index: strict > keeps the lowest index of a tie
      float d = dist2_rn(x[j], y[j], z[j], c.x, c.y, c.z);
      float m = fminf(md[j], d);
      md[j] = m;
      if (m > best) {
        best = m;
        bi = b0 + threadIdx.x + j * kFpsThreads;
      }
    }
    unsigned vb = best < 0.f ? 0u : __float_as_uint(best);
    warp_argmax(full, vb, bi);
    if (lane == 0) {
      sv[warp] = vb;
      si[warp] = bi;
    }
    __syncthreads();
    const int slot = s & 1;
    if (warp == 0) {
      vb = lane < kFpsThreads / 32 ? sv[lane] : 0u;
      bi = lane < kFpsThreads / 32 ? si[lane] : 0x7fffffff;
      warp_argmax(full, vb, bi);
      if (lane == 0) {
        FpsCand cd;
        cd.v = vb;
        cd.i = bi;
        const bool has = bi != 0x7fffffff;
        cd.x = has ? xs[bi - b0] : 0.f;
        cd.y = has ? ys[bi - b0] : 0.f;
        cd.z = has ? zs[bi - b0] : 0.f;
        cand[slot] = cd;
      }
    }
    // candidates of all CTAs are published; a slot is rewritten two steps later, after every CTA has
    // passed the barrier of the step in between (so nobody still reads it)
    cl.sync();
    FpsCand w;
    w.v = 0u;
    w.i = 0x7fffffff;
    w.x = w.y = w.z = 0.f;
    if (lane < nblk) w = *cl.map_shared_rank(&cand[slot], lane);
    unsigned wv = w.v;
    int wi = w.i;
    warp_argmax(full, wv, wi);
    // the lane that holds the winner hands out its position (a CTA without points has index INT_MAX
    // and loses every tie; the winner index is unique)
    const unsigned holders = __ballot_sync(full, lane < nblk && w.v == wv && w.i == wi);
    const int src = holders ? __ffs((int)holders) - 1 : 0;
    cur = wi;
    c.x = __shfl_sync(full, w.x, src);
    c.y = __shfl_sync(full, w.y, src);
    c.z = __shfl_sync(full, w.z, src);
  }
  cl.sync();  // no CTA leaves while a peer may still read its candidate slots
}

template <int PPT>
static int launch_fps_cluster(mvs_ctx* ctx, const float4* pts, int n, int n_nodes, int first, int* out,
                              int nblk) {
  int per = (n + nblk - 1) / nblk;
  size_t smem = (size_t)per * 12;
  auto kern = k_fps_cluster<PPT>;
  MVS_TRY(ensure_dyn_smem((const void*)kern, ctx->device, kFpsCap * 12));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(nblk);
  cfg.blockDim = dim3(kFpsThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = nblk;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  ProfScope ps(ctx, "fps_cluster");
  MVS_CUDA(cudaLaunchKernelEx(&cfg, kern, pts, n, n_nodes, first, out));
  return MVS_OK;
}

// ------------------------------------------------------------------ per-vertex node weights
// w = (1 - d/d_max)^2 over the k nearest nodes, d_max = distance to the
// (k+1)-th, normalised; FMA-free float32 in the oracle's order (oracle/deform.cpp
// orc_node_weights) so the weights are bit-reproducible.
__global__ void k_node_weights(const int* __restrict__ idx5, const float* __restrict__ d25,
                               int64_t n, int4* __restrict__ vidx, float4* __restrict__ vw) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int id[5];
  float dd[5];
#pragma unroll
  for (int a = 0; a < 5; ++a) {
    id[a] = idx5[i * 5 + a];
    dd[a] = d25[i * 5 + a];
  }
  int kk = 0;
#pragma unroll
  for (int a = 0; a < 4; ++a)
    if (id[a] >= 0) kk = a + 1;
  float dmax = id[4] >= 0 ? __fsqrt_rn(dd[4]) : 0.0f;
  float w[4];
  float sum = 0.0f;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    if (a < kk && dmax > 0.0f) {
      float r = __fdiv_rn(__fsqrt_rn(dd[a]), dmax);
      float u = __fsub_rn(1.0f, r);
      w[a] = __fmul_rn(u, u);
    } else {
      w[a] = 0.0f;
    }
    sum = __fadd_rn(sum, w[a]);
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    if (sum > 0.0f)
      w[a] = __fdiv_rn(w[a], sum);
    else
      w[a] = (a < kk) ? __fdiv_rn(1.0f, (float)kk) : 0.0f;
  }
  vidx[i] = make_int4(id[0], id[1], id[2], id[3]);
  vw[i] = make_float4(w[0], w[1], w[2], w[3]);
}

__device__ __forceinline__ void cswap(int& a, int& b, float& wa, float& wb) {
  unsigned ua = a < 0 ? 0x7fffffffu : (unsigned)a, ub = b < 0 ? 0x7fffffffu : (unsigned)b;
  if (ub < ua) {
    int t = a; a = b; b = t;
    float tw = wa; wa = wb; wb = tw;
  }
}

__device__ __forceinline__ unsigned spread10(unsigned v) {  // 10 bits -> every third bit
  v &= 0x3ffu;
  v = (v | (v << 16)) & 0x030000ffu;
  v = (v | (v << 8)) & 0x0300f00fu;
  v = (v | (v << 4)) & 0x030c30c3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}

// sort key: the Morton code of the centroid of the vertex's four nodes in the high word — a function
// of the node SET, so vertices with identical sets are adjacent (tuple runs), consecutive runs are
// neighbours on the surface (coherent walks, shared accumulators in the assembly groups) and a
// vertex shard touches a compact node subset; a hash of the ascending tuple in the low word keeps
// different sets with the same code apart.
__global__ void k_tuple_keys(const int4* __restrict__ vidx, const float4* __restrict__ g,
                             float3 bmin, float3 inv_ext, int64_t n,
                             unsigned long long* __restrict__ keys, int* __restrict__ vals) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int4 id = vidx[i];
  float d0 = 0, d1 = 0, d2 = 0, d3 = 0;
  cswap(id.x, id.y, d0, d1);
  cswap(id.z, id.w, d2, d3);
  cswap(id.x, id.z, d0, d2);
  cswap(id.y, id.w, d1, d3);
  cswap(id.y, id.z, d1, d2);
  unsigned h = 2166136261u;
  h = (h ^ (unsigned)id.x) * 16777619u;
  h = (h ^ (unsigned)id.y) * 16777619u;
  h = (h ^ (unsigned)id.z) * 16777619u;
  h = (h ^ (unsigned)id.w) * 16777619u;
  h ^= h >> 15;
  // centroid of the valid nodes, summed in ascending-id order (identical for every vertex of the set)
  int ids[4] = {id.x, id.y, id.z, id.w};
  float cx = 0.f, cy = 0.f, cz = 0.f, cnt = 0.f;
  for (int a = 0; a < 4; ++a)
    if (ids[a] >= 0) {
      float4 p = g[ids[a]];
      cx += p.x; cy += p.y; cz += p.z;
      cnt += 1.f;
    }
  float ic = cnt > 0.f ? 1.0f / cnt : 0.f;
  unsigned mx = (unsigned)fminf(fmaxf((cx * ic - bmin.x) * inv_ext.x, 0.f), 1023.f);
  unsigned my = (unsigned)fminf(fmaxf((cy * ic - bmin.y) * inv_ext.y, 0.f), 1023.f);
  unsigned mz = (unsigned)fminf(fmaxf((cz * ic - bmin.z) * inv_ext.z, 0.f), 1023.f);
  unsigned morton = spread10(mx) | (spread10(my) << 1) | (spread10(mz) << 2);
  keys[i] = ((unsigned long long)morton << 32) | h;
  vals[i] = (int)i;
}

__global__ void k_permute_vertices(const int* __restrict__ perm, int64_t n,
                                   const float4* __restrict__ v0, const float4* __restrict__ n0,
                                   const int4* __restrict__ idx_knn,
                                   const float4* __restrict__ w_knn, float4* __restrict__ v,
                                   float4* __restrict__ nrm, int4* __restrict__ vidx,
                                   float4* __restrict__ vw, int* __restrict__ inv_perm) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int o = perm[i];
  inv_perm[o] = (int)i;
  float4 p = v0[o];
  p.w = __int_as_float(o);
  v[i] = p;
  nrm[i] = n0[o];
  int4 id = idx_knn[o];
  float4 w = w_knn[o];
  cswap(id.x, id.y, w.x, w.y);
  cswap(id.z, id.w, w.z, w.w);
  cswap(id.x, id.z, w.x, w.z);
  cswap(id.y, id.w, w.y, w.w);
  cswap(id.y, id.z, w.y, w.z);
  vidx[i] = id;
  vw[i] = w;
}

// ------------------------------------------------------------------ block-sparse pattern
__device__ __forceinline__ void row_insert(int* __restrict__ tab, int cap, int row, int colv,
                                           int* __restrict__ overflow) {
  int* r = tab + (size_t)row * cap;
  unsigned h = ((unsigned)colv * 2654435761u) % (unsigned)cap;
  for (int probe = 0; probe < cap; ++probe) {
    int cur = r[h];
    if (cur == colv) return;
    if (cur == -1) {
      int old = atomicCAS(r + h, -1, colv);
      if (old == -1 || old == colv) return;
    }
    h = (h + 1) % (unsigned)cap;
  }
  *overflow = 1;
}

__global__ void k_pattern_vertices(const int4* __restrict__ vidx, int64_t n, int* __restrict__ tab,
                                   int cap, int* __restrict__ overflow) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int4 id = vidx[i];
  if (i > 0) {
    int4 pr = vidx[i - 1];
    if (pr.x == id.x && pr.y == id.y && pr.z == id.z && pr.w == id.w) return;  // same tuple
  }
  int a[4] = {id.x, id.y, id.z, id.w};
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (a[p] >= 0 && a[q] >= 0) row_insert(tab, cap, a[p], a[q], overflow);
}

__global__ void k_pattern_graph(const int* __restrict__ graph, int64_t m, int gk1,
                                int* __restrict__ tab, int cap, int* __restrict__ overflow) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= m * gk1) return;
  int j = (int)(t / gk1);
  int l = graph[t];
  if (t % gk1 == 0) row_insert(tab, cap, j, j, overflow);
  if (l < 0 || l == j) return;
  row_insert(tab, cap, j, l, overflow);
  row_insert(tab, cap, l, j, overflow);
}

// per row: compact non-empty slots to the front, sorted ascending; counts[j] = row length
__global__ void k_pattern_compact(int* __restrict__ tab, int cap, int64_t m,
                                  int* __restrict__ counts) {
  int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= m) return;
  int* r = tab + (size_t)j * cap;
  int cnt = 0;
  for (int s = 0; s < cap; ++s) {
    int c = r[s];
    if (c < 0) continue;
    r[s] = -1;
    // insertion into the sorted prefix r[0..cnt)
    int p = cnt;
    while (p > 0 && r[p - 1] > c) {
      r[p] = r[p - 1];
      --p;
    }
    r[p] = c;
    ++cnt;
  }
  counts[j] = cnt;
}

__global__ void k_pattern_gather(const int* __restrict__ tab, int cap, int64_t m,
                                 const int* __restrict__ row_ptr, int* __restrict__ col) {
  int64_t j = blockIdx.x;
  if (j >= m) return;
  int s = row_ptr[j], e = row_ptr[j + 1];
  for (int t = threadIdx.x; t < e - s; t += blockDim.x) col[s + t] = tab[(size_t)j * cap + t];
}

// ------------------------------------------------------------------ tuple runs
__global__ void k_run_heads(const int4* __restrict__ vidx, int64_t n, int* __restrict__ flag) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int h = 1;
  if (i > 0) {
    int4 a = vidx[i], b = vidx[i - 1];
    h = (a.x != b.x || a.y != b.y || a.z != b.z || a.w != b.w) ? 1 : 0;
  }
  flag[i] = h;
}

__device__ __forceinline__ int find_slot_dev(const int* __restrict__ row_ptr,
                                             const int* __restrict__ col, int j, int l) {
  int lo = row_ptr[j], hi = row_ptr[j + 1] - 1;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (col[mid] < l) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// scan[i] = run id of vertex i's run when i is a head; every vertex records its run id in v.w,
// heads fill the run's slot table (16 pattern slots (a,b) row-major, then the 4 node ids)
__global__ void k_run_table(const int4* __restrict__ vidx, const int* __restrict__ flag_scan,
                            int64_t n, const int* __restrict__ row_ptr,
                            const int* __restrict__ col, const float4* __restrict__ g,
                            float4* __restrict__ v, int* __restrict__ run_slots,
                            int* __restrict__ run_start, float4* __restrict__ run_g) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int4 id = vidx[i];
  bool head = true;
  if (i > 0) {
    int4 b = vidx[i - 1];
    head = id.x != b.x || id.y != b.y || id.z != b.z || id.w != b.w;
  }
  // exclusive scan of the head flags: a head's value is its run id, a non-head's is run id + 1
  int run = flag_scan[i] - (head ? 0 : 1);
  v[i].w = __int_as_float(run);
  if (!head) return;
  run_start[run] = (int)i;
  int js[4] = {id.x, id.y, id.z, id.w};
  int* out = run_slots + (size_t)run * 20;
  for (int a = 0; a < 4; ++a) {
    for (int b = 0; b < 4; ++b)
      out[4 * a + b] = (js[a] >= 0 && js[b] >= 0) ? find_slot_dev(row_ptr, col, js[a], js[b]) : -1;
    out[16 + a] = js[a];
    float4 gj = js[a] >= 0 ? g[js[a]] : make_float4(0.f, 0.f, 0.f, 0.f);
    gj.w = js[a] >= 0 ? 1.0f : 0.0f;
    run_g[(size_t)run * 4 + a] = gj;
  }
}

// tile t of a run: t < 10 the block tile (a, b), a <= b; t = 10..13 the gradient tile of node t - 10
__constant__ int c_tile_a[10] = {0, 1, 2, 3, 0, 0, 0, 1, 1, 2};
__constant__ int c_tile_b[10] = {0, 1, 2, 3, 1, 2, 3, 2, 3, 3};

// One thread per assembly group: the distinct destinations (first sys64 element) of the group's
// tiles and, per run, the group-local accumulator index of each tile (255 = no destination).
__global__ void k_group_table(const int* __restrict__ run_slots, int64_t nruns,
                              const int* __restrict__ uidx, int64_t nup, int64_t ngroups,
                              int* __restrict__ grp_n, int* __restrict__ grp_dst,
                              unsigned char* __restrict__ run_local) {
  int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (q >= ngroups) return;
  int* dst = grp_dst + (size_t)q * kGrpMax;
  int cnt = 0;
  int64_t r1 = min(nruns, (q + 1) * (int64_t)kGrpRuns);
  for (int64_t r = q * kGrpRuns; r < r1; ++r) {
    const int* rs = run_slots + (size_t)r * 20;
    for (int t = 0; t < 14; ++t) {
      int d;
      if (t < 10) {
        int sl = rs[4 * c_tile_a[t] + c_tile_b[t]];
        d = sl < 0 ? -1 : uidx[sl] * 16;
      } else {
        int j = rs[16 + t - 10];
        d = j < 0 ? -1 : (int)(nup * 16 + (int64_t)j * 12);
      }
      int loc = 255;
      if (d >= 0) {
        for (int i = 0; i < cnt; ++i)
          if (dst[i] == d) {
            loc = i;
            break;
          }
        if (loc == 255) {
          dst[cnt] = d;
          loc = cnt++;
        }
      }
      run_local[(size_t)r * 16 + t] = (unsigned char)loc;
    }
  }
  grp_n[q] = cnt;
}

// tslot[e] = slot of block (l, j) for the entry e = (j, l); flag[e] = 1 for the upper triangle
// (l >= j), whose exclusive scan numbers the blocks kept in sys64
__global__ void k_transposed_slots(const int* __restrict__ row_ptr, const int* __restrict__ col,
                                   int64_t m, int* __restrict__ tslot, int* __restrict__ upper) {
  int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= m) return;
  for (int t = row_ptr[j]; t < row_ptr[j + 1]; ++t) {
    tslot[t] = find_slot_dev(row_ptr, col, col[t], (int)j);
    upper[t] = col[t] >= j ? 1 : 0;
  }
}

// uidx[e]: rank of an upper block among the upper blocks; a lower block points at its transpose
__global__ void k_upper_index(const int* __restrict__ tslot, const int* __restrict__ scan,
                              int64_t nnz, int* __restrict__ uidx) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= nnz) return;
  int te = tslot[e];
  uidx[e] = scan[te < e ? te : e];
}

// ------------------------------------------------------------------ build_graph
int deform_build_graph(mvs_deform* h, int node_k, int graph_k) {
  mvs_ctx* ctx = h->ctx;
  MVS_TRY(ctx_ensure_device(ctx));
  if (node_k != 4) {
    set_error("build_graph: node_k must be 4 (the kernels are specialised for k = 4)");
    return MVS_ERR_INVALID_ARG;
  }
  if (graph_k < 1 || graph_k > 31) {
    set_error("build_graph: graph_k must be in 1..31");
    return MVS_ERR_INVALID_ARG;
  }
  if (h->m <= 0) {
    set_error("build_graph: no nodes (call uniform_sampling / sample_fps / set_nodes first)");
    return MVS_ERR_STATE;
  }
  int64_t n = h->n, m = h->m;
  cudaStream_t st = ctx->stream;
  h->node_k = node_k;
  h->graph_k = graph_k;
  h->graph_built = false;
  // 1. k+1 nearest nodes of every vertex, weights (kNN order, original vertex order)
  DevBuf<int> idx5;
  DevBuf<float> d25;
  MVS_TRY(idx5.reserve((size_t)n * 5));
  MVS_TRY(d25.reserve((size_t)n * 5));
  MVS_TRY(nn_knn(h->node_index, h->v0.p, n, 5, idx5.p, d25.p));
  MVS_TRY(h->vidx_knn.reserve((size_t)n));
  MVS_TRY(h->vw_knn.reserve((size_t)n));
  {
    ProfScope ps(ctx, "node_weights");
    k_node_weights<<<div_up(n, 256), 256, 0, st>>>(idx5.p, d25.p, n, h->vidx_knn.p, h->vw_knn.p);
  }
  MVS_CUDA(cudaGetLastError());
  // 2. sort vertices by (primary node, tuple hash)
  DevBuf<unsigned long long> keys, keys_out;
  DevBuf<int> vals;
  MVS_TRY(keys.reserve((size_t)n));
  MVS_TRY(keys_out.reserve((size_t)n));
  MVS_TRY(vals.reserve((size_t)n));
  MVS_TRY(h->perm.reserve((size_t)n));
  MVS_TRY(h->inv_perm.reserve((size_t)n));
  {
    float gmn[3], gmx[3];
    MVS_TRY(bbox_host(ctx, h->g.p, m, h->node_index->red, gmn, gmx));
    float3 bmin = make_float3(gmn[0], gmn[1], gmn[2]);
    float3 inv_ext;
    inv_ext.x = gmx[0] > gmn[0] ? 1023.999f / (gmx[0] - gmn[0]) : 0.f;
    inv_ext.y = gmx[1] > gmn[1] ? 1023.999f / (gmx[1] - gmn[1]) : 0.f;
    inv_ext.z = gmx[2] > gmn[2] ? 1023.999f / (gmx[2] - gmn[2]) : 0.f;
    ProfScope ps(ctx, "tuple_keys");
    k_tuple_keys<<<div_up(n, 256), 256, 0, st>>>(h->vidx_knn.p, h->g.p, bmin, inv_ext, n, keys.p,
                                                 vals.p);
  }
  {
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys.p, keys_out.p, vals.p, h->perm.p,
                                    (int)n, 0, 64, st);
    DevBuf<char> tmp;
    MVS_TRY(tmp.reserve(tmp_bytes));
    ProfScope ps(ctx, "tuple_sort", 8);
    MVS_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, keys.p, keys_out.p, vals.p,
                                             h->perm.p, (int)n, 0, 64, st));
    MVS_CUDA(cudaStreamSynchronize(st));
  }
  MVS_TRY(h->v.reserve((size_t)n));
  MVS_TRY(h->nrm.reserve((size_t)n));
  MVS_TRY(h->vidx.reserve((size_t)n));
  MVS_TRY(h->vw.reserve((size_t)n));
  {
    ProfScope ps(ctx, "permute_vertices");
    k_permute_vertices<<<div_up(n, 256), 256, 0, st>>>(h->perm.p, n, h->v0.p, h->nrm0.p,
                                                       h->vidx_knn.p, h->vw_knn.p, h->v.p,
                                                       h->nrm.p, h->vidx.p, h->vw.p,
                                                       h->inv_perm.p);
  }
  MVS_CUDA(cudaGetLastError());
  // 3. node graph: (K+1) nearest nodes including self (Deformation.cpp:108-153)
  int gk1 = graph_k + 1;
  MVS_TRY(h->graph.reserve((size_t)m * gk1));
  MVS_TRY(nn_knn(h->node_index, h->g.p, m, gk1, h->graph.p, nullptr));
  // 4. block pattern: per-row hash sets, grown on overflow
  DevBuf<int> tab, counts, flag;
  MVS_TRY(flag.reserve(1));
  MVS_TRY(h->row_ptr.reserve((size_t)m + 2));
  int cap = 64;
  for (;;) {
    MVS_TRY(tab.reserve((size_t)m * cap));
    MVS_CUDA(cudaMemsetAsync(tab.p, 0xff, sizeof(int) * (size_t)m * cap, st));
    MVS_CUDA(cudaMemsetAsync(flag.p, 0, sizeof(int), st));
    {
      ProfScope ps(ctx, "pattern_insert", 2);
      k_pattern_vertices<<<div_up(n, 256), 256, 0, st>>>(h->vidx.p, n, tab.p, cap, flag.p);
      k_pattern_graph<<<div_up(m * gk1, 256), 256, 0, st>>>(h->graph.p, m, gk1, tab.p, cap, flag.p);
    }
    int of = 0;
    MVS_TRY(d2h_sync(ctx, &of, flag.p, sizeof(int)));
    if (!of) break;
    cap *= 2;
    if (cap > 4096) {
      set_error("build_graph: node adjacency exceeds 4096 blocks per row");
      return MVS_ERR_CAPACITY;
    }
  }
  {
    ProfScope ps(ctx, "pattern_compact");
    k_pattern_compact<<<div_up(m, 128), 128, 0, st>>>(tab.p, cap, m, h->row_ptr.p);
  }
  MVS_CUDA(cudaMemsetAsync(h->row_ptr.p + m, 0, sizeof(int) * 2, st));
  MVS_TRY(exclusive_scan_inplace(ctx, h->row_ptr.p, m, h->node_index->scan_tmp));
  int nnz = 0;
  MVS_TRY(d2h_sync(ctx, &nnz, h->row_ptr.p + m, sizeof(int)));
  h->nnz = nnz;
  h->h_row_ptr.resize((size_t)m + 1);
  MVS_TRY(d2h_sync(ctx, h->h_row_ptr.data(), h->row_ptr.p, sizeof(int) * (m + 1)));
  h->pcg_plan_valid = false;
  MVS_TRY(h->col.reserve((size_t)nnz));
  {
    ProfScope ps(ctx, "pattern_gather");
    k_pattern_gather<<<(unsigned)m, 64, 0, st>>>(tab.p, cap, m, h->row_ptr.p, h->col.p);
  }
  MVS_CUDA(cudaGetLastError());
  // 5. runs of equal node tuples + their slot tables
  {
    DevBuf<int> rflag;
    MVS_TRY(rflag.reserve((size_t)n + 2));
    {
      ProfScope ps(ctx, "run_heads");
      k_run_heads<<<div_up(n, 256), 256, 0, st>>>(h->vidx.p, n, rflag.p);
    }
    MVS_CUDA(cudaMemsetAsync(rflag.p + n, 0, sizeof(int) * 2, st));
    MVS_TRY(exclusive_scan_inplace(ctx, rflag.p, n, h->node_index->scan_tmp));
    int nruns = 0;
    MVS_TRY(d2h_sync(ctx, &nruns, rflag.p + n, sizeof(int)));
    h->nruns = nruns;
    if ((int64_t)nnz * 16 + m * 12 >= (int64_t)1 << 31) {
      set_error("build_graph: normal-equation system exceeds 2^31 elements");
      return MVS_ERR_CAPACITY;
    }
    h->ngroups = (nruns + kGrpRuns - 1) / kGrpRuns;
    MVS_TRY(h->run_slots.reserve((size_t)nruns * 20));
    MVS_TRY(h->run_start.reserve((size_t)nruns + 1));
    MVS_TRY(h->run_g.reserve((size_t)nruns * 4));
    MVS_TRY(h->run_local.reserve((size_t)nruns * 16));
    MVS_TRY(h->grp_n.reserve((size_t)h->ngroups));
    MVS_TRY(h->grp_dst.reserve((size_t)h->ngroups * kGrpMax));
    MVS_TRY(h->tslot.reserve((size_t)nnz));
    MVS_TRY(h->uidx.reserve((size_t)nnz));
    {
      DevBuf<int> upper;
      MVS_TRY(upper.reserve((size_t)nnz + 2));
      {
        ProfScope ps(ctx, "transposed_slots");
        k_transposed_slots<<<div_up(m, 128), 128, 0, st>>>(h->row_ptr.p, h->col.p, m, h->tslot.p,
                                                           upper.p);
      }
      MVS_CUDA(cudaMemsetAsync(upper.p + nnz, 0, sizeof(int) * 2, st));
      MVS_TRY(exclusive_scan_inplace(ctx, upper.p, nnz, h->node_index->scan_tmp));
      int nup = 0;
      MVS_TRY(d2h_sync(ctx, &nup, upper.p + nnz, sizeof(int)));
      h->nup = nup;
      ProfScope ps(ctx, "upper_index");
      k_upper_index<<<div_up(nnz, 256), 256, 0, st>>>(h->tslot.p, upper.p, nnz, h->uidx.p);
      MVS_CUDA(cudaStreamSynchronize(st));  // `upper` is released in stream order
    }
    {
      ProfScope ps(ctx, "run_table", 2);
      k_run_table<<<div_up(n, 256), 256, 0, st>>>(h->vidx.p, rflag.p, n, h->row_ptr.p, h->col.p,
                                                  h->g.p, h->v.p, h->run_slots.p, h->run_start.p,
                                                  h->run_g.p);
      k_group_table<<<div_up(h->ngroups, 64), 64, 0, st>>>(h->run_slots.p, nruns, h->uidx.p, h->nup,
                                                           h->ngroups, h->grp_n.p, h->grp_dst.p,
                                                           h->run_local.p);
    }
    int n32 = (int)n;
    MVS_CUDA(cudaMemcpyAsync(h->run_start.p + nruns, &n32, sizeof(int), cudaMemcpyHostToDevice, st));
    MVS_CUDA(cudaGetLastError());
    MVS_CUDA(cudaStreamSynchronize(st));
    // host copy of the first vertex of every assembly group (+ n): vertex shards are snapped to
    // group boundaries, the unit of float32 accumulation, so sharded sums carry the unsharded bits
    {
      std::vector<int> rs((size_t)nruns + 1);
      MVS_TRY(d2h_sync(ctx, rs.data(), h->run_start.p, sizeof(int) * (nruns + 1)));
      h->h_grp_start.resize((size_t)h->ngroups + 1);
      for (int64_t q = 0; q < h->ngroups; ++q) h->h_grp_start[q] = rs[(size_t)q * kGrpRuns];
      h->h_grp_start[h->ngroups] = (int)n;
    }
    float mn[3], mx[3];
    MVS_TRY(bbox_host(ctx, h->v0.p, n, h->node_index->red, mn, mx));
    double d2 = 0;
    for (int a = 0; a < 3; ++a) d2 += ((double)mx[a] - mn[a]) * ((double)mx[a] - mn[a]);
    h->diag2 = d2 > 0 ? d2 : 1.0;
  }
  // 6. solver buffers
  if (h->peer) {  // a rebuilt graph gets a fresh exchange block (the peers rebuild theirs alike)
    h->sys.release();
    h->sys64.release();
    MVS_CUDA(cudaStreamSynchronize(st));
    peer_link_free(h->peer);
    h->peer = nullptr;
  }
  if (ctx->comm && ctx->nranks > 1 && ctx->nranks <= kPeerMaxRanks && !getenv("MVS_NO_PEER")) {
    // vertex-sharded solve: the two systems live in a block the other ranks can map (comm.cuh)
    h->sys.release();
    h->sys64.release();
    MVS_TRY(peer_block_alloc(ctx, h->sys64_count() + 2, h->sys_floats(), &h->peer));
    h->sys64.adopt(reinterpret_cast<long long*>(h->peer->block), h->sys64_count() + 2);
    h->sys.adopt(reinterpret_cast<float*>((char*)h->peer->block + h->peer->off_sysf), h->sys_floats());
  } else {
    MVS_TRY(h->sys.reserve(h->sys_floats()));
    MVS_TRY(h->sys64.reserve(h->sys64_count() + 2));  // + the assembly group counter
  }
  MVS_TRY(h->Minv.reserve((size_t)m * 144));
  MVS_TRY(h->x.reserve((size_t)m * 12));
  MVS_TRY(h->pcg.reserve((size_t)m * 12 * 6));
  MVS_TRY(h->scal.reserve(S_SIZE));
  MVS_TRY(h->partials.reserve((size_t)8 * ctx->num_sms * 4 + 64));
  MVS_TRY(h->ticket.reserve(1));
  MVS_CUDA(cudaMemsetAsync(h->ticket.p, 0, sizeof(unsigned), st));
  MVS_CUDA(cudaMemsetAsync(h->scal.p, 0, sizeof(double) * S_SIZE, st));
  MVS_TRY(h->resid.reserve((size_t)n));
  MVS_TRY(h->corr_prev.reserve((size_t)n));
  MVS_CUDA(cudaMemsetAsync(h->corr_prev.p, 0xff, sizeof(int) * n, st));
  h->corr_seeded = false;
  MVS_TRY(h->targets.reserve((size_t)n));
  MVS_TRY(h->cur0.reserve((size_t)n));
  MVS_TRY(h->curn0.reserve((size_t)n));
  MVS_TRY(gn_reset_state(h));
  MVS_CUDA(cudaStreamSynchronize(st));
  h->graph_built = true;
  h->cur_valid = false;
  return MVS_OK;
}

// ------------------------------------------------------------------ node correspondences
// Deformation/Deformation.cpp:266-357 for one node per thread, in FMA-free float64 so the
// selection keys are bit-identical to the host oracle.
struct RadiusCount {
  float qx, qy, qz, r2;
  int count;
  __device__ __forceinline__ float worst2() const { return r2; }
  __device__ __forceinline__ void point(float4 p) {
    if (dist2_rn(p.x, p.y, p.z, qx, qy, qz) <= r2) ++count;
  }
};

struct Top8 {
  double key[8], aplen[8], plen[8], pdist[8];
  int idx[8];
  int cnt;
  __device__ __forceinline__ void init() {
    cnt = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      key[i] = INFINITY;
      aplen[i] = INFINITY;
      idx[i] = 0x7fffffff;
      plen[i] = pdist[i] = 0.0;
    }
  }
  __device__ __forceinline__ static bool less(double k1, double a1, int i1, double k2, double a2,
                                              int i2) {
    if (k1 != k2) return k1 < k2;
    if (a1 != a2) return a1 < a2;
    return i1 < i2;
  }
  __device__ __forceinline__ void insert(double k, double a, int id, double pl, double pd) {
    ++cnt;
    if (!less(k, a, id, key[7], aplen[7], idx[7])) return;
    key[7] = k; aplen[7] = a; idx[7] = id; plen[7] = pl; pdist[7] = pd;
#pragma unroll
    for (int i = 7; i >= 1; --i) {
      if (less(key[i], aplen[i], idx[i], key[i - 1], aplen[i - 1], idx[i - 1])) {
        double t;
        t = key[i]; key[i] = key[i - 1]; key[i - 1] = t;
        t = aplen[i]; aplen[i] = aplen[i - 1]; aplen[i - 1] = t;
        t = plen[i]; plen[i] = plen[i - 1]; plen[i - 1] = t;
        t = pdist[i]; pdist[i] = pdist[i - 1]; pdist[i - 1] = t;
        int ti = idx[i]; idx[i] = idx[i - 1]; idx[i - 1] = ti;
      }
    }
  }
};

struct RadiusTop8 {
  float qx, qy, qz, r2;
  double o[3], nrm[3], nlen;
  const float4* tn;
  Top8 top;
  __device__ __forceinline__ float worst2() const { return r2; }
  __device__ __forceinline__ void point(float4 p) {
    if (!(dist2_rn(p.x, p.y, p.z, qx, qy, qz) <= r2)) return;
    int id = __float_as_int(p.w);
    float4 t = __ldg(tn + id);
    double dn = __dadd_rn(__dadd_rn(__dmul_rn(nrm[0], (double)t.x), __dmul_rn(nrm[1], (double)t.y)),
                          __dmul_rn(nrm[2], (double)t.z));
    if (!(dn > 0.0)) return;  // Deformation.cpp:306-312
    double d0 = __dsub_rn((double)p.x, o[0]), d1 = __dsub_rn((double)p.y, o[1]),
           d2 = __dsub_rn((double)p.z, o[2]);
    double pl = __ddiv_rn(
        __dadd_rn(__dadd_rn(__dmul_rn(d0, nrm[0]), __dmul_rn(d1, nrm[1])), __dmul_rn(d2, nrm[2])),
        nlen);
    double sq = __dsub_rn(
        __dadd_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)), __dmul_rn(d2, d2)),
        __dmul_rn(pl, pl));
    double pd = sq > 0.0 ? __dsqrt_rn(sq) : 0.0;
    top.insert(floor(__ddiv_rn(pd, 1e-6)), fabs(pl), id, pl, pd);
  }
};

__global__ void __launch_bounds__(64)
    k_node_targets(PyrParams tg, const float4* __restrict__ tpts,
                   const float4* __restrict__ tnrm, const float4* __restrict__ g,
                   const float4* __restrict__ gnrm, int64_t m, double proj_len_err,
                   double proj_dist_err, double* __restrict__ controls, int* __restrict__ valid) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= m) return;
  float4 q = g[i];
  float4 nq = gnrm[i];
  controls[3 * i] = (double)q.x;
  controls[3 * i + 1] = (double)q.y;
  controls[3 * i + 2] = (double)q.z;
  valid[i] = 0;
  KnnList<1> nn;
  nn.init(q.x, q.y, q.z);
  pyr_search(tg, q.x, q.y, q.z, nn);
  float radius = __fmul_rn(nn.d2[0], 2.0f);  // Deformation.cpp:288, squared units
  RadiusCount rc{q.x, q.y, q.z, radius, 0};
  pyr_search(tg, q.x, q.y, q.z, rc);
  if (rc.count >= 10000) return;  // Deformation.cpp:286-297,315: all maxResult slots fill -> skipped
  RadiusTop8 rt;
  rt.qx = q.x; rt.qy = q.y; rt.qz = q.z; rt.r2 = radius;
  rt.o[0] = (double)q.x; rt.o[1] = (double)q.y; rt.o[2] = (double)q.z;
  rt.nrm[0] = (double)nq.x; rt.nrm[1] = (double)nq.y; rt.nrm[2] = (double)nq.z;
  rt.nlen = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(rt.nrm[0], rt.nrm[0]),
                                           __dmul_rn(rt.nrm[1], rt.nrm[1])),
                                 __dmul_rn(rt.nrm[2], rt.nrm[2])));
  rt.tn = tnrm;
  rt.top.init();
  pyr_search(tg, q.x, q.y, q.z, rt);
  int nb = min(8, rt.top.cnt);
  if (nb == 0) return;
  double mlen = 0.0, mdist = 0.0, mp[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (j < nb) {
      mlen = __dadd_rn(mlen, rt.top.plen[j]);
      mdist = __dadd_rn(mdist, rt.top.pdist[j]);
      float4 t = __ldg(tpts + rt.top.idx[j]);
      mp[0] = __dadd_rn(mp[0], (double)t.x);
      mp[1] = __dadd_rn(mp[1], (double)t.y);
      mp[2] = __dadd_rn(mp[2], (double)t.z);
    }
  }
  double dnb = (double)nb;
  mlen = __ddiv_rn(mlen, dnb);
  mdist = __ddiv_rn(mdist, dnb);
  for (int a = 0; a < 3; ++a) mp[a] = __ddiv_rn(mp[a], dnb);
  if (mlen >= proj_len_err || mdist >= proj_dist_err) return;  // :350
  double d0 = __dsub_rn(mp[0], rt.o[0]), d1 = __dsub_rn(mp[1], rt.o[1]),
         d2 = __dsub_rn(mp[2], rt.o[2]);
  double dl = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)), __dmul_rn(d2, d2)));
  double cosv = fabs(__ddiv_rn(
      __dadd_rn(__dadd_rn(__dmul_rn(d0, rt.nrm[0]), __dmul_rn(d1, rt.nrm[1])), __dmul_rn(d2, rt.nrm[2])),
      __dmul_rn(dl, rt.nlen)));
  if (cosv < 0.1) return;  // :353
  valid[i] = 1;
  controls[3 * i] = mp[0];
  controls[3 * i + 1] = mp[1];
  controls[3 * i + 2] = mp[2];
}

// Deformation.cpp:362-381: one Jacobi sweep, displacement <- mean of neighbourhood displacements
__global__ void k_smooth_controls(const float4* __restrict__ g, const int* __restrict__ graph,
                                  int gk1, int64_t m, const double* __restrict__ cin,
                                  double* __restrict__ cout) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= m) return;
  double wgt = __ddiv_rn(1.0, (double)gk1);
  double acc[3] = {0.0, 0.0, 0.0};
  for (int e = 0; e < gk1; ++e) {
    int l = graph[i * gk1 + e];
    if (l < 0) continue;
    float4 o = g[l];
    acc[0] = __dadd_rn(acc[0], __dmul_rn(wgt, __dsub_rn(cin[3 * l], (double)o.x)));
    acc[1] = __dadd_rn(acc[1], __dmul_rn(wgt, __dsub_rn(cin[3 * l + 1], (double)o.y)));
    acc[2] = __dadd_rn(acc[2], __dmul_rn(wgt, __dsub_rn(cin[3 * l + 2], (double)o.z)));
  }
  float4 o = g[i];
  cout[3 * i] = __dadd_rn((double)o.x, acc[0]);
  cout[3 * i + 1] = __dadd_rn((double)o.y, acc[1]);
  cout[3 * i + 2] = __dadd_rn((double)o.z, acc[2]);
}

int deform_node_targets(mvs_deform* h, double proj_len_err, double proj_dist_err,
                        int smooth_sweeps, DevBuf<double>& controls, DevBuf<int>& valid) {
  mvs_ctx* ctx = h->ctx;
  if (!h->tgt_index || !h->tgt_index->pyr_built || !h->tn) {
    set_error("node_targets: target points and normals not set (mvs_deform_set_target)");
    return MVS_ERR_STATE;
  }
  if (!h->graph_built) MVS_TRY(deform_build_graph(h, h->node_k, h->graph_k));
  int64_t m = h->m;
  MVS_TRY(controls.reserve((size_t)m * 6));
  MVS_TRY(valid.reserve((size_t)m));
  {
    ProfScope ps(ctx, "node_targets");
    k_node_targets<<<div_up(m, 64), 64, 0, ctx->stream>>>(
        h->tgt_index->pyr, h->tp, h->tn, h->g.p, h->gnrm.p, m, proj_len_err, proj_dist_err,
        controls.p, valid.p);
  }
  double* a = controls.p;
  double* b = controls.p + 3 * m;
  for (int it = 0; it < smooth_sweeps; ++it) {
    ProfScope ps(ctx, "smooth_controls");
    k_smooth_controls<<<div_up(m, 128), 128, 0, ctx->stream>>>(h->g.p, h->graph.p, h->graph_k + 1,
                                                               m, a, b);
    std::swap(a, b);
  }
  if (a != controls.p)
    MVS_CUDA(cudaMemcpyAsync(controls.p, a, sizeof(double) * 3 * m, cudaMemcpyDeviceToDevice,
                             ctx->stream));
  MVS_CUDA(cudaGetLastError());
  return MVS_OK;
}

// ------------------------------------------------------------------ sampling entry points
static int set_nodes_from_host(mvs_deform* h, const std::vector<int32_t>& ids) {
  mvs_ctx* ctx = h->ctx;
  int64_t m = (int64_t)ids.size();
  for (int32_t id : ids)
    if (id < 0 || id >= h->n) {
      set_error("set_nodes: vertex id %d out of range", id);
      return MVS_ERR_INVALID_ARG;
    }
  if (m <= 0) {
    set_error("set_nodes: empty node set");
    return MVS_ERR_DEGENERATE;
  }
  h->node_ids = ids;
  h->m = m;
  MVS_TRY(h->node_ids_d.reserve((size_t)m));
  MVS_TRY(h->g.reserve((size_t)m));
  MVS_TRY(h->gnrm.reserve((size_t)m));
  MVS_CUDA(cudaMemcpyAsync(h->node_ids_d.p, ids.data(), sizeof(int) * m, cudaMemcpyHostToDevice,
                           ctx->stream));
  {
    ProfScope ps(ctx, "gather_nodes", 2);
    k_gather_f4<<<div_up(m, 128), 128, 0, ctx->stream>>>(h->v0.p, h->node_ids_d.p, m, h->g.p);
    k_gather_f4<<<div_up(m, 128), 128, 0, ctx->stream>>>(h->nrm0.p, h->node_ids_d.p, m, h->gnrm.p);
  }
  MVS_CUDA(cudaStreamSynchronize(ctx->stream));
  if (!h->node_index) MVS_TRY(mvs_nn_index_create(ctx, &h->node_index));
  MVS_TRY(nn_build(h->node_index, h->g.p, m, 2.0f));
  h->graph_built = false;
  return MVS_OK;
}

}  // namespace mvs

using namespace mvs;

extern "C" {

int mvs_deform_set_nodes(mvs_deform* h, const int32_t* vertex_ids, int64_t n_nodes) {
  MVS_REQUIRE(h && vertex_ids && n_nodes > 0, "bad argument");
  MVS_TRY(ctx_ensure_device(h->ctx));
  std::vector<int32_t> ids(vertex_ids, vertex_ids + n_nodes);
  return set_nodes_from_host(h, ids);
}

int mvs_deform_get_nodes(mvs_deform* h, int32_t* vertex_ids, int64_t capacity, int64_t* n_nodes) {
  MVS_REQUIRE(h, "null handle");
  if (n_nodes) *n_nodes = h->m;
  if (vertex_ids) {
    if (capacity < h->m) {
      set_error("get_nodes: capacity %lld < %lld", (long long)capacity, (long long)h->m);
      return MVS_ERR_INVALID_ARG;
    }
    memcpy(vertex_ids, h->node_ids.data(), sizeof(int32_t) * h->m);
  }
  return MVS_OK;
}

// Deformation/Deformation.cpp:63-106.  The k-NN lists of all vertices are computed on the device;
// the greedy keep/suppress scan is inherently sequential in vertex order and runs on the host
// (one pass over n*k ints), like the reference.
int mvs_deform_uniform_sampling(mvs_deform* h, int suppress_k) {
  MVS_REQUIRE(h, "null handle");
  if (suppress_k < 1 || suppress_k > 32) {
    set_error("uniform_sampling: suppress_k must be in 1..32");
    return MVS_ERR_INVALID_ARG;
  }
  mvs_ctx* ctx = h->ctx;
  MVS_TRY(ctx_ensure_device(ctx));
  int64_t n = h->n;
  mvs_nn_index* vi = nullptr;
  MVS_TRY(mvs_nn_index_create(ctx, &vi));
  DevBuf<int> nn;
  int st = nn.reserve((size_t)n * suppress_k);
  if (st == MVS_OK) st = nn_build(vi, h->v0.p, n);
  if (st == MVS_OK) st = nn_knn(vi, h->v0.p, n, suppress_k, nn.p, nullptr);
  std::vector<int32_t> hn((size_t)n * suppress_k);
  if (st == MVS_OK) st = d2h_sync(ctx, hn.data(), nn.p, sizeof(int) * hn.size());
  mvs_nn_index_destroy(vi);
  MVS_TRY(st);
  std::vector<char> removed((size_t)n, 0);
  std::vector<int32_t> ids;
  for (int64_t i = 0; i < n; ++i) {
    if (removed[i]) continue;
    ids.push_back((int32_t)i);
    for (int j = 0; j < suppress_k; ++j) {
      int32_t id = hn[i * suppress_k + j];
      if (id >= 0 && id != i) removed[id] = 1;
    }
  }
  return set_nodes_from_host(h, ids);
}

int mvs_deform_sample_fps(mvs_deform* h, int64_t n_nodes, int64_t first_vertex) {
  MVS_REQUIRE(h && n_nodes > 0 && first_vertex >= 0 && first_vertex < h->n, "bad argument");
  mvs_ctx* ctx = h->ctx;
  MVS_TRY(ctx_ensure_device(ctx));
  int64_t n = h->n;
  if (n_nodes > n) n_nodes = n;
  if (n <= (int64_t)kFpsCluster * kFpsCap) {
    // small set: the whole sampling in one cluster launch
    DevBuf<int> out;
    MVS_TRY(out.reserve((size_t)n_nodes));
    // cluster size: the fewest CTAs that hold the set at <= 12 points per thread (fewer CTAs leave
    // room for other streams' samplings), points per thread: the smallest unrolled variant
    // the per-node time is a chain of synchronisations, not arithmetic: 4 CTAs x 17 points per thread
    // take 0.92 ms for 512 nodes of 34.7K points, 8 CTAs x 9 points 1.10 ms
    int nblk = 4;
    if (const char* env = getenv("MVS_FPS_CLUSTER")) {
      int e = atoi(env);
      if (e == 1 || e == 2 || e == 4 || e == 8) nblk = e;
    }
    while (nblk < kFpsCluster && (int64_t)nblk * kFpsCap < n) nblk *= 2;
    const int per = (int)((n + nblk - 1) / nblk);
    const int ppt = (per + kFpsThreads - 1) / kFpsThreads;
    const float4* P = h->v0.p;
    const int nn = (int)n, nd = (int)n_nodes, f0 = (int)first_vertex;
    if (ppt <= 3)
      MVS_TRY(launch_fps_cluster<3>(ctx, P, nn, nd, f0, out.p, nblk));
    else if (ppt <= 6)
      MVS_TRY(launch_fps_cluster<6>(ctx, P, nn, nd, f0, out.p, nblk));
    else if (ppt <= 9)
      MVS_TRY(launch_fps_cluster<9>(ctx, P, nn, nd, f0, out.p, nblk));
    else if (ppt <= 12)
      MVS_TRY(launch_fps_cluster<12>(ctx, P, nn, nd, f0, out.p, nblk));
    else if (ppt <= 16)
      MVS_TRY(launch_fps_cluster<16>(ctx, P, nn, nd, f0, out.p, nblk));
    else
      MVS_TRY(launch_fps_cluster<kFpsMaxPpt>(ctx, P, nn, nd, f0, out.p, nblk));
    MVS_CUDA(cudaGetLastError());
    std::vector<int32_t> ids((size_t)n_nodes);
    MVS_TRY(d2h_sync(ctx, ids.data(), out.p, sizeof(int) * n_nodes));
    return set_nodes_from_host(h, ids);
  }
  int grid = grid_for(ctx, n, 256, 4);
  DevBuf<float> mind, pval;
  DevBuf<int> pidx, cur, out;
  DevBuf<unsigned> ticket;
  MVS_TRY(mind.reserve((size_t)n));
  MVS_TRY(pval.reserve((size_t)grid));
  MVS_TRY(pidx.reserve((size_t)grid));
  MVS_TRY(cur.reserve(1));
  MVS_TRY(out.reserve((size_t)n_nodes));
  MVS_TRY(ticket.reserve(1));
  std::vector<float> inf((size_t)n, INFINITY);
  MVS_CUDA(cudaMemcpyAsync(mind.p, inf.data(), sizeof(float) * n, cudaMemcpyHostToDevice, ctx->stream));
  int c0 = (int)first_vertex;
  MVS_CUDA(cudaMemcpyAsync(cur.p, &c0, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  MVS_CUDA(cudaMemsetAsync(ticket.p, 0, sizeof(unsigned), ctx->stream));
  MVS_CUDA(cudaStreamSynchronize(ctx->stream));
  {
    ProfScope ps(ctx, "fps", n_nodes);
    for (int64_t s = 0; s < n_nodes; ++s)
      k_fps_step<<<grid, 256, 0, ctx->stream>>>(h->v0.p, n, mind.p, cur.p, out.p, (int)s, pval.p,
                                                pidx.p, ticket.p);
  }
  MVS_CUDA(cudaGetLastError());
  std::vector<int32_t> ids((size_t)n_nodes);
  MVS_TRY(d2h_sync(ctx, ids.data(), out.p, sizeof(int) * n_nodes));
  return set_nodes_from_host(h, ids);
}

int mvs_deform_build_graph(mvs_deform* h, int node_k, int graph_k) {
  MVS_REQUIRE(h, "null handle");
  return deform_build_graph(h, node_k, graph_k);
}

}  // extern "C"
