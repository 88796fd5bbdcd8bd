// gn_assemble.cu — per-vertex passes of one Gauss-Newton iteration of the embedded-deformation
// solve (formulation: SURVEY.md §8(a) D7-D8):
//
//   k_deform_corr   v~_i = sum_j w_ij [A_j (v_i - g_j) + g_j + t_j]; exact 1-NN of v~_i in the
//                   target grid; resid_i = (v~_i - c_i, omega_i)            48 B/vertex in, 16 out
//   k_assemble_fit  fused residual + Jacobian + J^T J / J^T r assembly.  The fit Jacobian of a
//                   vertex w.r.t. node j is w_ij [d^T (x) I3 | I3], so every 12x12 block pair is
//                   kron(b_j b_l^T, I3) with b_j = w_ij (d, 1) in R^4: only the 4x4 factor is
//                   accumulated (16 MACs instead of 432).  Vertices are pre-sorted so that equal
//                   node 4-tuples are adjacent; a warp finds the groups of equal tuples among its
//                   32 vertices with __match_any_sync, stages the per-vertex records in shared
//                   memory, reduces each group in registers (14 4x4 tiles x 2 half-warps) and
//                   flushes one vector atomic per tile row.                  64 B/vertex in
//   k_node_terms    rotation (dense 9x9 per node) and smoothness (graph edges) terms
//   k_precond       12x12 block-Jacobi preconditioner: Cholesky inverse of each diagonal block
#include <stdlib.h>

#include <cmath>

#include "comm.cuh"
#include "deform.cuh"
#include "nn_pyramid.cuh"
#include "reduce.cuh"

namespace mvs {

// ------------------------------------------------------------------ helpers
struct NodeX {
  float a[12];  // a1 (0..2), a2 (3..5), a3 (6..8), t (9..11)
};

__device__ __forceinline__ NodeX load_node(const float* __restrict__ x, int j) {
  NodeX n;
  const float4* p = reinterpret_cast<const float4*>(x + (size_t)j * 12);
  float4 u0 = __ldg(p), u1 = __ldg(p + 1), u2 = __ldg(p + 2);
  n.a[0] = u0.x; n.a[1] = u0.y; n.a[2] = u0.z; n.a[3] = u0.w;
  n.a[4] = u1.x; n.a[5] = u1.y; n.a[6] = u1.z; n.a[7] = u1.w;
  n.a[8] = u2.x; n.a[9] = u2.y; n.a[10] = u2.z; n.a[11] = u2.w;
  return n;
}

// v~ = sum_a w_a [A_a (v - g_a) + g_a + t_a] over the vertex's nodes in ASCENDING node id order, in a
// fixed float32 sequence of explicit roundings (one product, two fused multiply-adds, two additions,
// one fused multiply-add per component and node) so that the oracle (oracle/deform.cpp
// deform_vertex_f32, std::fmaf in the same order) produces the SAME float32 query point and with it
// bit-identical correspondence indices (north_star: correspondence indices bit-exact).
__device__ __forceinline__ float3 deform_point(float4 v, int4 id, float4 w,
                                               const float4* __restrict__ g,
                                               const float* __restrict__ x) {
  int ids[4] = {id.x, id.y, id.z, id.w};
  float ws[4] = {w.x, w.y, w.z, w.w};
  float3 out = make_float3(0.f, 0.f, 0.f);
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    if (ids[a] < 0) continue;
    float4 gj = __ldg(g + ids[a]);
    NodeX n = load_node(x, ids[a]);
    float d0 = __fsub_rn(v.x, gj.x), d1 = __fsub_rn(v.y, gj.y), d2 = __fsub_rn(v.z, gj.z);
    float px = __fmaf_rn(n.a[6], d2, __fmaf_rn(n.a[3], d1, __fmul_rn(n.a[0], d0)));
    float py = __fmaf_rn(n.a[7], d2, __fmaf_rn(n.a[4], d1, __fmul_rn(n.a[1], d0)));
    float pz = __fmaf_rn(n.a[8], d2, __fmaf_rn(n.a[5], d1, __fmul_rn(n.a[2], d0)));
    px = __fadd_rn(__fadd_rn(px, gj.x), n.a[9]);
    py = __fadd_rn(__fadd_rn(py, gj.y), n.a[10]);
    pz = __fadd_rn(__fadd_rn(pz, gj.z), n.a[11]);
    out.x = __fmaf_rn(ws[a], px, out.x);
    out.y = __fmaf_rn(ws[a], py, out.y);
    out.z = __fmaf_rn(ws[a], pz, out.z);
  }
  return out;
}

__global__ void k_reset_state(float* __restrict__ x, int64_t m) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= m * 12) return;
  int k = (int)(t % 12);
  x[t] = (k == 0 || k == 4 || k == 8) ? 1.0f : 0.0f;
}

int gn_reset_state(mvs_deform* h) {
  ProfScope ps(h->ctx, "gn_reset");
  k_reset_state<<<div_up(h->m * 12, 256), 256, 0, h->ctx->stream>>>(h->x.p, h->m);
  MVS_CUDA(cudaGetLastError());
  return MVS_OK;
}

// ------------------------------------------------------------------ deform + correspondences
struct Best1 {
  float qx, qy, qz;
  float d2;
  int idx;
  float cx, cy, cz;
  __device__ __forceinline__ float worst2() const { return d2; }
  __device__ __forceinline__ void point(float4 p) {
    float d = dist2_rn(p.x, p.y, p.z, qx, qy, qz);
    int id = __float_as_int(p.w);
    if (d < d2 || (d == d2 && id < idx)) {
      d2 = d;
      idx = id;
      cx = p.x; cy = p.y; cz = p.z;
    }
  }
};

// flat index of a launch -> sorted vertex inside the launch's sub-ranges (-1 past the end)
__device__ __forceinline__ int64_t range_index(const ShardRanges& R, int64_t t) {
  if (t >= R.voff[R.n]) return -1;
  int k = 0;
  while (t >= R.voff[k + 1]) ++k;
  return R.b[k] + (t - R.voff[k]);
}

// MULTI = false: one contiguous range (the single-GPU case) passed as two scalars
template <bool MULTI>
__global__ void __launch_bounds__(128, 8)
    k_deform_corr(PyrParams tg, const float4* __restrict__ tpts, int* __restrict__ corr_prev,
                  const float4* __restrict__ tnrm,
                  const float4* __restrict__ v, const float4* __restrict__ nrm,
                  const int4* __restrict__ vidx, const float4* __restrict__ vw,
                  const float4* __restrict__ g, const float* __restrict__ x, ShardRanges R,
                  int64_t begin, int64_t end, float max_d2, float min_cos,
                  float4* __restrict__ resid) {
  int64_t i;
  if (MULTI) {
    i = range_index(R, blockIdx.x * (int64_t)blockDim.x + threadIdx.x);
    if (i < 0) return;
  } else {
    i = begin + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= end) return;
  }
  float4 vi = __ldg(v + i);
  float3 q = deform_point(vi, __ldg(vidx + i), __ldg(vw + i), g, x);
  Best1 b;
  b.qx = q.x; b.qy = q.y; b.qz = q.z;
  b.d2 = INFINITY;
  b.idx = 0x7fffffff;
  b.cx = b.cy = b.cz = 0.f;
  // seed with the previous iteration's correspondence: the bound starts tight, so the walk only
  // touches the few cells around the query
  int prev = corr_prev[i];
  if (prev >= 0) {
    float4 c = __ldg(tpts + prev);
    c.w = __int_as_float(prev);
    b.point(c);
  }
  pyr_search(tg, q.x, q.y, q.z, b);
  corr_prev[i] = b.idx;
  float om = 1.0f;
  if (max_d2 > 0.f && b.d2 > max_d2) om = 0.0f;
  if (min_cos > -1.0f && tnrm) {
    float4 ni = __ldg(nrm + i);
    float4 nt = __ldg(tnrm + b.idx);
    // double, FMA-free: the accept/reject decision is reproducible on the host
    double dn = __dadd_rn(__dadd_rn(__dmul_rn((double)ni.x, (double)nt.x),
                                    __dmul_rn((double)ni.y, (double)nt.y)),
                          __dmul_rn((double)ni.z, (double)nt.z));
    if (dn < (double)min_cos) om = 0.0f;
  }
  resid[i] = make_float4(q.x - b.cx, q.y - b.cy, q.z - b.cz, om);
}

__global__ void k_deform_resid(const float4* __restrict__ v, const int4* __restrict__ vidx,
                               const float4* __restrict__ vw, const float4* __restrict__ g,
                               const float* __restrict__ x, const float4* __restrict__ targets,
                               ShardRanges R, float4* __restrict__ resid) {
  int64_t i = range_index(R, blockIdx.x * (int64_t)blockDim.x + threadIdx.x);
  if (i < 0) return;
  float4 t = __ldg(targets + i);
  if (!(t.w > 0.f)) {
    resid[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  float3 q = deform_point(__ldg(v + i), __ldg(vidx + i), __ldg(vw + i), g, x);
  resid[i] = make_float4(q.x - t.x, q.y - t.y, q.z - t.z, t.w);
}

// the rank's sub-ranges in launch batches of at most kMaxShardRanges
static std::vector<ShardRanges> range_batches(const mvs_deform* h) {
  std::vector<ShardRanges> out;
  const auto ranges = h->shard_ranges();
  for (size_t k0 = 0; k0 < ranges.size(); k0 += kMaxShardRanges) {
    ShardRanges R;
    memset(&R, 0, sizeof(R));
    R.n = (int)std::min<size_t>(kMaxShardRanges, ranges.size() - k0);
    for (int k = 0; k < R.n; ++k) {
      R.b[k] = ranges[k0 + k].first;
      R.e[k] = ranges[k0 + k].second;
      R.voff[k + 1] = R.voff[k] + (R.e[k] - R.b[k]);
      R.g_lo[k] = h->group_at(R.b[k]);
      int g_hi = R.e[k] >= h->n ? (int)h->ngroups : h->group_at(R.e[k]);
      R.goff[k + 1] = R.goff[k] + std::max(0, g_hi - R.g_lo[k]);
    }
    out.push_back(R);
  }
  return out;
}

int gn_deform_and_correspond(mvs_deform* h) {
  mvs_ctx* ctx = h->ctx;
  const auto batches = range_batches(h);
  if (batches.empty()) return MVS_OK;
  if (h->prm.corr_mode == 1) {
    if (!h->tgt_index || !h->tgt_index->pyr_built) {
      set_error("corr_mode 1 needs a target (mvs_deform_set_target*)");
      return MVS_ERR_STATE;
    }
  } else if (!h->have_targets) {
    set_error("fixed-target mode needs per-vertex targets");
    return MVS_ERR_STATE;
  }
  for (const ShardRanges& R : batches) {
    const int64_t nv = R.voff[R.n];
    if (nv <= 0) continue;
    if (h->prm.corr_mode == 1) {
      ProfScope ps(ctx, "gn_deform_corr");
      if (R.n == 1) {
        ShardRanges none;
        none.n = 0;
        k_deform_corr<false><<<div_up(nv, 128), 128, 0, ctx->stream>>>(
            h->tgt_index->pyr, h->tp, h->corr_prev.p, h->tn, h->v.p, h->nrm.p, h->vidx.p, h->vw.p,
            h->g.p, h->x.p, none, R.b[0], R.e[0], (float)h->prm.max_corr_dist2,
            (float)h->prm.min_corr_cos, h->resid.p);
      } else {
        k_deform_corr<true><<<div_up(nv, 128), 128, 0, ctx->stream>>>(
            h->tgt_index->pyr, h->tp, h->corr_prev.p, h->tn, h->v.p, h->nrm.p, h->vidx.p, h->vw.p,
            h->g.p, h->x.p, R, 0, 0, (float)h->prm.max_corr_dist2, (float)h->prm.min_corr_cos,
            h->resid.p);
      }
    } else {
      ProfScope ps(ctx, "gn_deform_resid");
      k_deform_resid<<<div_up(nv, 256), 256, 0, ctx->stream>>>(
          h->v.p, h->vidx.p, h->vw.p, h->g.p, h->x.p, h->targets.p, R, h->resid.p);
    }
  }
  MVS_CUDA(cudaGetLastError());
  return MVS_OK;
}

// ------------------------------------------------------------------ fit assembly
constexpr int kAsmWarps = 4;
constexpr int kRecStride = 20;

__device__ __forceinline__ int find_slot(const int* __restrict__ row_ptr,
                                         const int* __restrict__ col, int j, int l) {
  int lo = __ldg(row_ptr + j), hi = __ldg(row_ptr + j + 1) - 1;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (__ldg(col + mid) < l) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// order-independent accumulation: value * 2^k rounded to int64, summed with integer atomics
// `flag` (the context's pinned solver flag, may be null): raised when a value is not finite or does not
// fit the fixed-point range — the bound behind gn_fx_scale assumes targets within reach of the mesh
__device__ __forceinline__ void fx_add(long long* addr, float v, float scale, int* flag = nullptr) {
  const float sv = v * scale;
  if (flag && !(fabsf(sv) < 4.0e18f)) {
    volatile int* f = reinterpret_cast<volatile int*>(flag);
    f[0] = 3;
    f[1] = __float_as_int(v);      // diagnostics for the error text
    f[2] = __float_as_int(scale);
  }
  atomicAdd(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__float2ll_rn(sv));
}

// The kernel is bound by the issue rate of the global reductions (RED: ~1.3 cycles per lane per SM),
// not by the 48 B/vertex stream, so the design minimises reductions per vertex:
//   * only the upper triangle of J^T J is accumulated (10 of the 16 block pairs of a vertex; the
//     conversion pass mirrors it);
//   * a warp owns a GROUP of kGrpRuns consecutive tuple runs (a run = consecutive sorted vertices
//     that share their four nodes, ~22 vertices at C3).  Runs of a group mostly share their primary
//     node, so their tiles hit the same blocks: the group keeps one shared-memory accumulator per
//     distinct destination (table built with the graph) and flushes each ONCE, as fixed-point
//     integer atomics (exactly commutative: bitwise reproducible);
//   * per run: the per-vertex records b_a = sqrt(w_fit omega) w_a (v - g_a, 1), a = 0..3, and the
//     scaled residual are staged in shared memory, 14 tile lanes x 2 half-warps accumulate the 10
//     products b_a b_b^T (a <= b) and the 4 gradient tiles b_a r^T in registers;
//   * the next 32 vertices (resid, v, vw: 3 x 16 B each; node ids are not read — the run table
//     carries the node positions) are in flight while the current ones are processed.
struct AsmRaw {
  float4 rs, vi, w;
};

// 4 CTAs per SM: 121 registers, no spills — 57.5 us per launch at C3 against 62.5 at 5 CTAs
// (96 registers, 24 B of spill), 65 at 3 and 6, 84 at 8 (64 registers, 144 B of spill)
constexpr int kAsmCtasPerSm = 4;

__global__ void __launch_bounds__(kAsmWarps * 32, kAsmCtasPerSm)
    k_assemble_fit(const float4* __restrict__ v, const float4* __restrict__ vw,
                   const float4* __restrict__ resid, const float4* __restrict__ run_g,
                   const unsigned char* __restrict__ run_local, const int* __restrict__ run_start,
                   const int* __restrict__ grp_n, const int* __restrict__ grp_dst,
                   long long* __restrict__ sys64, int nup16, int nruns, ShardRanges R, float w_fit,
                   float fx_scale, unsigned* __restrict__ next_group, int* flag) {
  __shared__ __align__(16) float rec[kAsmWarps][32][kRecStride];
  __shared__ __align__(16) float accs[kAsmWarps][kGrpMax][16];
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int half = lane >> 4, t = lane & 15;
  int ta, tb;
  switch (t) {
    case 0: ta = 0; tb = 0; break;
    case 1: ta = 1; tb = 1; break;
    case 2: ta = 2; tb = 2; break;
    case 3: ta = 3; tb = 3; break;
    case 4: ta = 0; tb = 1; break;
    case 5: ta = 0; tb = 2; break;
    case 6: ta = 0; tb = 3; break;
    case 7: ta = 1; tb = 2; break;
    case 8: ta = 1; tb = 3; break;
    case 9: ta = 2; tb = 3; break;
    default: ta = (t - 10) & 3; tb = 4; break;  // gradient tiles read the residual (cols 16..19)
  }
  int64_t begin = 0, end = 1;  // the sub-range of the current group (set per ticket)
  auto fetch = [&](int64_t i) {
    AsmRaw r;
    i = min(i, end - 1);  // lanes past the run are masked when the pass is processed
    r.rs = __ldg(resid + i);
    r.vi = __ldg(v + i);
    r.w = __ldg(vw + i);
    return r;
  };
  // groups are handed out dynamically (run lengths vary); the sums are integer, so the order in
  // which groups finish cannot change the result
  for (;;) {
    int q = 0;
    if (lane == 0) q = (int)atomicAdd(next_group, 1u);
    q = __shfl_sync(full, q, 0);
    if (q >= R.goff[R.n]) break;
    // ticket -> (sub-range, group): sub-ranges start and end on group boundaries
    int kr = 0;
    while (q >= R.goff[kr + 1]) ++kr;
    q = R.g_lo[kr] + (q - R.goff[kr]);
    begin = R.b[kr];
    end = R.e[kr];
    const int run_lo = R.g_lo[kr] * kGrpRuns;
    const int run_hi = min((R.g_lo[kr] + (R.goff[kr + 1] - R.goff[kr])) * kGrpRuns, nruns) - 1;
    const int r_base = q * kGrpRuns;
    const int r_first = max(r_base, run_lo), r_last = min(r_base + kGrpRuns - 1, run_hi);
    // lane l <= kGrpRuns: first vertex of run r_base + l
    const int my_rs = __ldg(run_start + min(r_base + min(lane, kGrpRuns), nruns));
    const int L = __ldg(grp_n + q);
    for (int k = lane; k < L * 4; k += 32)
      reinterpret_cast<float4*>(&accs[warp][0][0])[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    int run = r_first;
    int64_t rs1 = min((int64_t)__shfl_sync(full, my_rs, run - r_base + 1), end);
    int64_t base = max((int64_t)__shfl_sync(full, my_rs, run - r_base), begin);
    AsmRaw raw = fetch(base + lane);
    float4 mg = make_float4(0.f, 0.f, 0.f, 0.f);
    int mloc = 255;
    if (lane < 4) mg = __ldg(run_g + (size_t)run * 4 + lane);
    if (lane < 14) mloc = __ldg(run_local + (size_t)run * 16 + lane);
    bool new_run = true;
    float gx[4], gy[4], gz[4], gv[4];
    float acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = 0.f;
    int loc = 255;
    __syncwarp();
    for (;;) {
      if (new_run) {
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          gx[a] = __shfl_sync(full, mg.x, a);
          gy[a] = __shfl_sync(full, mg.y, a);
          gz[a] = __shfl_sync(full, mg.z, a);
          gv[a] = __shfl_sync(full, mg.w, a);
        }
        loc = mloc;
      }
      const int cnt = (int)min((int64_t)32, rs1 - base);
      // ---- next pass: same run, or the first pass of the next run of the group
      int nrun = run;
      int64_t nbase = base + 32, nrs1 = rs1;
      if (nbase >= rs1) {
        nrun = run + 1;
        if (nrun <= r_last) {
          nbase = max((int64_t)__shfl_sync(full, my_rs, nrun - r_base), begin);
          nrs1 = min((int64_t)__shfl_sync(full, my_rs, nrun - r_base + 1), end);
        }
      }
      const bool has_next = nrun <= r_last;
      AsmRaw raw_n = raw;
      if (has_next) {
        raw_n = fetch(nbase + lane);
        if (nrun != run) {
          if (lane < 4) mg = __ldg(run_g + (size_t)nrun * 4 + lane);
          if (lane < 14) mloc = __ldg(run_local + (size_t)nrun * 16 + lane);
        }
      }
      // ---- records of this pass
      float r0[kRecStride];
#pragma unroll
      for (int k = 0; k < kRecStride; ++k) r0[k] = 0.f;
      if (lane < cnt && raw.rs.w > 0.f) {
        float s = sqrtf(w_fit * raw.rs.w);
        float ws[4] = {raw.w.x, raw.w.y, raw.w.z, raw.w.w};
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          if (gv[a] > 0.f) {
            float sw = s * ws[a];
            r0[4 * a + 0] = sw * (raw.vi.x - gx[a]);
            r0[4 * a + 1] = sw * (raw.vi.y - gy[a]);
            r0[4 * a + 2] = sw * (raw.vi.z - gz[a]);
            r0[4 * a + 3] = sw;
          }
        }
        r0[16] = s * raw.rs.x;
        r0[17] = s * raw.rs.y;
        r0[18] = s * raw.rs.z;
      }
      {
        float4* dst = reinterpret_cast<float4*>(&rec[warp][lane][0]);
#pragma unroll
        for (int k = 0; k < 5; ++k)
          dst[k] = make_float4(r0[4 * k], r0[4 * k + 1], r0[4 * k + 2], r0[4 * k + 3]);
      }
      __syncwarp();
      if (t < 14) {
        // members mi = half, half + 2, ...; records past cnt are zero (the whole pass is rewritten
        // every time), so the loop runs two members per step without a tail
        const int cnt4 = (cnt + 3) & ~3;
        for (int mi = half; mi < cnt4; mi += 4) {
          const float4 ba0 = *reinterpret_cast<const float4*>(&rec[warp][mi][4 * ta]);
          const float4 bb0 = *reinterpret_cast<const float4*>(&rec[warp][mi][4 * tb]);
          const float4 ba1 = *reinterpret_cast<const float4*>(&rec[warp][mi + 2][4 * ta]);
          const float4 bb1 = *reinterpret_cast<const float4*>(&rec[warp][mi + 2][4 * tb]);
          acc[0] += ba0.x * bb0.x; acc[1] += ba0.x * bb0.y; acc[2] += ba0.x * bb0.z; acc[3] += ba0.x * bb0.w;
          acc[4] += ba0.y * bb0.x; acc[5] += ba0.y * bb0.y; acc[6] += ba0.y * bb0.z; acc[7] += ba0.y * bb0.w;
          acc[8] += ba0.z * bb0.x; acc[9] += ba0.z * bb0.y; acc[10] += ba0.z * bb0.z; acc[11] += ba0.z * bb0.w;
          acc[12] += ba0.w * bb0.x; acc[13] += ba0.w * bb0.y; acc[14] += ba0.w * bb0.z; acc[15] += ba0.w * bb0.w;
          acc[0] += ba1.x * bb1.x; acc[1] += ba1.x * bb1.y; acc[2] += ba1.x * bb1.z; acc[3] += ba1.x * bb1.w;
          acc[4] += ba1.y * bb1.x; acc[5] += ba1.y * bb1.y; acc[6] += ba1.y * bb1.z; acc[7] += ba1.y * bb1.w;
          acc[8] += ba1.z * bb1.x; acc[9] += ba1.z * bb1.y; acc[10] += ba1.z * bb1.z; acc[11] += ba1.z * bb1.w;
          acc[12] += ba1.w * bb1.x; acc[13] += ba1.w * bb1.y; acc[14] += ba1.w * bb1.z; acc[15] += ba1.w * bb1.w;
        }
      }
      new_run = nrun != run;
      if (new_run) {
        // ---- the run is complete: fold the two half-warps, add the tiles to the group accumulators
#pragma unroll
        for (int k = 0; k < 16; ++k) acc[k] += __shfl_xor_sync(full, acc[k], 16);
        if (half == 0 && t < 14 && loc != 255) {
          float4* o = reinterpret_cast<float4*>(&accs[warp][loc][0]);
          if (t < 10) {
            float4 o0 = o[0], o1 = o[1], o2 = o[2], o3 = o[3];
            o0.x += acc[0]; o0.y += acc[1]; o0.z += acc[2]; o0.w += acc[3];
            o1.x += acc[4]; o1.y += acc[5]; o1.z += acc[6]; o1.w += acc[7];
            o2.x += acc[8]; o2.y += acc[9]; o2.z += acc[10]; o2.w += acc[11];
            o3.x += acc[12]; o3.y += acc[13]; o3.z += acc[14]; o3.w += acc[15];
            o[0] = o0; o[1] = o1; o[2] = o2; o[3] = o3;
          } else {
            // rhs_j[3c + r] = sum b_a[c] * r[r]   (acc[4c + r], r < 3)
            float4 o0 = o[0], o1 = o[1], o2 = o[2];
            o0.x += acc[0]; o0.y += acc[1]; o0.z += acc[2]; o0.w += acc[4];
            o1.x += acc[5]; o1.y += acc[6]; o1.z += acc[8]; o1.w += acc[9];
            o2.x += acc[10]; o2.y += acc[12]; o2.z += acc[13]; o2.w += acc[14];
            o[0] = o0; o[1] = o1; o[2] = o2;
          }
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) acc[k] = 0.f;
      }
      __syncwarp();  // this pass's records are consumed; accumulator updates are visible
      if (!has_next) break;
      run = nrun;
      base = nbase;
      rs1 = nrs1;
      raw = raw_n;
    }
    // ---- flush the group: 16 consecutive lanes hit 16 consecutive int64
    const int* gd = grp_dst + (size_t)q * kGrpMax;
    for (int k = lane; k < L * 16; k += 32) {
      const int tile = k >> 4, e = k & 15;
      const int d = __ldg(gd + tile);
      if (e < (d >= nup16 ? 12 : 16)) fx_add(sys64 + d + e, accs[warp][tile][e], fx_scale, flag);
    }
    __syncwarp();
  }
}

// fixed point [U: nup x 16 | rhs] -> float32 system [B: nnz x 16 | rhs]; a lower-triangle block is the
// transpose of its accumulated mirror (tslot[e] = slot of the transposed block; rows ascend, so e is
// a lower-triangle entry iff tslot[e] < e; uidx[e] = where the block or its mirror sits in U)
__global__ void k_sys_convert(const long long* __restrict__ in, float* __restrict__ out, int64_t n,
                              int64_t nnz16, int64_t nup16, const int* __restrict__ tslot,
                              const int* __restrict__ uidx, float inv_scale) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    int64_t src;
    if (i < nnz16) {
      int e = (int)(i >> 4), el = (int)(i & 15);
      bool lower = __ldg(tslot + e) < e;
      src = (int64_t)__ldg(uidx + e) * 16 + (lower ? (((el & 3) << 2) | (el >> 2)) : el);
    } else {
      src = nup16 + (i - nnz16);
    }
    out[i] = (float)((double)in[src] * (double)inv_scale);
  }
}

// ------------------------------------------------------------------ node terms
// 16 lanes per node: lanes 0..8 each build one row of the dense 9x9 rotation block (and its
// gradient entry); every lane takes the graph edges e = lane, lane+16, ... of the smoothness term.
// The node's own 4x4 diagonal factor and gradient are reduced over the 16 lanes with shuffles and
// flushed once; the (j,l), (l,j), (l,l) contributions of an edge go out as atomics.
constexpr int kNodeLanes = 16;

__global__ void __launch_bounds__(128)
    k_node_terms(const float4* __restrict__ g, const float* __restrict__ x,
                 const int* __restrict__ graph, int gk1, int64_t m,
                 const int* __restrict__ row_ptr, const int* __restrict__ col,
                 const int* __restrict__ uidx, long long* __restrict__ B, float* __restrict__ Rot,
                 long long* __restrict__ rhs, float w_rot, float w_reg, float fxs, int add_fixed) {
  const unsigned full = 0xffffffffu;
  int64_t j = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / kNodeLanes;
  int k = threadIdx.x & (kNodeLanes - 1);
  bool on = j < m;
  int jj = on ? (int)j : 0;
  NodeX n = load_node(x, jj);
  const float* a1 = n.a;
  const float* a2 = n.a + 3;
  const float* a3 = n.a + 6;
  float gl[12];
#pragma unroll
  for (int p = 0; p < 12; ++p) gl[p] = 0.f;
  if (on && k < 9) {
    float res[6];
    res[0] = a1[0] * a2[0] + a1[1] * a2[1] + a1[2] * a2[2];
    res[1] = a1[0] * a3[0] + a1[1] * a3[1] + a1[2] * a3[2];
    res[2] = a2[0] * a3[0] + a2[1] * a3[1] + a2[2] * a3[2];
    res[3] = a1[0] * a1[0] + a1[1] * a1[1] + a1[2] * a1[2] - 1.0f;
    res[4] = a2[0] * a2[0] + a2[1] * a2[1] + a2[2] * a2[2] - 1.0f;
    res[5] = a3[0] * a3[0] + a3[1] * a3[1] + a3[2] * a3[2] - 1.0f;
    // column p of the 6x9 Jacobian, for p = k and for every q
    auto Jcol = [&](int p, float out[6]) {
      int c = p / 3, r = p - 3 * c;  // unknown a_{c+1}[r]
      out[0] = c == 0 ? a2[r] : (c == 1 ? a1[r] : 0.f);
      out[1] = c == 0 ? a3[r] : (c == 2 ? a1[r] : 0.f);
      out[2] = c == 1 ? a3[r] : (c == 2 ? a2[r] : 0.f);
      out[3] = c == 0 ? 2.f * a1[r] : 0.f;
      out[4] = c == 1 ? 2.f * a2[r] : 0.f;
      out[5] = c == 2 ? 2.f * a3[r] : 0.f;
    };
    float Jp[6];
    Jcol(k, Jp);
    float gp = 0.f;
#pragma unroll
    for (int e = 0; e < 6; ++e) gp += Jp[e] * res[e];
    float* R = Rot + (size_t)j * 81 + 9 * k;
#pragma unroll
    for (int q = 0; q < 9; ++q) {
      float Jq[6];
      Jcol(q, Jq);
      float sacc = 0.f;
#pragma unroll
      for (int e = 0; e < 6; ++e) sacc += Jp[e] * Jq[e];
      R[q] = w_rot * sacc;
    }
    // gradient entry k goes to lane-local gl[k] (static index via unrolled select)
#pragma unroll
    for (int p = 0; p < 9; ++p)
      if (p == k) gl[p] = w_rot * gp;
  }
  float Bjj[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) Bjj[q] = 0.f;
  if (on) {
    float4 gj = __ldg(g + j);
    for (int e = k; e < gk1; e += kNodeLanes) {
      int l = __ldg(graph + (size_t)j * gk1 + e);
      if (l < 0 || l == j) continue;
      float4 gv = __ldg(g + l);
      NodeX nl = load_node(x, l);
      float ev[3] = {gv.x - gj.x, gv.y - gj.y, gv.z - gj.z};
      float r[3];
#pragma unroll
      for (int rr = 0; rr < 3; ++rr)
        r[rr] = (n.a[rr] * ev[0] + n.a[3 + rr] * ev[1] + n.a[6 + rr] * ev[2] - ev[rr]) +
                (n.a[9 + rr] - nl.a[9 + rr]);
      float a[4] = {ev[0], ev[1], ev[2], 1.0f};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int d = 0; d < 4; ++d) Bjj[4 * c + d] += w_reg * a[c] * a[d];
#pragma unroll
        for (int rr = 0; rr < 3; ++rr) gl[3 * c + rr] += w_reg * a[c] * r[rr];
      }
      // coupling block: B_jl[c][3] = B_lj[3][c] = -w a_c; only the upper-triangle copy is
      // accumulated (k_sys_convert mirrors it)
      long long* Bup = B + (size_t)__ldg(uidx + find_slot(row_ptr, col, (int)j, l)) * 16;
      long long* Bll = B + (size_t)__ldg(uidx + find_slot(row_ptr, col, l, l)) * 16;
      if (add_fixed) {
#pragma unroll
        for (int c = 0; c < 4; ++c) fx_add(Bup + (j < l ? 4 * c + 3 : 12 + c), -w_reg * a[c], fxs);
        fx_add(Bll + 15, w_reg, fxs);
#pragma unroll
        for (int rr = 0; rr < 3; ++rr) fx_add(rhs + (size_t)l * 12 + 9 + rr, -w_reg * r[rr], fxs);
      }
    }
  }
  // reduce the node's own contributions over its 16 lanes
#pragma unroll
  for (int o = kNodeLanes / 2; o > 0; o >>= 1) {
#pragma unroll
    for (int q = 0; q < 16; ++q) Bjj[q] += __shfl_xor_sync(full, Bjj[q], o);
#pragma unroll
    for (int p = 0; p < 12; ++p) gl[p] += __shfl_xor_sync(full, gl[p], o);
  }
  if (on && k < 4 && add_fixed) {
    long long* Bd = B + (size_t)__ldg(uidx + find_slot(row_ptr, col, (int)j, (int)j)) * 16;
    // lane k flushes row k of the 4x4 factor and 3 gradient entries
    float b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f, g0 = 0.f, g1 = 0.f, g2 = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (c == k) {
        b0 = Bjj[4 * c]; b1 = Bjj[4 * c + 1]; b2 = Bjj[4 * c + 2]; b3 = Bjj[4 * c + 3];
        g0 = gl[3 * c]; g1 = gl[3 * c + 1]; g2 = gl[3 * c + 2];
      }
    fx_add(Bd + 4 * k, b0, fxs);
    fx_add(Bd + 4 * k + 1, b1, fxs);
    fx_add(Bd + 4 * k + 2, b2, fxs);
    fx_add(Bd + 4 * k + 3, b3, fxs);
    fx_add(rhs + (size_t)j * 12 + 3 * k, g0, fxs);
    fx_add(rhs + (size_t)j * 12 + 3 * k + 1, g1, fxs);
    fx_add(rhs + (size_t)j * 12 + 3 * k + 2, g2, fxs);
  }
}

// ------------------------------------------------------------------ preconditioner
// 12x12 block-Jacobi preconditioner: 16 lanes per node, lane k holds row k of the diagonal block
// D_j = kron(B_jj, I3) + Rot_j + lambda I in float64 registers and the group inverts it in place by
// Gauss-Jordan elimination (D is symmetric positive definite: no pivoting), exchanging the pivot
// row with shuffles.
__device__ __forceinline__ double shfl_d(double v, int src) {
  return __shfl_sync(0xffffffffu, v, src);
}

// The diagonal factor is read from the fixed-point system and converted exactly as k_sys_convert
// does (same float value), so the kernel does not wait for the conversion pass: the two run side by
// side.
__global__ void __launch_bounds__(128)
    k_precond(const int* __restrict__ row_ptr, const int* __restrict__ col,
              const int* __restrict__ uidx, const long long* __restrict__ U, float inv_scale,
              const float* __restrict__ Rot, int64_t m, double lambda, float* __restrict__ Minv,
              const float* __restrict__ Bf) {
  int64_t j = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / kNodeLanes;
  int lane = threadIdx.x & 31;
  int k = lane & (kNodeLanes - 1);
  int gbase = lane & ~(kNodeLanes - 1);
  bool on = j < m;
  bool act = on && k < 12;
  double a[12];
  double diag0 = 1.0;
#pragma unroll
  for (int q = 0; q < 12; ++q) a[q] = (q == k) ? 1.0 : 0.0;  // identity rows for idle lanes
  if (act) {
    const int dslot = find_slot(row_ptr, col, (int)j, (int)j);
    const long long* Bd = U + (size_t)__ldg(uidx + dslot) * 16;
    const float* Bdf = Bf ? Bf + (size_t)dslot * 16 : nullptr;  // the converted system (same float values)
    int ka = k / 3, kr = k - 3 * ka;
#pragma unroll
    for (int q = 0; q < 12; ++q) {
      int qa = q / 3, qr = q - 3 * qa;
      double v = 0.0;
      if (qr == kr)
        v = Bdf ? (double)__ldg(Bdf + 4 * ka + qa)
                : (double)(float)((double)__ldg(Bd + 4 * ka + qa) * (double)inv_scale);
      if (k < 9 && q < 9) v += (double)__ldg(Rot + (size_t)j * 81 + 9 * k + q);
      if (q == k) v += lambda;
      a[q] = v;
    }
#pragma unroll
    for (int q = 0; q < 12; ++q)
      if (q == k) diag0 = a[q];
  }
  bool bad = false;
#pragma unroll
  for (int c = 0; c < 12; ++c) {
    double piv = shfl_d(a[c], gbase + c);
    if (!(piv > 0.0)) bad = true;
    double d = 1.0 / piv;
    double f = a[c];  // this row's entry in the pivot column
    double prow[12];
#pragma unroll
    for (int q = 0; q < 12; ++q) prow[q] = shfl_d(a[q], gbase + c) * d;  // scaled pivot row
    if (k == c) {
#pragma unroll
      for (int q = 0; q < 12; ++q) a[q] = (q == c) ? d : prow[q];
    } else {
#pragma unroll
      for (int q = 0; q < 12; ++q) a[q] = (q == c) ? -f * d : a[q] - f * prow[q];
    }
  }
  // every lane of a node saw the same pivots, so `bad` is uniform over the node's 16 lanes:
  // fall back to the inverse of the diagonal (Jacobi) for that node
  if (act) {
    float* out = Minv + (size_t)j * 144 + 12 * k;
#pragma unroll
    for (int q = 0; q < 12; ++q)
      out[q] = bad ? ((q == k && diag0 > 0.0) ? (float)(1.0 / diag0) : 0.f) : (float)a[q];
  }
}

// ------------------------------------------------------------------ all-reduce + conversion over peer memory
// Vertex-sharded solve over NVLink (comm.cuh), reads only — every transfer is a coalesced peer-to-peer
// LOAD of contiguous memory:
//   1. reduce-scatter: rank r owns the slice [lo, hi) of the fixed-point words; it loads that slice of
//      every rank's partial system (16 bytes per lane), adds — integers: the total does not depend on
//      the order — and writes the float value (the conversion k_sys_convert applies) into its own
//      slice of `redf`;
//   2. all-gather: after a barrier every rank copies the other ranks' slices of `redf` into its own;
//   3. the float system [B | rhs] is expanded locally from `redf` (lower triangle mirrored).
// Per rank and iteration (N-1)/N x (8 + 4) bytes per fixed-point word cross NVLink, against
// 2 (N-1)/N x 8 for ncclAllReduce(int64) followed by a conversion pass.
__global__ void __launch_bounds__(256)
    k_peer_reduce(PeerLinkDev L, int64_t lo, int64_t hi, float inv_scale) {
  // lo, hi are even: two words per lane
  int64_t i = lo + 2 * (blockIdx.x * (int64_t)blockDim.x + threadIdx.x);
  const int64_t stride = 2 * (int64_t)gridDim.x * blockDim.x;
  float* out = L.redf[L.rank];
  for (; i < hi; i += stride) {
    long long s0 = 0, s1 = 0;
    for (int r = 0; r < L.nranks; ++r) {
      const longlong2 v = __ldcg(reinterpret_cast<const longlong2*>(L.sys64[r] + i));
      s0 += v.x;
      s1 += v.y;
    }
    *reinterpret_cast<float2*>(out + i) =
        make_float2((float)((double)s0 * (double)inv_scale), (float)((double)s1 * (double)inv_scale));
  }
}

__global__ void __launch_bounds__(256)
    k_peer_gather(PeerLinkDev L, int64_t per, int64_t n) {
  // words [r per, (r + 1) per) belong to rank r; per is a multiple of 4
  int64_t i = 4 * (blockIdx.x * (int64_t)blockDim.x + threadIdx.x);
  const int64_t stride = 4 * (int64_t)gridDim.x * blockDim.x;
  float* out = L.redf[L.rank];
  for (; i < n; i += stride) {
    const int owner = (int)(i / per);
    if (owner == L.rank) continue;
    *reinterpret_cast<float4*>(out + i) = __ldcg(reinterpret_cast<const float4*>(L.redf[owner] + i));
  }
}

// [U: nup x 16 | rhs] in float -> float system [B: nnz x 16 | rhs] (k_sys_convert's index map)
__global__ void k_sys_expand(const float* __restrict__ in, float* __restrict__ out, int64_t n,
                             int64_t nnz16, int64_t nup16, const int* __restrict__ tslot,
                             const int* __restrict__ uidx) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    int64_t src;
    if (i < nnz16) {
      int e = (int)(i >> 4), el = (int)(i & 15);
      bool lower = __ldg(tslot + e) < e;
      src = (int64_t)__ldg(uidx + e) * 16 + (lower ? (((el & 3) << 2) | (el >> 2)) : el);
    } else {
      src = nup16 + (i - nnz16);
    }
    out[i] = __ldg(in + src);
  }
}

// Below this many vertices the node terms and the preconditioner run on the main stream: on a view
// mesh (35K vertices) the kernels they would overlap last a few microseconds, while every fork/join
// is four driver calls (8 per GN iteration) that the workers of mvs_stitch_pairs contend for.
constexpr int64_t kSideStreamMinVerts = 200000;

// fixed-point scale: the largest power of two that cannot overflow int64 for this problem (every
// entry is bounded by w_fit omega_max n max(1, D^2) plus the node terms)
static float gn_fx_scale(mvs_deform* h) {
  double bound = 4.0 * (h->prm.w_fit * (double)h->omega_max * (double)h->n * std::max(1.0, h->diag2) +
                        h->prm.w_reg * 64.0 * std::max(1.0, h->diag2) + h->prm.w_rot * 64.0 + 1.0);
  int ex = 0;
  frexp(bound, &ex);  // bound < 2^ex
  h->fx_scale = ldexp(1.0, std::max(0, std::min(40, 62 - ex)));
  return (float)h->fx_scale;
}

static int gn_zero_system(mvs_deform* h, cudaStream_t st) {
  // the word after the system is the assembly's group counter: cleared by the same memset
  MVS_CUDA(cudaMemsetAsync(h->sys64.p, 0, sizeof(long long) * (h->sys64_count() + 1), st));
  return MVS_OK;
}

static int gn_launch_node_terms(mvs_deform* h, float fxs, cudaStream_t st, int add_fixed = 1) {
  ProfScope ps(h->ctx, "gn_node_terms", 1, st);
  k_node_terms<<<div_up(h->m * kNodeLanes, 128), 128, 0, st>>>(
      h->g.p, h->x.p, h->graph.p, h->graph_k + 1, h->m, h->row_ptr.p, h->col.p, h->uidx.p, h->B64(),
      h->Rot(), h->rhs64(), (float)h->prm.w_rot, (float)h->prm.w_reg, fxs, add_fixed);
  MVS_CUDA(cudaGetLastError());
  return MVS_OK;
}

// The node terms (rotation + smoothness) depend only on the node state, and their fixed-point
// atomics commute with the fit pass's: on one rank they are zeroed-and-accumulated on the side
// stream while the main stream runs the correspondence pass, and joined before the fit pass.
// (Vertex-sharded solves add them once, after the all-reduce, on the main stream.)
int gn_node_terms_forked(mvs_deform* h) {
  mvs_ctx* ctx = h->ctx;
  h->node_terms_forked = false;
  if (ctx->comm && ctx->nranks > 1) return MVS_OK;
  if (h->n < kSideStreamMinVerts) return MVS_OK;  // small mesh: the fork/join calls cost more than they hide
  float fxs = gn_fx_scale(h);
  MVS_TRY(ctx_side_fork(ctx));
  MVS_TRY(gn_zero_system(h, ctx->side));
  MVS_TRY(gn_launch_node_terms(h, fxs, ctx->side));
  h->node_terms_forked = true;
  return MVS_OK;
}

int gn_assemble(mvs_deform* h) {
  mvs_ctx* ctx = h->ctx;
  cudaStream_t st = ctx->stream;
  float fxs = gn_fx_scale(h);
  const bool forked = h->node_terms_forked;
  h->node_terms_forked = false;
  if (forked)
    MVS_TRY(ctx_side_join(ctx));
  else
    MVS_TRY(gn_zero_system(h, st));
  unsigned* group_counter = reinterpret_cast<unsigned*>(h->sys64.p + h->sys64_count());
  bool counter_clean = true;  // zeroed together with the system
  for (const ShardRanges& R : range_batches(h)) {
    const int64_t ngr = R.goff[R.n], nv = R.voff[R.n];
    if (ngr <= 0 || nv <= 0) continue;
    if (!counter_clean) MVS_CUDA(cudaMemsetAsync(group_counter, 0, sizeof(long long), st));
    counter_clean = false;
    ProfScope ps(ctx, "gn_assemble_fit");
    int64_t warps_needed = std::min<int64_t>(ngr, (nv + 63) / 64);
    int grid = (int)std::min<int64_t>((warps_needed + kAsmWarps - 1) / kAsmWarps,
                                      (int64_t)ctx->num_sms * kAsmCtasPerSm);
    k_assemble_fit<<<grid, kAsmWarps * 32, 0, st>>>(
        h->v.p, h->vw.p, h->resid.p, h->run_g.p, h->run_local.p, h->run_start.p, h->grp_n.p,
        h->grp_dst.p, h->sys64.p, (int)(h->nup * 16), (int)h->nruns, R, (float)h->prm.w_fit, fxs,
        group_counter, ctx->solver_flag);
  }
  MVS_CUDA(cudaGetLastError());
  const float inv_scale = (float)(1.0 / h->fx_scale);
  if (ctx->comm && ctx->nranks > 1 && h->peer && !h->peer->ready && !h->peer->failed)
    MVS_TRY(peer_link_connect(ctx, h->peer));  // collective; the ranks agree on ready / failed
  if (ctx->comm && ctx->nranks > 1 && h->peer && h->peer->ready) {
    // vertex-sharded solve over peer memory: the node terms enter the sum once (rank 0; every rank
    // still needs the rotation blocks, which are node-local), then barrier -> fused reduce + convert
    // -> barrier, and the preconditioner reads the converted float system
    MVS_TRY(gn_launch_node_terms(h, fxs, st, ctx->rank == 0 ? 1 : 0));
    {
      ProfScope ps(ctx, "gn_peer_barrier_in");
      MVS_TRY(peer_barrier(ctx, h->peer));  // every rank's partial system is complete
    }
    const int64_t n64 = (int64_t)h->sys64_count();
    // slice per rank: a multiple of 4 words (16-byte lanes in both kernels)
    const int64_t per = ((n64 + ctx->nranks - 1) / ctx->nranks + 3) / 4 * 4;
    const int64_t lo = std::min<int64_t>(n64, per * ctx->rank), hi = std::min<int64_t>(n64, lo + per);
    {
      ProfScope ps(ctx, "gn_peer_reduce");
      if (hi > lo)
        k_peer_reduce<<<grid_for(ctx, (hi - lo + 1) / 2, 256, 8), 256, 0, st>>>(h->peer->dev, lo, hi,
                                                                              inv_scale);
    }
    {
      ProfScope ps(ctx, "gn_peer_barrier_mid");
      MVS_TRY(peer_barrier(ctx, h->peer));  // every slice is reduced; nobody reads the partials any more
    }
    {
      ProfScope ps(ctx, "gn_peer_gather", 2);
      k_peer_gather<<<grid_for(ctx, (n64 + 3) / 4, 256, 8), 256, 0, st>>>(h->peer->dev, per, n64);
      const int64_t nout = h->nnz * 16 + h->m * 12;
      k_sys_expand<<<grid_for(ctx, nout, 256, 4), 256, 0, st>>>(
          h->peer->dev.redf[ctx->rank], h->sys.p, nout, h->nnz * 16, h->nup * 16, h->tslot.p,
          h->uidx.p);
    }
    {
      ProfScope ps(ctx, "gn_precond");
      k_precond<<<div_up(h->m * kNodeLanes, 128), 128, 0, st>>>(
          h->row_ptr.p, h->col.p, h->uidx.p, h->B64(), inv_scale, h->Rot(), h->m, h->prm.lambda,
          h->Minv.p, h->B());
    }
    MVS_CUDA(cudaGetLastError());
    return MVS_OK;
  }
  if (ctx->comm && ctx->nranks > 1) {
    // vertex-sharded solve: sum the partial J^T J blocks and J^T r over ranks (NVLink/NVSwitch);
    // integer sums make the result independent of the reduction order NCCL picks
    ProfScope ps(ctx, "gn_allreduce", 0);
    MVS_TRY(comm_allreduce_sum_i64(ctx, h->sys64.p, h->sys64_count()));
  }
  if (!forked) MVS_TRY(gn_launch_node_terms(h, fxs, st));
  // conversion (main stream) and block-Jacobi inverses (side stream) both read the finished
  // fixed-point system
  const bool side = h->n >= kSideStreamMinVerts;
  cudaStream_t pst = side ? ctx->side : st;
  if (side) MVS_TRY(ctx_side_fork(ctx));
  {
    ProfScope ps(ctx, "gn_precond", 1, pst);
    k_precond<<<div_up(h->m * kNodeLanes, 128), 128, 0, pst>>>(
        h->row_ptr.p, h->col.p, h->uidx.p, h->B64(), inv_scale, h->Rot(), h->m, h->prm.lambda,
        h->Minv.p, nullptr);
  }
  {
    ProfScope ps(ctx, "gn_sys_convert");
    // sys = [B | rhs | Rot]: B and rhs are contiguous in both buffers
    int64_t nout = h->nnz * 16 + h->m * 12;
    k_sys_convert<<<grid_for(ctx, nout, 256, 4), 256, 0, st>>>(
        h->sys64.p, h->sys.p, nout, h->nnz * 16, h->nup * 16, h->tslot.p, h->uidx.p, inv_scale);
  }
  MVS_CUDA(cudaGetLastError());
  if (side) MVS_TRY(ctx_side_join(ctx));
  return MVS_OK;
}

// ------------------------------------------------------------------ final vertices + normals
__global__ void k_deform_final(const float4* __restrict__ v, const int4* __restrict__ vidx,
                               const float4* __restrict__ vw, const float4* __restrict__ g,
                               const float* __restrict__ x, const int* __restrict__ perm,
                               int64_t n, float4* __restrict__ cur0) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 vi = __ldg(v + i);
  float3 q = deform_point(vi, __ldg(vidx + i), __ldg(vw + i), g, x);
  cur0[__ldg(perm + i)] = make_float4(q.x, q.y, q.z, 0.f);
}

// Deformation/Deformation.h:111-128 FacetNormal + :86-109 VertexNormal
__global__ void k_facet_normals(const float4* __restrict__ p, const int* __restrict__ facets,
                                int64_t nf, float* __restrict__ nacc /* n x 4 */) {
  int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (f >= nf) return;
  int a = facets[3 * f], b = facets[3 * f + 1], c = facets[3 * f + 2];
  float4 p1 = __ldg(p + a), p2 = __ldg(p + b), p3 = __ldg(p + c);
  float ux = p2.x - p1.x, uy = p2.y - p1.y, uz = p2.z - p1.z;
  float vx = p3.x - p1.x, vy = p3.y - p1.y, vz = p3.z - p1.z;
  float cx = uy * vz - uz * vy, cy = uz * vx - ux * vz, cz = ux * vy - uy * vx;
  float l = sqrtf(cx * cx + cy * cy + cz * cz);
  if (!(l > 0.f)) return;
  float il = 1.0f / l;
  cx *= il; cy *= il; cz *= il;
  int vs[3] = {a, b, c};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    atomicAdd(nacc + 4 * (size_t)vs[k], cx);
    atomicAdd(nacc + 4 * (size_t)vs[k] + 1, cy);
    atomicAdd(nacc + 4 * (size_t)vs[k] + 2, cz);
  }
}

__global__ void k_normalize(float4* __restrict__ nrm, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 q = nrm[i];
  float l = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z);
  if (l > 0.f) nrm[i] = make_float4(q.x / l, q.y / l, q.z / l, 0.f);
}

int gn_final_vertices(mvs_deform* h) {
  mvs_ctx* ctx = h->ctx;
  cudaStream_t st = ctx->stream;
  int64_t n = h->n;
  {
    ProfScope ps(ctx, "deform_final");
    k_deform_final<<<div_up(n, 256), 256, 0, st>>>(h->v.p, h->vidx.p, h->vw.p, h->g.p, h->x.p,
                                                   h->perm.p, n, h->cur0.p);
  }
  return gn_recompute_normals(h);
}

int gn_recompute_normals(mvs_deform* h) {
  mvs_ctx* ctx = h->ctx;
  cudaStream_t st = ctx->stream;
  int64_t n = h->n;
  MVS_CUDA(cudaMemsetAsync(h->curn0.p, 0, sizeof(float4) * n, st));
  if (h->nf > 0) {
    ProfScope ps(ctx, "vertex_normals", 2);
    k_facet_normals<<<div_up(h->nf, 256), 256, 0, st>>>(h->cur0.p, h->facets.p, h->nf,
                                                        reinterpret_cast<float*>(h->curn0.p));
    k_normalize<<<div_up(n, 256), 256, 0, st>>>(h->curn0.p, n);
  }
  MVS_CUDA(cudaGetLastError());
  h->cur_valid = true;
  return MVS_OK;
}

// ------------------------------------------------------------------ energies
__global__ void __launch_bounds__(256)
    k_energy_fit(const float4* __restrict__ resid, int64_t begin, int64_t end, double w_fit,
                 double* partials, unsigned* ticket, double* out) {
  __shared__ double smem[32];
  double v[1] = {0.0};
  int64_t i = begin + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < end; i += stride) {
    float4 r = __ldg(resid + i);
    if (r.w > 0.f)
      v[0] += w_fit * (double)r.w * ((double)r.x * r.x + (double)r.y * r.y + (double)r.z * r.z);
  }
  grid_reduce_sum<1>(v, smem, partials, ticket, out);
}

__global__ void __launch_bounds__(128)
    k_energy_nodes(const float4* __restrict__ g, const float* __restrict__ x,
                   const int* __restrict__ graph, int gk1, int64_t m, double w_rot, double w_reg,
                   double* partials, unsigned* ticket, double* out) {
  __shared__ double smem[64];
  double v[2] = {0.0, 0.0};
  int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; j < m; j += stride) {
    NodeX n = load_node(x, (int)j);
    double a[9];
    for (int k = 0; k < 9; ++k) a[k] = n.a[k];
    double r0 = a[0] * a[3] + a[1] * a[4] + a[2] * a[5];
    double r1 = a[0] * a[6] + a[1] * a[7] + a[2] * a[8];
    double r2 = a[3] * a[6] + a[4] * a[7] + a[5] * a[8];
    double r3 = a[0] * a[0] + a[1] * a[1] + a[2] * a[2] - 1.0;
    double r4 = a[3] * a[3] + a[4] * a[4] + a[5] * a[5] - 1.0;
    double r5 = a[6] * a[6] + a[7] * a[7] + a[8] * a[8] - 1.0;
    v[0] += w_rot * (r0 * r0 + r1 * r1 + r2 * r2 + r3 * r3 + r4 * r4 + r5 * r5);
    float4 gj = __ldg(g + j);
    for (int e = 0; e < gk1; ++e) {
      int l = __ldg(graph + (size_t)j * gk1 + e);
      if (l < 0 || l == j) continue;
      float4 gv = __ldg(g + l);
      NodeX nl = load_node(x, l);
      double ev[3] = {(double)gv.x - gj.x, (double)gv.y - gj.y, (double)gv.z - gj.z};
      for (int rr = 0; rr < 3; ++rr) {
        double r = (a[rr] * ev[0] + a[3 + rr] * ev[1] + a[6 + rr] * ev[2] - ev[rr]) +
                   ((double)n.a[9 + rr] - (double)nl.a[9 + rr]);
        v[1] += w_reg * r * r;
      }
    }
  }
  grid_reduce_sum<2>(v, smem, partials, ticket, out);
}

int gn_energy(mvs_deform* h, double energy[3]) {
  mvs_ctx* ctx = h->ctx;
  cudaStream_t st = ctx->stream;
  MVS_CUDA(cudaMemsetAsync(h->scal.p + S_EFIT, 0, sizeof(double) * 3, st));
  const auto ranges = h->shard_ranges();
  double efit_extra = 0.0;  // sub-ranges beyond the first: summed on the host (not a hot path)
  for (size_t k = 0; k < ranges.size(); ++k) {
    const int64_t b = ranges[k].first, e = ranges[k].second;
    if (k > 0) {
      double part = 0.0;
      MVS_TRY(d2h_sync(ctx, &part, h->scal.p + S_EFIT, sizeof(double)));
      efit_extra += part;
    }
    ProfScope ps(ctx, "energy_fit");
    int grid = grid_for(ctx, e - b, 256, 4);
    MVS_TRY(h->partials.reserve((size_t)grid * 2 + 16));
    k_energy_fit<<<grid, 256, 0, st>>>(h->resid.p, b, e, h->prm.w_fit, h->partials.p, h->ticket.p,
                                       h->scal.p + S_EFIT);
  }
  if (efit_extra != 0.0) {
    double last = 0.0;
    MVS_TRY(d2h_sync(ctx, &last, h->scal.p + S_EFIT, sizeof(double)));
    last += efit_extra;
    MVS_TRY(h2d(ctx, h->scal.p + S_EFIT, &last, sizeof(double)));
    MVS_CUDA(cudaStreamSynchronize(st));
  }
  if (ctx->comm && ctx->nranks > 1) MVS_TRY(comm_allreduce_sum_f64(ctx, h->scal.p + S_EFIT, 1));
  {
    ProfScope ps(ctx, "energy_nodes");
    int grid = grid_for(ctx, h->m, 128, 2);
    MVS_TRY(h->partials.reserve((size_t)grid * 2 + 16));
    k_energy_nodes<<<grid, 128, 0, st>>>(h->g.p, h->x.p, h->graph.p, h->graph_k + 1, h->m,
                                         h->prm.w_rot, h->prm.w_reg, h->partials.p, h->ticket.p,
                                         h->scal.p + S_EROT);
  }
  MVS_CUDA(cudaGetLastError());
  MVS_TRY(d2h_sync(ctx, energy, h->scal.p + S_EFIT, sizeof(double) * 3));
  return MVS_OK;
}

}  // namespace mvs
