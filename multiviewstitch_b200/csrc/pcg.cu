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
//   * two grid-wide exchanges per iteration (after p.q and after r.z).  The barrier and the dot
//     product are ONE step: every CTA release-stores its float64 partial into its slot of a small
//     ring, and every CTA polls all slots until none holds the "not yet written" sentinel — one
//     store-to-visible latency instead of arrive-atomic + spin + partial read-back.  The reduction
//     order is fixed (lane -> warp -> CTA partial -> every CTA adds the partials in CTA order), so
//     the solve is bit-reproducible run to run;
//   * the sparse product is gathered entry-per-lane: lane k of a row takes pattern entries k, k+16
//     of that row, loads the whole neighbour record (12 x (z, p) = six 16-byte loads) — all loads of
//     a row are in flight at once, one L2 round trip — and a 15-shuffle reduce-scatter hands every
//     lane its own component;
//   * the search direction of a neighbour row is rebuilt on the fly as z_l + beta p_prev_l, which
//     saves the third barrier a separate "p = z + beta p" phase would need.
// Standard PCG recurrences, identical to oracle/deform.cpp pcg(); vectors float32, scalars float64.
#include <stdlib.h>

#include <algorithm>

#include "deform.cuh"
#include "reduce.cuh"

namespace mvs {

constexpr int kRowLanes = 16;
// CTA size is a launch parameter (256: two CTAs per SM, 512: one): fewer, larger CTAs halve the
// arrivals of every grid exchange.  MVS_PCG_THREADS overrides the planner's choice.
constexpr int kPcgMaxThreads = 512;
constexpr int kPcgMaxWarps = kPcgMaxThreads / 32;
constexpr int kPcgDefaultThreads = 512;  // C3: 149.1 us per 20 iterations against 151 at 256

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
  double* ring;   // kRing x grid x 2 exchange slots, all-ones at launch
  unsigned long long* words;  // (2 iters + 1) x 4 fixed-point exchange words, zero at launch
  // the other half of the exchange buffer (used by the previous launch): this launch re-arms it for
  // the next one, so no memset sits between two solves
  unsigned long long* next_ring;
  unsigned long long* next_words;
  int ring_count, word_count;
  double* scal;   // S_* slots
  int64_t m;
  int iters;
  float lambda;
  double tol;
  int rows_per_cta;
  int stage_smem;  // 1: matrix slice staged in shared memory
  int* diverged;   // pinned host word raised when a partial sum leaves the representable range
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

constexpr unsigned long long kSlotEmpty = ~0ull;  // bit pattern no sum can produce (see below)
constexpr int kRing = 4;                           // phases of slots in flight (>= 3 required)

__device__ __forceinline__ void st_flag(double* p, double v) {
  unsigned long long b = (unsigned long long)__double_as_longlong(v);
  if (b == kSlotEmpty) b = 0x7ff8000000000000ull;  // a NaN stays a NaN, never the sentinel
  asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" ::"l"(p), "l"(b) : "memory");
}
__device__ __forceinline__ void fence_acq_rel() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ unsigned long long ld_word(const unsigned long long* p) {
  unsigned long long b;
  asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(b) : "l"(p) : "memory");
  return b;
}
__device__ __forceinline__ unsigned long long ld_flag(const double* p) {
  unsigned long long b;
  asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(b) : "l"(p) : "memory");
  return b;
}

// Grid-wide barrier + all-reduce of N doubles in one exchange.
//   ring: kRing x ncta x 2 doubles, all kSlotEmpty at launch.  Phase ph uses ring[ph % kRing].
//   CTA c: (1) fixed-order CTA sum; (2) thread 0 re-arms its slot of phase ph-2 (every CTA consumed
//   that phase before it could write phase ph-1, which this CTA has already seen complete), fences
//   (publishing the vector entries the CTA wrote before the call, ordered by the __syncthreads) and
//   stores the partial; (3) warp 0 polls all slots — 8 independent relaxed loads per lane in flight,
//   no L1 invalidation inside the spin — then fences once (acquire) and adds the partials in a
//   fixed order; (4) the totals reach the other warps through shared memory.  Consumers read the
//   exchanged vectors with ld.global.cg (L2), never through L1.
// Every CTA runs the same number of phases (control flow depends only on the identical totals).
template <int N, int T>
__device__ __forceinline__ void grid_exchange(double (&v)[N], double* red, double* ring,
                                              unsigned& phase, int ncta, double (&out)[N]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int W = T / 32;
  constexpr int U = 8;           // slots per lane in flight
  double* red1 = red;                       // W x 2: CTA partial
  double* red2 = red + kPcgMaxWarps * 2;    // 2: totals
#pragma unroll
  for (int j = 0; j < N; ++j) v[j] = warp_sum(v[j]);
  if (lane == 0)
#pragma unroll
    for (int j = 0; j < N; ++j) red1[warp * 2 + j] = v[j];
  __syncthreads();
  double* cur = ring + (size_t)(phase % kRing) * ncta * 2;
  if (warp == 0) {
    if (lane == 0) {
      double* old = ring + ((size_t)((phase + 2) % kRing) * ncta + blockIdx.x) * 2;
      reinterpret_cast<unsigned long long*>(old)[0] = kSlotEmpty;
      reinterpret_cast<unsigned long long*>(old)[1] = kSlotEmpty;
      double s[N];
#pragma unroll
      for (int j = 0; j < N; ++j) {
        s[j] = 0.0;
        for (int w = 0; w < W; ++w) s[j] += red1[w * 2 + j];
      }
      fence_acq_rel();
#pragma unroll
      for (int j = 0; j < N; ++j) st_flag(cur + (size_t)blockIdx.x * 2 + j, s[j]);
    }
    __syncwarp();
    double t[N];
#pragma unroll
    for (int j = 0; j < N; ++j) t[j] = 0.0;
    for (int c0 = 0; c0 < ncta; c0 += 32 * U) {
      unsigned long long b[U][N];
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int j = 0; j < N; ++j) b[u][j] = (c0 + u * 32 + lane < ncta) ? kSlotEmpty : 0ull;
      unsigned spins = 0;
      for (;;) {
        bool ok = true;
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int j = 0; j < N; ++j)
            if (b[u][j] == kSlotEmpty) b[u][j] = ld_flag(cur + (size_t)(c0 + u * 32 + lane) * 2 + j);
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int j = 0; j < N; ++j) ok = ok && (b[u][j] != kSlotEmpty);
        if (ok) break;
        if (++spins > (1u << 26)) __trap();  // a lost CTA becomes an error, not a hang
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int j = 0; j < N; ++j) t[j] += __longlong_as_double((long long)b[u][j]);
    }
    fence_acq_rel();
#pragma unroll
    for (int j = 0; j < N; ++j) t[j] = warp_sum(t[j]);
    if (lane == 0)
#pragma unroll
      for (int j = 0; j < N; ++j) red2[j] = t[j];
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < N; ++j) out[j] = red2[j];
  ++phase;
}

// ---- fast exchange: integer atomics on one word per (scalar, limb)
// Measured on B200 (scratch/barrier_bench.cu, 256 CTAs x 256 threads, us per barrier + all-reduce of
// two doubles): cooperative grid.sync + partial read-back 5.2, polled slots 5.3, arrival counter +
// slot read 3.0, THIS scheme 2.2.  Each CTA partial is split into two signed 44-bit limbs of a fixed-
// point number with unit 2^E (E fixed per quantity at the start of the solve, 19 bits of headroom
// above and 67 bits below the initial magnitude); a limb is biased to be non-negative and added with
// ONE 64-bit atomic whose top 10 bits count arrivals, so the same word is the barrier flag and the
// sum.  Integer addition commutes: the total is independent of arrival order (bit-reproducible), and
// every CTA decodes the same value.  Words are zero at launch and used once.
constexpr int kLimbBits = 44;
constexpr long long kLimbBias = 1ll << 43;
constexpr long long kLimbMax = (1ll << 43) - 1;
constexpr int kCountShift = 54;

struct FxUnit {
  double lo, hi, inv_lo, inv_hi;  // 2^E, 2^(E+44) and their reciprocals
};
__device__ __forceinline__ FxUnit fx_unit_for(double magnitude) {
  int x = (magnitude > 0.0 && magnitude < INFINITY) ? ilogb(magnitude) : 0;
  int E = min(max(x - 67, -900), 800);
  FxUnit u;
  u.lo = ldexp(1.0, E);
  u.hi = ldexp(1.0, E + kLimbBits);
  u.inv_lo = ldexp(1.0, -E);
  u.inv_hi = ldexp(1.0, -(E + kLimbBits));
  return u;
}
__device__ __forceinline__ long long fx_limb(double v, const FxUnit& u, int limb, int* flag) {
  // a non-finite partial contributes nothing and one beyond the fixed-point range is clamped: either
  // way the solve is diverging — raise the context's flag so the host call returns an error
  if (!(fabs(v) < INFINITY)) {
    v = 0.0;
    if (flag) *reinterpret_cast<volatile int*>(flag) = 1;
  }
  double hr = rint(v * u.inv_hi);
  if (flag && fabs(hr) > (double)kLimbMax) *reinterpret_cast<volatile int*>(flag) = 2;
  double h = fmin(fmax(hr, (double)-kLimbMax), (double)kLimbMax);
  if (limb == 0) return (long long)h;
  double rem = v - h * u.hi;  // exact
  double l = fmin(fmax(rint(rem * u.inv_lo), (double)-kLimbMax), (double)kLimbMax);
  return (long long)l;
}

// FENCE = false: a pure all-reduce.  The p.q exchange needs no memory ordering — the search
// directions a CTA stored before it are read by other CTAs only after the NEXT (fenced) exchange —
// so it skips the release/acquire fences (each waits for the CTA's outstanding stores).
// `post(tot, extra)` runs on ONE thread of the CTA between the decode and the closing barrier and
// leaves derived scalars (the step lengths: one float64 division each) in extra[0..1], which every
// thread reads after the barrier — instead of every thread issuing its own float64 division.
template <int N, int T, bool FENCE = true, class Post>
__device__ __forceinline__ void grid_exchange_fx(double (&v)[N], double* red,
                                                 unsigned long long* words, unsigned& wi, int ncta,
                                                 const FxUnit (&unit)[N], double (&out)[N],
                                                 Post&& post, int* flag = nullptr) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int W = T / 32;
  double* red1 = red;                     // W x 2: CTA partial
  double* red2 = red + kPcgMaxWarps * 2;  // N x 2 limb sums
#pragma unroll
  for (int j = 0; j < N; ++j) v[j] = warp_sum(v[j]);
  if (lane == 0)
#pragma unroll
    for (int j = 0; j < N; ++j) red1[warp * 2 + j] = v[j];
  __syncthreads();
  if (threadIdx.x < 2 * N) {
    const int j = threadIdx.x >> 1, limb = threadIdx.x & 1;
    double s = 0.0;
    for (int w = 0; w < W; ++w) s += red1[w * 2 + j];
    FxUnit u = unit[0];
    if (N > 1 && j == 1) u = unit[N - 1];
    long long fxv = fx_limb(s, u, limb, flag);
    unsigned long long enc = (unsigned long long)(fxv + kLimbBias) + (1ull << kCountShift);
    if (FENCE && threadIdx.x == 0) fence_acq_rel();  // publish what this CTA wrote before the call
    __syncwarp((1u << (2 * N)) - 1u);
    unsigned long long* wd = words + wi + threadIdx.x;
    atomicAdd(wd, enc);
    unsigned long long b;
    unsigned spins = 0;
    while (((b = ld_word(wd)) >> kCountShift) < (unsigned long long)ncta)
      if (++spins > (1u << 26)) __trap();  // a lost CTA becomes an error, not a hang
    if (FENCE && threadIdx.x == 0) fence_acq_rel();
    long long sum = (long long)(b & ((1ull << kCountShift) - 1ull)) - (long long)ncta * kLimbBias;
    red2[threadIdx.x] = (double)sum;
    __syncwarp((1u << (2 * N)) - 1u);
    if (threadIdx.x == 0) {
      double tot[N];
#pragma unroll
      for (int j = 0; j < N; ++j) tot[j] = red2[2 * j] * unit[j].hi + red2[2 * j + 1] * unit[j].lo;
      post(tot, red2 + 2 * N);
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < N; ++j) out[j] = red2[2 * j] * unit[j].hi + red2[2 * j + 1] * unit[j].lo;
  wi += 2 * N;
}

// y[0..11] partial products of this lane -> lane k of the 16-lane group returns sum over the group
// of y[k] (k < 12).  Reduce-scatter by halving: 8 + 4 + 2 + 1 shuffles.
__device__ __forceinline__ float group_reduce_scatter(const float (&y)[12], int k) {
  const unsigned full = 0xffffffffu;
  float a8[8], a4[4], a2[2];
  const bool u8 = k & 8, u4 = k & 4, u2 = k & 2, u1 = k & 1;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float hi = i < 4 ? y[8 + i] : 0.f;
    float send = u8 ? y[i] : hi;
    float keep = u8 ? hi : y[i];
    a8[i] = keep + __shfl_xor_sync(full, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float send = u4 ? a8[i] : a8[4 + i];
    float keep = u4 ? a8[4 + i] : a8[i];
    a4[i] = keep + __shfl_xor_sync(full, send, 4);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    float send = u2 ? a4[i] : a4[2 + i];
    float keep = u2 ? a4[2 + i] : a4[i];
    a2[i] = keep + __shfl_xor_sync(full, send, 2);
  }
  float send = u1 ? a2[0] : a2[1];
  float keep = u1 ? a2[1] : a2[0];
  return keep + __shfl_xor_sync(full, send, 1);
}

// One pattern entry: acc[3a + r] += sum_b B[a][b] (z_l[3b + r] + beta p_l[3b + r]); `zq` are the six
// 16-byte pieces of the neighbour's 12 (z, p) pairs.
__device__ __forceinline__ void entry_mac(float (&acc)[12], const float4* __restrict__ Bt,
                                          const float4 (&zq)[6], float beta) {
  float pv[12];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    pv[2 * i] = zq[i].x + beta * zq[i].y;
    pv[2 * i + 1] = zq[i].z + beta * zq[i].w;
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    float4 b = Bt[a];
#pragma unroll
    for (int r = 0; r < 3; ++r)
      acc[3 * a + r] += b.x * pv[r] + b.y * pv[3 + r] + b.z * pv[6 + r] + b.w * pv[9 + r];
  }
}

// zp[buf][row][k] = (z, p): z of the row after the update of an iteration, p its search direction;
// iteration `it` reads buffer (it+1)&1 (written during iteration it-1) and writes buffer it&1.
template <int RS, int T>
__global__ void __launch_bounds__(T, 512 / T) k_pcg_persistent(PcgArgs a) {
  extern __shared__ float4 smem_dyn[];
  __shared__ double red[kPcgMaxWarps * 2 + 8];
  constexpr int kGroups = T / kRowLanes;  // row groups of this CTA
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
  float2* zp = reinterpret_cast<float2*>(a.zbuf);  // 2 x m x 12 float2
  unsigned phase = 0;
  for (int i = blockIdx.x * T + threadIdx.x; i < a.ring_count + a.word_count;
       i += gridDim.x * T) {
    if (i < a.ring_count)
      a.next_ring[i] = kSlotEmpty;
    else
      a.next_words[i - a.ring_count] = 0ull;
  }

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
    for (int i = threadIdx.x; i < nb; i += T) smem_dyn[i] = __ldg(Bsrc + i);
    for (int i = threadIdx.x; i < nz1 - nz0; i += T) scol[i] = __ldg(csrc + i);
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
    // the pattern range of the row is known to all 16 lanes (every lane gathers entries)
    t0s[s] = on ? __ldg(a.row_ptr + j) - nz0 : 0;
    t1s[s] = on ? __ldg(a.row_ptr + j + 1) - nz0 : 0;
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
  double rz, bn2;
  {
    double tot[2];
    grid_exchange<2, T>(acc, red, a.ring, phase, ncta, tot);
    rz = tot[0];
    bn2 = tot[1];
  }
  double rn2 = bn2;
  double rz_old = 0.0;
  unsigned wi = 0;
  const FxUnit unit_b[2] = {fx_unit_for(rz), fx_unit_for(bn2)};
  FxUnit unit_a[1] = {fx_unit_for(0.0)};
  int it = 0;
  float beta_next = 0.f;
  for (; it < a.iters; ++it) {
    if (a.tol > 0.0 && rn2 <= a.tol * a.tol * bn2) break;  // uniform over the grid
    // ---- phase A: p = z + beta p ; q = (H + lambda) p ; p.q
    const float beta = beta_next;  // rz / rz_old, divided once per CTA inside the last exchange
    const float4* zprev = reinterpret_cast<const float4*>(zp + (size_t)((it + 1) & 1) * m * 12);
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
      // gather: this lane's pattern entries, two entries (12 x 16 B) in flight
      float ya[12];
#pragma unroll
      for (int q = 0; q < 12; ++q) ya[q] = 0.f;
      const int t1 = t1s[s];
      // (one entry at a time when many row slots already fill the register file)
      constexpr bool kTwo = RS <= 2;
      for (int t = t0s[s] + k; t < t1; t += (kTwo ? 2 : 1) * kRowLanes) {
        const int tb = t + kRowLanes;
        const bool two = kTwo && tb < t1;
        const int la = cs[t], lb = two ? cs[tb] : la;
        float4 za[6], zb[6];
        const float4* sa = zprev + (size_t)la * 6;
        const float4* sb = zprev + (size_t)lb * 6;
#pragma unroll
        for (int i = 0; i < 6; ++i) za[i] = __ldcg(sa + i);
        if (two)
#pragma unroll
          for (int i = 0; i < 6; ++i) zb[i] = __ldcg(sb + i);
        entry_mac(ya, Bs + (size_t)t * 4, za, beta);
        if (two) entry_mac(ya, Bs + (size_t)tb * 4, zb, beta);
      }
      float y = group_reduce_scatter(ya, k);
      y += a.lambda * p[s];
#pragma unroll
      for (int q = 0; q < 9; ++q) {
        float pqv = __shfl_sync(full, p[s], gbase + q);
        if (RS == 1)
          y += rot[q] * pqv;
        else if (act && k < 9)
          y += __ldg(a.Rot + j * 81 + k * 9 + q) * pqv;
      }
      if (!act) y = 0.f;
      qv[s] = y;
      pq[0] += (double)p[s] * (double)y;
    }
    double pqa[1];
    float alpha;
    double* extra = red + kPcgMaxWarps * 2;
    if (it == 0) {  // magnitude of p.q unknown yet: exact slot exchange, then fix its unit
      grid_exchange<1, T>(pq, red, a.ring, phase, ncta, pqa);
      unit_a[0] = fx_unit_for(fabs(pqa[0]));
      alpha = (pqa[0] > 0.0 && rz > 0.0) ? (float)(rz / pqa[0]) : 0.f;
    } else {
      grid_exchange_fx<1, T, false>(pq, red, a.words, wi, ncta, unit_a, pqa,
                                 [&](const double (&tot)[1], double* ex) {
                                   ex[0] = (tot[0] > 0.0 && rz > 0.0) ? (double)(float)(rz / tot[0]) : 0.0;
                                 },
                                 a.diverged);
      alpha = (float)extra[2];
    }
    // ---- phase B: delta += alpha p ; r -= alpha q ; z = M^-1 r ; r.z, |r|^2
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
    rz_old = rz;
    {
      double tot[2];
      grid_exchange_fx<2, T>(acc, red, a.words, wi, ncta, unit_b, tot,
                          [&](const double (&t2)[2], double* ex) {
                            ex[0] = rz_old > 0.0 ? (double)(float)(t2[0] / rz_old) : 0.0;
                          },
                          a.diverged);
      rz = tot[0];
      rn2 = tot[1];
      beta_next = (float)extra[4];
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

template <int RS, int T>
static int launch_pcg_t(mvs_deform* h, PcgArgs& a, int grid, size_t smem) {
  auto kern = k_pcg_persistent<RS, T>;
  // the limit was raised once, by the planner (ensure_dyn_smem): no attribute call per launch
  void* args[] = {&a};
  MVS_CUDA(cudaLaunchCooperativeKernel((void*)kern, dim3(grid), dim3(T), args, smem, h->ctx->stream));
  return MVS_OK;
}
template <int RS>
static int launch_pcg(mvs_deform* h, PcgArgs& a, int grid, size_t smem) {
  return h->pcg_threads == 512 ? launch_pcg_t<RS, 512>(h, a, grid, smem)
                               : launch_pcg_t<RS, 256>(h, a, grid, smem);
}

template <int RS, int T>
static int max_coresident_t(size_t smem, int num_sms, int* out) {
  auto kern = k_pcg_persistent<RS, T>;
  if (smem > 48 * 1024) {
    int dev = 0;
    cudaGetDevice(&dev);
    // raised once to the planner's ceiling (200 KB), not per candidate size
    if (ensure_dyn_smem((const void*)kern, dev, 200 * 1024) != MVS_OK) {
      cudaGetLastError();
      *out = 0;
      return MVS_OK;
    }
  }
  int per_sm = 0;
  MVS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, T, smem));
  *out = per_sm * num_sms;
  return MVS_OK;
}
template <int RS>
static int max_coresident(size_t smem, int num_sms, int threads, int* out) {
  return threads == 512 ? max_coresident_t<RS, 512>(smem, num_sms, out)
                        : max_coresident_t<RS, 256>(smem, num_sms, out);
}

// Picks the row-slot count RS, the grid and whether the matrix slice fits in shared memory.
static int plan_pcg(mvs_deform* h) {
  if (h->pcg_plan_valid) return MVS_OK;
  h->pcg_launch_per_phase = getenv("MVS_PCG_LAUNCHES") != nullptr;  // test switch
  if (h->pcg_launch_per_phase) {
    h->pcg_plan_valid = true;
    return MVS_OK;
  }
  mvs_ctx* ctx = h->ctx;
  int64_t m = h->m;
  // the smallest row-slot count whose grid is co-resident: a CTA walks its RS slots one after the
  // other in every phase, so 20,000 rows run 5 slots on 250 CTAs instead of 8 slots on 157
  const int rs_opts[7] = {1, 2, 3, 4, 5, 6, 8};
  int t_opts[2] = {kPcgDefaultThreads, kPcgDefaultThreads == 256 ? 512 : 256};
  int nt_opts = 2;
  if (const char* env = getenv("MVS_PCG_THREADS")) {
    int t = atoi(env);
    if (t == 256 || t == 512) {
      t_opts[0] = t;
      nt_opts = 1;
    }
  }
  for (int stage = 1; stage >= 0; --stage) {
    for (int ti = 0; ti < nt_opts; ++ti) {
      const int T = t_opts[ti];
      const int kGroups = T / kRowLanes;
      for (int oi = 0; oi < 7; ++oi) {
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
          case 1: MVS_TRY(max_coresident<1>(smem, ctx->num_sms, T, &cap)); break;
          case 2: MVS_TRY(max_coresident<2>(smem, ctx->num_sms, T, &cap)); break;
          case 3: MVS_TRY(max_coresident<3>(smem, ctx->num_sms, T, &cap)); break;
          case 4: MVS_TRY(max_coresident<4>(smem, ctx->num_sms, T, &cap)); break;
          case 5: MVS_TRY(max_coresident<5>(smem, ctx->num_sms, T, &cap)); break;
          case 6: MVS_TRY(max_coresident<6>(smem, ctx->num_sms, T, &cap)); break;
          default: MVS_TRY(max_coresident<8>(smem, ctx->num_sms, T, &cap)); break;
        }
        // the exchange words count arrivals in 10 bits (kCountShift): fewer than 1024 CTAs
        if (grid <= cap && grid < 1024) {
          h->pcg_rs = RS;
          h->pcg_threads = T;
          h->pcg_grid = grid;
          h->pcg_smem = smem;
          h->pcg_stage = stage;
          h->pcg_plan_valid = true;
          return MVS_OK;
        }
      }
    }
  }
  // not co-resident: the launch-per-phase solver (pcg_solve_launches)
  h->pcg_launch_per_phase = true;
  h->pcg_plan_valid = true;
  return MVS_OK;
}

// ------------------------------------------------------------------ launch-per-phase solver
// Graphs whose rows are not co-resident in one cooperative grid (more than ~37,000 nodes: the
// reference's own 16-NN sampling rule, Deformation.cpp:63-106, yields 60-100 K nodes on a 1M-vertex
// mesh) run the same recurrences as four launches per iteration — product + p.q, update + r.z,
// direction, one-thread scalar roll — with the scalars kept on the device (no host synchronisation)
// and the same fixed-order float64 reductions (reduce.cuh), so the solve stays bit-reproducible.  At
// that size every launch streams tens of MB, so launch latency no longer dominates.
constexpr int kLThreads = 256;
constexpr int kLGroups = kLThreads / kRowLanes;

struct PcgLArgs {
  const int* row_ptr;
  const int* col;
  const float* B;
  const float* Rot;
  const float* Minv;
  const float* rhs;
  float* x;
  float *r, *z, *p, *q, *d;  // m x 12 each
  double* scal;              // S_RZ, S_PQ, S_RZ_NEW, S_BNORM2, S_RNORM2
  double* partials;
  unsigned* ticket;
  int64_t m;
  float lambda;
  double tol;
};

// z_row = M^-1 r_row over the 16-lane group (lane k holds r[k])
__device__ __forceinline__ float precond_row(const float* __restrict__ Minv, int64_t j, int k,
                                             int gbase, float rk, bool act) {
  float zv = 0.f;
#pragma unroll
  for (int q = 0; q < 12; ++q) {
    float rq = __shfl_sync(0xffffffffu, rk, gbase + q);
    if (act) zv += __ldg(Minv + j * 144 + k * 12 + q) * rq;
  }
  return zv;
}

__global__ void __launch_bounds__(kLThreads) k_pcgl_init(PcgLArgs a) {
  __shared__ double smem[64];
  const int lane = threadIdx.x & 31, k = lane & (kRowLanes - 1), gbase = lane & ~(kRowLanes - 1);
  int64_t j = (int64_t)blockIdx.x * kLGroups + threadIdx.x / kRowLanes;
  bool act = j < a.m && k < 12;
  float rk = act ? -__ldg(a.rhs + j * 12 + k) : 0.f;
  float zv = precond_row(a.Minv, j, k, gbase, rk, act);
  if (act) {
    a.r[j * 12 + k] = rk;
    a.z[j * 12 + k] = zv;
    a.p[j * 12 + k] = zv;
    a.d[j * 12 + k] = 0.f;
  }
  double v[2] = {(double)rk * (double)zv, (double)rk * (double)rk};
  double out[2];
  if (grid_reduce_sum<2>(v, smem, a.partials, a.ticket, out) && threadIdx.x == 0) {
    a.scal[S_RZ] = out[0];
    a.scal[S_BNORM2] = out[1];
    a.scal[S_RNORM2] = out[1];
    a.scal[S_PQ] = 0.0;  // iterations run (as the persistent kernel reports them)
  }
}

// q = (H + lambda I) p, p.q
__global__ void __launch_bounds__(kLThreads) k_pcgl_product(PcgLArgs a) {
  __shared__ double smem[32];
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, k = lane & (kRowLanes - 1), gbase = lane & ~(kRowLanes - 1);
  int64_t j = (int64_t)blockIdx.x * kLGroups + threadIdx.x / kRowLanes;
  bool on = j < a.m, act = on && k < 12;
  const bool stop = a.tol > 0.0 && a.scal[S_RNORM2] <= a.tol * a.tol * a.scal[S_BNORM2];
  float pk = act ? a.p[j * 12 + k] : 0.f;
  float ya[12];
#pragma unroll
  for (int q = 0; q < 12; ++q) ya[q] = 0.f;
  if (on && !stop) {
    const int t0 = __ldg(a.row_ptr + j), t1 = __ldg(a.row_ptr + j + 1);
    for (int t = t0 + k; t < t1; t += kRowLanes) {
      const int l = __ldg(a.col + t);
      const float4* pl = reinterpret_cast<const float4*>(a.p + (size_t)l * 12);
      float4 u0 = __ldg(pl), u1 = __ldg(pl + 1), u2 = __ldg(pl + 2);
      float pv[12] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w, u2.x, u2.y, u2.z, u2.w};
      const float4* Bt = reinterpret_cast<const float4*>(a.B) + (size_t)t * 4;
#pragma unroll
      for (int aa = 0; aa < 4; ++aa) {
        float4 b = __ldg(Bt + aa);
#pragma unroll
        for (int r = 0; r < 3; ++r)
          ya[3 * aa + r] += b.x * pv[r] + b.y * pv[3 + r] + b.z * pv[6 + r] + b.w * pv[9 + r];
      }
    }
  }
  float y = group_reduce_scatter(ya, k);
  y += a.lambda * pk;
#pragma unroll
  for (int q = 0; q < 9; ++q) {
    float pq = __shfl_sync(full, pk, gbase + q);
    if (act && k < 9) y += __ldg(a.Rot + j * 81 + k * 9 + q) * pq;
  }
  if (!act || stop) y = 0.f;
  if (act) a.q[j * 12 + k] = y;
  double v[1] = {(double)pk * (double)y};
  double out[1];
  if (grid_reduce_sum<1>(v, smem, a.partials, a.ticket, out) && threadIdx.x == 0)
    a.scal[S_PQ_CUR] = out[0];  // p.q of this iteration
}

// alpha ; d += alpha p ; r -= alpha q ; z = M^-1 r ; r.z, |r|^2
__global__ void __launch_bounds__(kLThreads) k_pcgl_update(PcgLArgs a) {
  __shared__ double smem[64];
  const int lane = threadIdx.x & 31, k = lane & (kRowLanes - 1), gbase = lane & ~(kRowLanes - 1);
  int64_t j = (int64_t)blockIdx.x * kLGroups + threadIdx.x / kRowLanes;
  bool act = j < a.m && k < 12;
  const double rz = a.scal[S_RZ], pq = a.scal[S_PQ_CUR];
  const bool stop = a.tol > 0.0 && a.scal[S_RNORM2] <= a.tol * a.tol * a.scal[S_BNORM2];
  const float alpha = (!stop && pq > 0.0 && rz > 0.0) ? (float)(rz / pq) : 0.f;
  float rk = 0.f, zv = 0.f;
  if (act) {
    rk = a.r[j * 12 + k] - alpha * a.q[j * 12 + k];
    a.d[j * 12 + k] += alpha * a.p[j * 12 + k];
    a.r[j * 12 + k] = rk;
  }
  zv = precond_row(a.Minv, j, k, gbase, rk, act);
  if (act) a.z[j * 12 + k] = zv;
  double v[2] = {(double)rk * (double)zv, (double)rk * (double)rk};
  double out[2];
  if (grid_reduce_sum<2>(v, smem, a.partials, a.ticket, out) && threadIdx.x == 0) {
    a.scal[S_RZ_NEW] = out[0];
    if (!stop) {
      a.scal[S_RNORM2] = out[1];
      a.scal[S_PQ] += 1.0;
    }
  }
}

// beta = rz_new / rz ; p = z + beta p.  Every block reads both scalars; k_pcgl_roll (next launch)
// then moves rz_new into rz.
__global__ void __launch_bounds__(256) k_pcgl_direction(PcgLArgs a) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const double rz = a.scal[S_RZ], rzn = a.scal[S_RZ_NEW];
  const bool stop = a.tol > 0.0 && a.scal[S_RNORM2] <= a.tol * a.tol * a.scal[S_BNORM2];
  const float beta = (!stop && rz > 0.0) ? (float)(rzn / rz) : 0.f;
  if (t < a.m * 12 && !stop) a.p[t] = a.z[t] + beta * a.p[t];
}
__global__ void k_pcgl_roll(double* scal) { scal[S_RZ] = scal[S_RZ_NEW]; }

__global__ void __launch_bounds__(256) k_pcgl_finish(PcgLArgs a, float* __restrict__ delta) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= a.m * 12) return;
  float dv = a.d[t];
  a.x[t] += dv;
  delta[t] = dv;
}

static int pcg_solve_launches(mvs_deform* h) {
  mvs_ctx* ctx = h->ctx;
  cudaStream_t st = ctx->stream;
  const int64_t m = h->m;
  const size_t N = (size_t)m * 12;
  PcgLArgs a;
  a.row_ptr = h->row_ptr.p;
  a.col = h->col.p;
  a.B = h->B();
  a.Rot = h->Rot();
  a.Minv = h->Minv.p;
  a.rhs = h->rhs();
  a.x = h->x.p;
  a.r = h->pcg.p;
  a.z = h->pcg.p + N;
  a.p = h->pcg.p + 2 * N;
  a.q = h->pcg.p + 3 * N;
  a.d = h->pcg.p + 5 * N;
  float* delta = h->pcg.p + 4 * N;
  a.scal = h->scal.p;
  const int grid = (int)((m + kLGroups - 1) / kLGroups);
  MVS_TRY(h->partials.reserve((size_t)grid * 2 + 16));
  a.partials = h->partials.p;
  a.ticket = h->ticket.p;
  a.m = m;
  a.lambda = (float)h->prm.lambda;
  a.tol = h->prm.pcg_tol;
  ProfScope ps(ctx, "pcg_solve", 2 + 4 * std::max(h->prm.pcg_iters, 0));
  k_pcgl_init<<<grid, kLThreads, 0, st>>>(a);
  for (int it = 0; it < h->prm.pcg_iters; ++it) {
    k_pcgl_product<<<grid, kLThreads, 0, st>>>(a);
    k_pcgl_update<<<grid, kLThreads, 0, st>>>(a);
    k_pcgl_direction<<<div_up((int64_t)N, 256), 256, 0, st>>>(a);
    k_pcgl_roll<<<1, 1, 0, st>>>(a.scal);
  }
  k_pcgl_finish<<<div_up((int64_t)N, 256), 256, 0, st>>>(a, delta);
  MVS_CUDA(cudaGetLastError());
  return MVS_OK;
}

int pcg_solve(mvs_deform* h) {
  mvs_ctx* ctx = h->ctx;
  MVS_TRY(plan_pcg(h));
  if (h->pcg_launch_per_phase) return pcg_solve_launches(h);
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
  // exchange slots of the fused barrier + reduction: "empty" = all-ones
  // two halves [ring | words]: a launch uses one and re-arms the other (the previous launch's)
  size_t ring_doubles = (size_t)kRing * h->pcg_grid * 2;
  size_t nwords = (size_t)(2 * std::max(h->prm.pcg_iters, 0) + 1) * 4;
  size_t half = ring_doubles + nwords;
  if (h->pcg_xchg_ring != ring_doubles || h->pcg_xchg_words != nwords) {
    MVS_TRY(h->pcg_xchg.reserve(2 * half));
    for (int k = 0; k < 2; ++k) {
      MVS_CUDA(cudaMemsetAsync(h->pcg_xchg.p + k * half, 0xff, ring_doubles * sizeof(double), ctx->stream));
      MVS_CUDA(cudaMemsetAsync(h->pcg_xchg.p + k * half + ring_doubles, 0, nwords * sizeof(double),
                               ctx->stream));
    }
    h->pcg_xchg_ring = ring_doubles;
    h->pcg_xchg_words = nwords;
    h->pcg_epoch = 0;
  }
  double* cur = h->pcg_xchg.p + (h->pcg_epoch & 1) * half;
  double* nxt = h->pcg_xchg.p + ((h->pcg_epoch + 1) & 1) * half;
  ++h->pcg_epoch;
  a.ring = cur;
  a.words = reinterpret_cast<unsigned long long*>(cur + ring_doubles);
  a.next_ring = reinterpret_cast<unsigned long long*>(nxt);
  a.next_words = reinterpret_cast<unsigned long long*>(nxt + ring_doubles);
  a.ring_count = (int)ring_doubles;
  a.word_count = (int)nwords;
  a.scal = h->scal.p;
  a.m = m;
  a.iters = h->prm.pcg_iters;
  a.lambda = (float)h->prm.lambda;
  a.tol = h->prm.pcg_tol;
  a.rows_per_cta = h->pcg_rs * (h->pcg_threads / kRowLanes);
  a.stage_smem = h->pcg_stage;
  a.diverged = ctx->solver_flag;
  ProfScope ps(ctx, "pcg_solve");
  switch (h->pcg_rs) {
    case 1: MVS_TRY(launch_pcg<1>(h, a, h->pcg_grid, h->pcg_smem)); break;
    case 2: MVS_TRY(launch_pcg<2>(h, a, h->pcg_grid, h->pcg_smem)); break;
    case 3: MVS_TRY(launch_pcg<3>(h, a, h->pcg_grid, h->pcg_smem)); break;
    case 4: MVS_TRY(launch_pcg<4>(h, a, h->pcg_grid, h->pcg_smem)); break;
    case 5: MVS_TRY(launch_pcg<5>(h, a, h->pcg_grid, h->pcg_smem)); break;
    case 6: MVS_TRY(launch_pcg<6>(h, a, h->pcg_grid, h->pcg_smem)); break;
    default: MVS_TRY(launch_pcg<8>(h, a, h->pcg_grid, h->pcg_smem)); break;
  }
  MVS_CUDA(cudaGetLastError());
  return MVS_OK;
}

}  // namespace mvs
