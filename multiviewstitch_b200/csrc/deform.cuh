// deform.cuh — device-resident state of one Deformation object (Deformation/Deformation.h:224-252)
// and the internal interfaces between deform_graph.cu (graph construction, correspondences),
// gn_assemble.cu (per-vertex passes + normal-equation assembly), pcg.cu (solver) and
// deform_api.cu (C ABI).
//
// Data layout in HBM (all float32, 16-byte records, SoA across arrays):
//   v[n], nrm[n]          float4  rest vertices (.w = tuple-run id) / normals in SORTED order: vertices are permuted
//                                 once at graph-build time so that vertices sharing the same
//                                 4-node set are adjacent (key = primary node | tuple hash);
//                                 perm[i] = original vertex id of sorted position i
//   vidx[n] int4, vw[n] float4    the k=4 nearest nodes of each vertex (ascending node id) and
//                                 their normalised weights w = (1 - d/d_max)^2
//   resid[n] float4               (deformed - target, omega) written by the correspondence pass
//   g[m], gnrm[m] float4          node positions / normals;  x[m][12] node state (a1,a2,a3,t)
//   normal equations, Kronecker-compressed:  one contiguous float buffer
//       sys = [ B: nnz x 16 | rhs: m x 12 | Rot: m x 81 ]
//   block-CSR pattern row_ptr[m+1], col[nnz] (ascending, diagonal included)
#pragma once
#include <algorithm>
#include <utility>
#include <vector>

#include "common.cuh"
#include "nn_grid.cuh"

namespace mvs {
struct ArapState;  // arap.cu
void arap_free(ArapState* a);
struct PeerLink;   // comm.cuh
}  // namespace mvs

struct mvs_deform {
  mvs_ctx* ctx = nullptr;
  int64_t n = 0, nf = 0, m = 0;
  // original-order inputs
  mvs::DevBuf<float4> v0, nrm0;  // rest vertices / normals, original order
  mvs::DevBuf<int> facets;       // 3*nf, original vertex ids
  // nodes
  std::vector<int32_t> node_ids;
  mvs::DevBuf<int> node_ids_d;
  mvs::DevBuf<float4> g, gnrm;
  mvs_nn_index* node_index = nullptr;
  int node_k = 4, graph_k = 8;
  bool graph_built = false;
  // sorted per-vertex data
  mvs::DevBuf<int> perm;      // sorted position -> original id
  mvs::DevBuf<int> inv_perm;  // original id -> sorted position
  mvs::DevBuf<float4> v, nrm;
  mvs::DevBuf<int4> vidx;
  mvs::DevBuf<float4> vw;
  mvs::DevBuf<int4> vidx_knn;  // kNN order (nearest first), original vertex order — API view
  mvs::DevBuf<float4> vw_knn;
  mvs::DevBuf<int> graph;  // m x (graph_k+1)
  // pattern + system
  mvs::DevBuf<int> row_ptr, col;
  int64_t nnz = 0;
  std::vector<int> h_row_ptr;  // host copy of row_ptr (solver launch planning)
  std::vector<int> h_grp_start;  // first sorted vertex of every assembly group, then n
  bool pcg_plan_valid = false;
  bool pcg_launch_per_phase = false;  // graph too large for the persistent cooperative solver
  int pcg_rs = 1, pcg_grid = 0, pcg_stage = 0, pcg_threads = 256;
  size_t pcg_smem = 0;
  mvs::DevBuf<float> sys;
  // order-independent accumulation: J^T J / J^T r are summed as 64-bit fixed point (integer atomics
  // commute exactly), then converted to the float32 system above; only the upper triangle of the
  // block pattern is stored: layout [U: nup x 16 | rhs: m x 12]
  mvs::DevBuf<long long> sys64;
  // vertex-sharded solves: sys64 and sys are views into ONE cudaMalloc'd exchange block that the other
  // ranks map through CUDA IPC (comm.cuh); null on single-rank contexts
  mvs::PeerLink* peer = nullptr;
  double fx_scale = 1.0;   // power of two, set per solve from the magnitude bound
  double diag2 = 1.0;      // squared bbox diagonal of the rest mesh
  float omega_max = 1.0f;  // largest per-vertex constraint weight
  // runs of equal node 4-tuples in the sorted vertex order: v[i].w = run id; per run the 16 pattern
  // slots (j_a, j_b) and the 4 node ids, so the assembly never searches the block pattern
  mvs::DevBuf<int> run_slots;  // nruns x 20
  mvs::DevBuf<int> run_start;  // nruns + 1: first sorted vertex of each run
  int64_t nruns = 0;
  // assembly groups: kGrpRuns consecutive runs share one shared-memory accumulator set.  Only the
  // upper triangle of J^T J is accumulated (block (j,l), j <= l); tslot mirrors it on conversion.
  mvs::DevBuf<float4> run_g;            // nruns x 4: positions of the run's four nodes (.w = valid)
  mvs::DevBuf<unsigned char> run_local; // nruns x 16: group-local accumulator of each of the 14 tiles
  mvs::DevBuf<int> grp_n;               // ngroups: distinct accumulators of the group
  mvs::DevBuf<int> grp_dst;             // ngroups x kGrpMax: first sys64 element of each accumulator
  mvs::DevBuf<int> tslot;               // nnz: pattern slot of the transposed block
  mvs::DevBuf<int> uidx;                // nnz: index of the block (or of its transpose) in sys64
  int64_t ngroups = 0;
  int64_t nup = 0;                      // upper-triangle blocks, diagonal included
  mvs::DevBuf<float> Minv;  // m x 144
  mvs::DevBuf<float> x;     // m x 12
  mvs::DevBuf<float> pcg;   // r, z, p0, p1, q, delta : 6 x m x 12
  mvs::DevBuf<double> scal; // PCG scalars + energies
  mvs::DevBuf<double> partials;
  mvs::DevBuf<unsigned> ticket;
  // target scan
  mvs_nn_index* tgt_index = nullptr;
  mvs::DevBuf<float4> tp_own, tn_own;
  const float4* tp = nullptr;
  const float4* tn = nullptr;
  int64_t nt = 0;
  // per-vertex work arrays (sorted order)
  mvs::DevBuf<float4> resid;    // (v~ - c, omega)
  mvs::DevBuf<int> corr_prev;   // last correspondence of each vertex (seed of the next search)
  bool corr_seeded = false;     // corr_prev holds valid seeds for the current target
  mvs::DevBuf<double> pcg_xchg;  // persistent PCG: two halves of [exchange ring | exchange words]
  size_t pcg_xchg_ring = 0, pcg_xchg_words = 0;
  unsigned pcg_epoch = 0;
  bool node_terms_forked = false;  // this iteration's node terms are in flight on the side stream
  mvs::DevBuf<float4> targets;  // (c, omega) for fixed-target modes
  bool have_targets = false;
  mvs::DevBuf<float4> cur0;     // deformed vertices, ORIGINAL order
  mvs::DevBuf<float4> curn0;    // deformed normals, original order
  bool cur_valid = false;
  // solver run state
  mvs_gn_params prm{};
  bool gn_active = false;
  int64_t v_begin = 0, v_end = -1;  // shard of sorted positions owned by this rank
  // the rank's range is worked through as `shard_parts` consecutive sub-ranges, one launch each — the
  // partial sums `shard_parts` ranks would produce, added by the same integer atomics the all-reduce
  // performs (mvs_deform_set_shard_parts: single-GPU check of the sharded arithmetic)
  int shard_parts = 1;
  // interleaved sharding (mvs_deform_set_shard_interleaved): [0, n) is cut into il_nranks x il_chunks
  // chunks and this rank takes chunks il_rank, il_rank + il_nranks, ... — the cost of the
  // correspondence pass varies over the mesh (far regions search longer), contiguous ranges do not
  // balance it.  il_nranks = 0: the plain range [v_begin, v_end)
  int il_rank = 0, il_nranks = 0, il_chunks = 1;
  double energy[3] = {0, 0, 0};
  // as-rigid-as-possible solver state (arap.cu), built on first use
  mvs::ArapState* arap = nullptr;

  float* B() { return sys.p; }
  // rhs before Rot: m*81 floats would break the 16-byte alignment the vector reductions need
  float* rhs() { return sys.p + (size_t)nnz * 16; }
  float* Rot() { return sys.p + (size_t)nnz * 16 + (size_t)m * 12; }
  size_t sys_floats() const { return (size_t)nnz * 16 + (size_t)m * 81 + (size_t)m * 12; }
  size_t sys64_count() const { return (size_t)nup * 16 + (size_t)m * 12; }
  long long* B64() { return sys64.p; }
  long long* rhs64() { return sys64.p + (size_t)nup * 16; }
  // A requested boundary is moved DOWN to the start of the assembly group that contains it (0 and n
  // stay): ranks that split [0, n) at common points still tile it exactly, and no group — the unit
  // in which float32 partial sums are formed before they become integers — is split between ranks.
  int64_t snap(int64_t i) const {
    if (i <= 0 || i >= n || h_grp_start.empty()) return i < 0 ? 0 : (i > n ? n : i);
    auto it = std::upper_bound(h_grp_start.begin(), h_grp_start.end(), (int)i);
    return (int64_t) * (it - 1);
  }
  // index of the assembly group that starts at sorted vertex i (i is 0, n or a snapped boundary)
  int group_at(int64_t i) const {
    if (i <= 0) return 0;
    auto it = std::lower_bound(h_grp_start.begin(), h_grp_start.end(), (int)i);
    return (int)(it - h_grp_start.begin());
  }
  int64_t shard_begin() const { return snap(v_begin); }
  int64_t shard_end() const { return v_end < 0 ? n : snap(v_end); }
  // the sorted-vertex sub-ranges this rank works through, each snapped to assembly groups
  std::vector<std::pair<int64_t, int64_t>> shard_ranges() const {
    std::vector<std::pair<int64_t, int64_t>> out;
    if (il_nranks > 0) {
      const int64_t T = (int64_t)il_nranks * il_chunks;
      for (int64_t k = il_rank; k < T; k += il_nranks) {
        int64_t b = k == 0 ? 0 : snap(n * k / T), e = k + 1 == T ? n : snap(n * (k + 1) / T);
        if (e > b) out.emplace_back(b, e);
      }
      return out;
    }
    const int64_t b0 = shard_begin(), e0 = shard_end();
    const int parts = shard_parts < 1 ? 1 : shard_parts;
    const int64_t per = (e0 - b0 + parts - 1) / parts;
    for (int part = 0; part < parts; ++part) {
      int64_t lo = std::min(e0, b0 + per * part), hi = std::min(e0, b0 + per * (part + 1));
      int64_t b = part == 0 ? b0 : std::max(b0, snap(lo));
      int64_t e = part == parts - 1 ? e0 : std::max(b0, snap(hi));
      if (e > b) out.emplace_back(b, e);
    }
    return out;
  }
};

namespace mvs {

#ifndef MVS_GRP_RUNS
#define MVS_GRP_RUNS 4
#endif
constexpr int kGrpRuns = MVS_GRP_RUNS;  // runs per assembly group
constexpr int kGrpMax = 14 * kGrpRuns;  // 10 upper-triangle block tiles + 4 gradient tiles per run

// the sub-ranges of one launch of the per-vertex passes (a rank's interleaved chunks): up to 16, mapped
// from the flat thread / ticket index inside the kernel so that the chunks fill waves together
constexpr int kMaxShardRanges = 16;
struct ShardRanges {
  int n;
  int64_t b[kMaxShardRanges], e[kMaxShardRanges];
  int64_t voff[kMaxShardRanges + 1];  // prefix sums of the vertex counts
  int g_lo[kMaxShardRanges];          // first assembly group of each range
  int goff[kMaxShardRanges + 1];      // prefix sums of the group counts
};

// scal[] slots
enum { S_RZ = 0, S_PQ = 1, S_RZ_NEW = 2, S_BNORM2 = 3, S_RNORM2 = 4, S_PQ_CUR = 5, S_EFIT = 8, S_EROT = 9,
       S_EREG = 10, S_SIZE = 16 };

// deform_graph.cu
int deform_build_graph(mvs_deform* h, int node_k, int graph_k);
int deform_node_targets(mvs_deform* h, double proj_len_err, double proj_dist_err,
                        int smooth_sweeps, DevBuf<double>& controls, DevBuf<int>& valid);

// gn_assemble.cu
int gn_reset_state(mvs_deform* h);
int gn_deform_and_correspond(mvs_deform* h);   // writes resid (mode 1: NN; else fixed targets)
int gn_node_terms_forked(mvs_deform* h);        // zero + node terms on the side stream (one rank)
int gn_assemble(mvs_deform* h);                 // [zero +] fit pass [+ allreduce + node terms] + precond
int gn_final_vertices(mvs_deform* h);           // cur0 / curn0 in original order
int gn_energy(mvs_deform* h, double energy[3]); // at the current state with the current resid
int gn_recompute_normals(mvs_deform* h);        // curn0 from cur0 + facets (Deformation.h:86-128)

// arap.cu
int arap_solve(mvs_deform* h, const double* ctrl_xyz_host, const int32_t* ctrl_ids_host, int64_t m,
               const mvs_arap_params* prm, double* energy_out, int* iters_out);

// pcg.cu
int pcg_solve(mvs_deform* h);
// device-resident building blocks of mvs_stitch_pairs (pipeline.cu, pointset.cu, deform_api.cu)
int depth2model_dev(mvs_ctx* ctx, const float* inv_depth_host, const mvs_camera* cam, double min_dsp,
                    double max_dsp, double smooth_threshold, DevBuf<double>* xyz, DevBuf<int>* tex,
                    DevBuf<int>* facets, int64_t* n_points, int64_t* n_facets);
int apply_srt_dev_f64(mvs_ctx* ctx, double* xyz_dev, int64_t n, double scale, const double R[9],
                      const double t[3]);
int deform_create_dev_f64(mvs_ctx* ctx, const double* xyz_dev, int64_t n, const int* facets_dev,
                          int64_t nf, mvs_deform** out);
int deform_vertices_dev(mvs_deform* h, const float4** cur);
// model2depth.cu: inverse-depth image (host) -> valid world points compacted in pixel order (device)
int unproject_compact_f4(mvs_ctx* ctx, const float* inv_depth, const mvs_camera* cam, double min_dsp,
                         double max_dsp, DevBuf<float4>& out, DevBuf<int>& scan_tmp, int64_t* count);

}  // namespace mvs
