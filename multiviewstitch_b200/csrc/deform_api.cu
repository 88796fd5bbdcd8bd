// deform_api.cu — C ABI of the Deformation object (Deformation/Deformation.h:224-252): lifetime,
// target upload, the Deform() drivers and the getters.  All arithmetic lives in deform_graph.cu,
// gn_assemble.cu and pcg.cu.
#include <stdio.h>

#include <algorithm>
#include <cmath>

#include "comm.cuh"
#include "deform.cuh"
#include "nn_pyramid.cuh"

namespace mvs {

__global__ void k_f4_to_f64xyz(const float4* __restrict__ in, double* __restrict__ out, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = in[i];
  out[3 * i] = p.x;
  out[3 * i + 1] = p.y;
  out[3 * i + 2] = p.z;
}
__global__ void k_f4_to_f32xyz(const float4* __restrict__ in, float* __restrict__ out, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = in[i];
  out[3 * i] = p.x;
  out[3 * i + 1] = p.y;
  out[3 * i + 2] = p.z;
}

// targets (original vertex order: xyz doubles + float weight) -> sorted float4 (c, omega)
__global__ void k_targets_permute(const double* __restrict__ t, const float* __restrict__ w,
                                  const int* __restrict__ perm, int64_t n,
                                  float4* __restrict__ out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int o = perm[i];
  float om = w ? w[o] : 1.0f;
  out[i] = make_float4((float)t[3 * (size_t)o], (float)t[3 * (size_t)o + 1],
                       (float)t[3 * (size_t)o + 2], om);
}

// out[perm[i]] = in[i]: sorted vertex order -> original vertex order
__global__ void k_unpermute_i32(const int* __restrict__ in, const int* __restrict__ perm, int64_t n,
                                int* __restrict__ out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[perm[i]] = in[i];
}

// node controls (m x 3 doubles) -> per-vertex targets on the node vertices only
__global__ void k_targets_from_nodes(const double* __restrict__ controls,
                                     const int* __restrict__ valid, int only_valid,
                                     const int* __restrict__ node_ids,
                                     const int* __restrict__ inv_perm, int64_t m,
                                     float4* __restrict__ targets) {
  int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= m) return;
  if (only_valid && !valid[j]) return;
  int pos = inv_perm[node_ids[j]];
  targets[pos] = make_float4((float)controls[3 * j], (float)controls[3 * j + 1],
                             (float)controls[3 * j + 2], 1.0f);
}

// Deformation/Deformation.cpp:186-203: 1-NN target of each node; keep if d2 <= threshold and
// cos(dir, normal) >= 0.5 (the 6-D tree has scale = 0, i.e. it is a plain position search).
__global__ void k_node_targets_threshold(PyrParams tg, const float4* __restrict__ g,
                                         const float4* __restrict__ gnrm, int64_t m,
                                         double threshold, double* __restrict__ controls,
                                         int* __restrict__ valid) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= m) return;
  float4 q = g[i];
  float4 nq = gnrm[i];
  valid[i] = 0;
  controls[3 * i] = q.x;
  controls[3 * i + 1] = q.y;
  controls[3 * i + 2] = q.z;
  KnnList<1> nn;
  nn.init(q.x, q.y, q.z);
  struct V {
    KnnList<1>* l;
    float cx, cy, cz;
    __device__ float worst2() const { return l->worst2(); }
    __device__ void point(float4 p) {
      float before = l->d2[0];
      int bi = l->idx[0];
      l->point(p);
      if (l->d2[0] != before || l->idx[0] != bi) {
        cx = p.x; cy = p.y; cz = p.z;
      }
    }
  } vis{&nn, 0.f, 0.f, 0.f};
  pyr_search(tg, q.x, q.y, q.z, vis);
  if ((double)nn.d2[0] > threshold) return;  // float dists[0] against the double threshold (:198)
  double d0 = __dsub_rn((double)vis.cx, (double)q.x), d1 = __dsub_rn((double)vis.cy, (double)q.y),
         d2 = __dsub_rn((double)vis.cz, (double)q.z);
  double n0 = nq.x, n1 = nq.y, n2 = nq.z;
  double dot = __dadd_rn(__dadd_rn(__dmul_rn(d0, n0), __dmul_rn(d1, n1)), __dmul_rn(d2, n2));
  double dl = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)), __dmul_rn(d2, d2)));
  double nl = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(n0, n0), __dmul_rn(n1, n1)), __dmul_rn(n2, n2)));
  double c = __ddiv_rn(dot, __dmul_rn(dl, nl));
  if (c < 0.5) return;  // NaN (dl == 0) compares false, as in the reference
  valid[i] = 1;
  controls[3 * i] = vis.cx;
  controls[3 * i + 1] = vis.cy;
  controls[3 * i + 2] = vis.cz;
}

static int ensure_graph(mvs_deform* h) {
  if (h->m <= 0) {
    // Deformation.cpp:248-250: sample when no nodes exist yet
    MVS_TRY(mvs_deform_uniform_sampling(h, 16));
  }
  if (!h->graph_built) MVS_TRY(deform_build_graph(h, h->node_k, h->graph_k));
  return MVS_OK;
}

static int validate_params(const mvs_gn_params* p) {
  if (!p) {
    set_error("null mvs_gn_params");
    return MVS_ERR_INVALID_ARG;
  }
  if (p->gn_iters < 0 || p->pcg_iters < 0 || p->corr_mode < 0 || p->corr_mode > 2 ||
      !(p->w_fit >= 0) || !(p->w_rot >= 0) || !(p->w_reg >= 0) || !(p->lambda >= 0)) {
    set_error("invalid mvs_gn_params");
    return MVS_ERR_INVALID_ARG;
  }
  return MVS_OK;
}

static int gn_begin_impl(mvs_deform* h, const mvs_gn_params* p) {
  MVS_TRY(validate_params(p));
  MVS_TRY(ensure_graph(h));
  h->prm = *p;
  MVS_TRY(gn_reset_state(h));
  // a solve never inherits correspondence seeds from an earlier one
  MVS_CUDA(cudaMemsetAsync(h->corr_prev.p, 0xff, sizeof(int) * h->n, h->ctx->stream));
  h->corr_seeded = false;
  h->gn_active = true;
  h->cur_valid = false;
  return MVS_OK;
}

// after a stream synchronise: did a solver kernel raise the context's divergence flag?
static int check_solver_flag(mvs_deform* h) {
  int* f = h->ctx->solver_flag;
  if (f && *reinterpret_cast<volatile int*>(f)) {
    const int why = *f;
    float dv = 0.f, ds = 0.f;
    memcpy(&dv, f + 1, 4);
    memcpy(&ds, f + 2, 4);
    f[0] = f[1] = f[2] = 0;
    if (why == 1)
      set_error("Gauss-Newton solve diverged: a PCG partial sum is not finite (check the inputs and weights)");
    else if (why == 2)
      set_error("Gauss-Newton solve diverged: a PCG partial sum left the range of the exchange words");
    else
      set_error("Gauss-Newton solve diverged: a normal-equation entry (%g at fixed-point scale %g) is not "
                "finite or leaves the fixed-point range (targets far outside the mesh's reach?)", dv, ds);
    return MVS_ERR_DEGENERATE;
  }
  return MVS_OK;
}

static int gn_step_impl(mvs_deform* h) {
  MVS_TRY(gn_node_terms_forked(h));  // side stream, overlaps the correspondence pass
  int rc = gn_deform_and_correspond(h);
  if (rc != MVS_OK) {  // keep the streams ordered before reporting the error
    if (h->node_terms_forked) ctx_side_join(h->ctx);
    h->node_terms_forked = false;
    return rc;
  }
  MVS_TRY(gn_assemble(h));
  MVS_TRY(pcg_solve(h));  // also applies x += delta
  return MVS_OK;
}

// overwrite_initial_geometry() (Deformation.cpp:400): the deformed mesh becomes the rest mesh
static int bake(mvs_deform* h) {
  mvs_ctx* ctx = h->ctx;
  int64_t n = h->n;
  MVS_CUDA(cudaMemcpyAsync(h->v0.p, h->cur0.p, sizeof(float4) * n, cudaMemcpyDeviceToDevice,
                           ctx->stream));
  if (h->nf > 0)
    MVS_CUDA(cudaMemcpyAsync(h->nrm0.p, h->curn0.p, sizeof(float4) * n, cudaMemcpyDeviceToDevice,
                             ctx->stream));
  h->graph_built = false;  // rebuilt lazily from the new rest geometry
  h->have_targets = false;
  if (h->arap) {  // cotangent weights belong to the old rest geometry
    arap_free(h->arap);
    h->arap = nullptr;
  }
  if (h->m > 0) {
    std::vector<int32_t> ids = h->node_ids;
    MVS_TRY(mvs_deform_set_nodes(h, ids.data(), (int64_t)ids.size()));
  }
  return MVS_OK;
}

static int set_node_targets(mvs_deform* h, DevBuf<double>& controls, DevBuf<int>& valid,
                            int only_valid) {
  mvs_ctx* ctx = h->ctx;
  MVS_CUDA(cudaMemsetAsync(h->targets.p, 0, sizeof(float4) * h->n, ctx->stream));
  {
    ProfScope ps(ctx, "targets_from_nodes");
    k_targets_from_nodes<<<div_up(h->m, 128), 128, 0, ctx->stream>>>(
        controls.p, valid.p, only_valid, h->node_ids_d.p, h->inv_perm.p, h->m, h->targets.p);
  }
  MVS_CUDA(cudaGetLastError());
  h->have_targets = true;
  h->omega_max = 1.0f;
  return MVS_OK;
}

}  // namespace mvs

using namespace mvs;

extern "C" {

int mvs_gn_default_params(mvs_gn_params* p) {
  if (!p) return MVS_ERR_INVALID_ARG;
  p->gn_iters = 10;
  p->pcg_iters = 20;
  p->corr_mode = 0;
  p->outer_rounds = 1;
  p->w_rot = 1.0;
  p->w_reg = 10.0;
  p->w_fit = 100.0;
  p->lambda = 1e-6;
  p->pcg_tol = 0.0;
  p->max_corr_dist2 = 0.0;
  p->min_corr_cos = -2.0;
  return MVS_OK;
}

static int create_common(mvs_ctx* ctx, int64_t n, const int32_t* facets, int64_t nf,
                         mvs_deform** out, mvs_deform** hp) {
  if (!out) return MVS_ERR_INVALID_ARG;
  *out = nullptr;
  if (n <= 0) {
    set_error("Deformation: empty mesh");
    return MVS_ERR_DEGENERATE;
  }
  if (nf < 0 || (nf > 0 && !facets)) {
    set_error("Deformation: bad facet array");
    return MVS_ERR_INVALID_ARG;
  }
  for (int64_t i = 0; i < 3 * nf; ++i)
    if (facets[i] < 0 || facets[i] >= n) {
      // the reference aborts on an invalid polyhedron (Deformation.cpp:42-45); we return a code
      set_error("Deformation: facet index %d out of range", facets[i]);
      return MVS_ERR_INVALID_ARG;
    }
  if (!ctx) ctx = default_ctx();
  if (!ctx) return MVS_ERR_NO_DEVICE;
  MVS_TRY(ctx_ensure_device(ctx));
  mvs_deform* h = new mvs_deform();
  h->ctx = ctx;
  ctx_retain(ctx);
  h->n = n;
  h->nf = nf;
  mvs_gn_default_params(&h->prm);
  *hp = h;
  return MVS_OK;
}

static int finish_create(mvs_deform* h, const int32_t* facets, mvs_deform** out) {
  mvs_ctx* ctx = h->ctx;
  int st = h->facets.reserve((size_t)3 * h->nf + 1);
  if (st == MVS_OK && h->nf > 0) st = h2d(ctx, h->facets.p, facets, sizeof(int) * 3 * h->nf);
  if (st == MVS_OK) {
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
      set_error("create: %s", cudaGetErrorString(e));
      st = MVS_ERR_CUDA;
    }
  }
  if (st != MVS_OK) {
    mvs_deform_destroy(h);
    return st;
  }
  *out = h;
  return MVS_OK;
}

int mvs_deform_create(mvs_ctx* ctx, const double* pts, const double* normals, int64_t n,
                      const int32_t* facets, int64_t nf, mvs_deform** out) {
  MVS_REQUIRE(pts, "null points");
  mvs_deform* h = nullptr;
  MVS_TRY(create_common(ctx, n, facets, nf, out, &h));
  int st = upload_xyz_f64(h->ctx, pts, n, h->v0);
  if (st == MVS_OK) {
    if (normals)
      st = upload_xyz_f64(h->ctx, normals, n, h->nrm0);
    else {
      st = h->nrm0.reserve((size_t)n);
      if (st == MVS_OK) cudaMemsetAsync(h->nrm0.p, 0, sizeof(float4) * n, h->ctx->stream);
    }
  }
  if (st != MVS_OK) {
    mvs_deform_destroy(h);
    return st;
  }
  return finish_create(h, facets, out);
}

int mvs_deform_create_f32(mvs_ctx* ctx, const float* pts, const float* normals, int64_t n,
                          const int32_t* facets, int64_t nf, mvs_deform** out) {
  MVS_REQUIRE(pts, "null points");
  mvs_deform* h = nullptr;
  MVS_TRY(create_common(ctx, n, facets, nf, out, &h));
  int st = upload_xyz_f32(h->ctx, pts, n, h->v0);
  if (st == MVS_OK) {
    if (normals)
      st = upload_xyz_f32(h->ctx, normals, n, h->nrm0);
    else {
      st = h->nrm0.reserve((size_t)n);
      if (st == MVS_OK) cudaMemsetAsync(h->nrm0.p, 0, sizeof(float4) * n, h->ctx->stream);
    }
  }
  if (st != MVS_OK) {
    mvs_deform_destroy(h);
    return st;
  }
  return finish_create(h, facets, out);
}

}  // extern "C"

namespace mvs {
__global__ void k_xyz_f64_to_f4_dev(const double* __restrict__ in, float4* __restrict__ out, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = make_float4((float)in[3 * i], (float)in[3 * i + 1], (float)in[3 * i + 2], 0.f);
}

// Deformation(points, normals = none, facets) from DEVICE-resident doubles and facets: the same
// float32 rounding as mvs_deform_create (upload, then cast on the device), no host copies.  The
// facets come from depth2model_dev and index [0, n) by construction.
int deform_create_dev_f64(mvs_ctx* ctx, const double* xyz_dev, int64_t n, const int* facets_dev,
                          int64_t nf, mvs_deform** out) {
  mvs_deform* h = nullptr;
  int32_t dummy = 0;
  MVS_TRY(create_common(ctx, n, nf > 0 ? &dummy : nullptr, 0, out, &h));
  h->nf = nf;
  int st = h->v0.reserve((size_t)n);
  if (st == MVS_OK) st = h->nrm0.reserve((size_t)n);
  if (st == MVS_OK) st = h->facets.reserve((size_t)3 * nf + 1);
  if (st == MVS_OK) {
    cudaStream_t s = h->ctx->stream;
    ProfScope ps(h->ctx, "convert_xyz");
    k_xyz_f64_to_f4_dev<<<div_up(n, 256), 256, 0, s>>>(xyz_dev, h->v0.p, n);
    cudaMemsetAsync(h->nrm0.p, 0, sizeof(float4) * n, s);
    if (nf > 0)
      cudaMemcpyAsync(h->facets.p, facets_dev, sizeof(int) * 3 * nf, cudaMemcpyDeviceToDevice, s);
    if (cudaGetLastError() != cudaSuccess) st = MVS_ERR_CUDA;
  }
  if (st != MVS_OK) {
    mvs_deform_destroy(h);
    return st;
  }
  *out = h;
  return MVS_OK;
}

// the deformed vertices (float4, original order) where they lie on the device
int deform_vertices_dev(mvs_deform* h, const float4** cur) {
  *cur = h->cur_valid ? h->cur0.p : h->v0.p;
  return MVS_OK;
}
}  // namespace mvs

extern "C" {

int mvs_deform_destroy(mvs_deform* h) {
  if (!h) return MVS_OK;
  cudaSetDevice(h->ctx->device);
  cudaStreamSynchronize(h->ctx->stream);
  if (h->node_index) mvs_nn_index_destroy(h->node_index);
  if (h->tgt_index) mvs_nn_index_destroy(h->tgt_index);
  if (h->arap) arap_free(h->arap);
  if (h->peer) {
    h->sys.release();
    h->sys64.release();
    peer_link_free(h->peer);
    h->peer = nullptr;
  }
  mvs_ctx* ctx = h->ctx;
  delete h;
  ctx_release(ctx);
  return MVS_OK;
}

static int set_target_index(mvs_deform* h, int64_t nt) {
  MVS_TRY(nn_build_pyramid(h->tgt_index, h->tp, nt, kPyrPpcFar));
  // correspondences of a previous target are no seeds for this one
  if (h->corr_prev.p && h->n > 0)
    MVS_CUDA(cudaMemsetAsync(h->corr_prev.p, 0xff, sizeof(int) * h->n, h->ctx->stream));
  h->corr_seeded = false;
  return MVS_OK;
}

static int target_common(mvs_deform* h, int64_t nt) {
  MVS_TRY(ctx_ensure_device(h->ctx));
  if (nt <= 0) {
    set_error("set_target: empty target");
    return MVS_ERR_DEGENERATE;
  }
  if (!h->tgt_index) MVS_TRY(mvs_nn_index_create(h->ctx, &h->tgt_index));
  h->nt = nt;
  return MVS_OK;
}

int mvs_deform_set_target(mvs_deform* h, const double* tpts, const double* tnormals, int64_t nt) {
  MVS_REQUIRE(h && tpts, "null argument");
  MVS_TRY(target_common(h, nt));
  MVS_TRY(upload_xyz_f64(h->ctx, tpts, nt, h->tp_own));
  h->tp = h->tp_own.p;
  h->tn = nullptr;
  if (tnormals) {
    MVS_TRY(upload_xyz_f64(h->ctx, tnormals, nt, h->tn_own));
    h->tn = h->tn_own.p;
  }
  return set_target_index(h, nt);
}

int mvs_deform_set_target_f32(mvs_deform* h, const float* tpts, const float* tnormals,
                              int64_t nt) {
  MVS_REQUIRE(h && tpts, "null argument");
  MVS_TRY(target_common(h, nt));
  MVS_TRY(upload_xyz_f32(h->ctx, tpts, nt, h->tp_own));
  h->tp = h->tp_own.p;
  h->tn = nullptr;
  if (tnormals) {
    MVS_TRY(upload_xyz_f32(h->ctx, tnormals, nt, h->tn_own));
    h->tn = h->tn_own.p;
  }
  return set_target_index(h, nt);
}

int mvs_deform_set_target_depth(mvs_deform* h, const float* inv_depth, const mvs_camera* cam,
                                double min_dsp, double max_dsp, int64_t* n_target) {
  MVS_REQUIRE(h && inv_depth && cam && cam->w > 0 && cam->h > 0, "bad argument");
  MVS_TRY(ctx_ensure_device(h->ctx));
  int64_t nt = 0;
  DevBuf<int> scan_tmp;
  MVS_TRY(unproject_compact_f4(h->ctx, inv_depth, cam, min_dsp, max_dsp, h->tp_own, scan_tmp, &nt));
  if (n_target) *n_target = nt;
  MVS_TRY(target_common(h, nt));
  h->tp = h->tp_own.p;
  h->tn = nullptr;
  return set_target_index(h, nt);
}

int mvs_deform_set_target_dev_f32(mvs_deform* h, const float* tpts_dev, const float* tnormals_dev,
                                  int64_t nt) {
  MVS_REQUIRE(h && tpts_dev, "null argument");
  MVS_TRY(target_common(h, nt));
  h->tp = reinterpret_cast<const float4*>(tpts_dev);
  h->tn = reinterpret_cast<const float4*>(tnormals_dev);
  return set_target_index(h, nt);
}

int mvs_deform_set_targets(mvs_deform* h, const double* targets, const float* weights, int64_t n) {
  MVS_REQUIRE(h && targets, "null argument");
  if (n != h->n) {
    set_error("set_targets: expected %lld targets, got %lld", (long long)h->n, (long long)n);
    return MVS_ERR_INVALID_ARG;
  }
  mvs_ctx* ctx = h->ctx;
  MVS_TRY(ctx_ensure_device(ctx));
  MVS_TRY(ensure_graph(h));
  float wmax = 1.0f;
  if (weights)
    for (int64_t i = 0; i < n; ++i) {
      if (!(weights[i] >= 0.f) || !std::isfinite(weights[i])) {
        set_error("set_targets: weight %lld is negative or not finite", (long long)i);
        return MVS_ERR_INVALID_ARG;
      }
      wmax = std::max(wmax, weights[i]);
    }
  h->omega_max = wmax;  // bounds the fixed-point accumulation of the assembly
  double* dt = nullptr;
  float* dw = nullptr;
  MVS_CUDA(cudaMallocAsync((void**)&dt, sizeof(double) * 3 * n, ctx->stream));
  int st = h2d(ctx, dt, targets, sizeof(double) * 3 * n);
  if (st == MVS_OK && weights) {
    cudaError_t e = cudaMallocAsync((void**)&dw, sizeof(float) * n, ctx->stream);
    if (e != cudaSuccess) st = MVS_ERR_CUDA;
    if (st == MVS_OK) st = h2d(ctx, dw, weights, sizeof(float) * n);
  }
  if (st == MVS_OK) {
    ProfScope ps(ctx, "targets_permute");
    k_targets_permute<<<div_up(n, 256), 256, 0, ctx->stream>>>(dt, dw, h->perm.p, n, h->targets.p);
  }
  cudaFreeAsync(dt, ctx->stream);
  if (dw) cudaFreeAsync(dw, ctx->stream);
  MVS_TRY(st);
  MVS_CUDA(cudaStreamSynchronize(ctx->stream));
  h->have_targets = true;
  return MVS_OK;
}

int mvs_deform_node_targets(mvs_deform* h, double proj_len_err, double proj_dist_err,
                            int smooth_sweeps, double* controls_xyz, int32_t* valid) {
  MVS_REQUIRE(h, "null handle");
  MVS_TRY(ctx_ensure_device(h->ctx));
  MVS_TRY(ensure_graph(h));
  DevBuf<double> c;
  DevBuf<int> v;
  MVS_TRY(deform_node_targets(h, proj_len_err, proj_dist_err, smooth_sweeps, c, v));
  if (controls_xyz)
    MVS_CUDA(cudaMemcpyAsync(controls_xyz, c.p, sizeof(double) * 3 * h->m, cudaMemcpyDeviceToHost,
                             h->ctx->stream));
  if (valid)
    MVS_CUDA(cudaMemcpyAsync(valid, v.p, sizeof(int) * h->m, cudaMemcpyDeviceToHost,
                             h->ctx->stream));
  MVS_CUDA(cudaStreamSynchronize(h->ctx->stream));
  return MVS_OK;
}

int mvs_deform_gn_begin(mvs_deform* h, const mvs_gn_params* params) {
  MVS_REQUIRE(h, "null handle");
  MVS_TRY(ctx_ensure_device(h->ctx));
  return gn_begin_impl(h, params);
}

int mvs_deform_gn_step(mvs_deform* h) {
  MVS_REQUIRE(h, "null handle");
  if (!h->gn_active) {
    set_error("gn_step before gn_begin");
    return MVS_ERR_STATE;
  }
  MVS_TRY(ctx_ensure_device(h->ctx));
  return gn_step_impl(h);
}

int mvs_deform_solve(mvs_deform* h, const mvs_gn_params* params) {
  MVS_REQUIRE(h, "null handle");
  mvs_gn_params def;
  if (!params) {
    mvs_gn_default_params(&def);
    params = &def;
  }
  MVS_TRY(ctx_ensure_device(h->ctx));
  MVS_TRY(gn_begin_impl(h, params));
  for (int it = 0; it < params->gn_iters; ++it) {
    int rc = gn_step_impl(h);
    if (rc != MVS_OK) {
      h->gn_active = false;
      return rc;
    }
  }
  MVS_TRY(gn_final_vertices(h));
  MVS_CUDA(cudaStreamSynchronize(h->ctx->stream));
  h->gn_active = false;
  return check_solver_flag(h);
}

int mvs_deform_get_correspondences(mvs_deform* h, int32_t* out_idx) {
  MVS_REQUIRE(h && out_idx, "null argument");
  mvs_ctx* ctx = h->ctx;
  MVS_TRY(ctx_ensure_device(ctx));
  if (!h->graph_built || !h->corr_prev.p) {
    set_error("get_correspondences: graph not built");
    return MVS_ERR_STATE;
  }
  int* tmp = nullptr;
  MVS_CUDA(cudaMallocAsync((void**)&tmp, sizeof(int) * h->n, ctx->stream));
  k_unpermute_i32<<<div_up(h->n, 256), 256, 0, ctx->stream>>>(h->corr_prev.p, h->perm.p, h->n, tmp);
  int rc = d2h_async(ctx, out_idx, tmp, sizeof(int) * h->n);
  cudaFreeAsync(tmp, ctx->stream);
  MVS_TRY(rc);
  MVS_CUDA(cudaStreamSynchronize(ctx->stream));
  return MVS_OK;
}

int mvs_deform_gn_end(mvs_deform* h) {
  MVS_REQUIRE(h, "null handle");
  if (!h->gn_active) {
    set_error("gn_end before gn_begin");
    return MVS_ERR_STATE;
  }
  MVS_TRY(ctx_ensure_device(h->ctx));
  MVS_TRY(gn_final_vertices(h));
  MVS_CUDA(cudaStreamSynchronize(h->ctx->stream));
  h->gn_active = false;
  return check_solver_flag(h);
}

int mvs_deform_run(mvs_deform* h, double proj_len_err, double proj_dist_err,
                   const mvs_gn_params* params) {
  MVS_REQUIRE(h, "null handle");
  mvs_gn_params def;
  if (!params) {
    mvs_gn_default_params(&def);
    params = &def;
  }
  MVS_TRY(validate_params(params));
  MVS_TRY(ctx_ensure_device(h->ctx));
  int rounds = params->outer_rounds > 0 ? params->outer_rounds : 1;
  for (int round = 0; round < rounds; ++round) {  // Deformation.cpp:252-253 `while(counter--)`
    MVS_TRY(ensure_graph(h));
    if (params->corr_mode == 0) {
      DevBuf<double> c;
      DevBuf<int> v;
      // Deformation.cpp:266-381; every node becomes a constraint, valid or not (:384 commented)
      MVS_TRY(deform_node_targets(h, proj_len_err, proj_dist_err, 2, c, v));
      MVS_TRY(set_node_targets(h, c, v, 0));
    }
    MVS_TRY(gn_begin_impl(h, params));
    for (int it = 0; it < params->gn_iters; ++it) MVS_TRY(gn_step_impl(h));
    MVS_TRY(gn_final_vertices(h));
    h->gn_active = false;
    MVS_TRY(bake(h));
  }
  MVS_CUDA(cudaStreamSynchronize(h->ctx->stream));
  return MVS_OK;
}

static int node_targets_threshold_impl(mvs_deform* h, double threshold, DevBuf<double>& c,
                                       DevBuf<int>& v) {
  mvs_ctx* ctx = h->ctx;
  if (!h->tgt_index || !h->tgt_index->pyr_built) {
    set_error("run_threshold: target not set");
    return MVS_ERR_STATE;
  }
  if (h->m <= 0) {
    set_error("run_threshold: no nodes (the reference leaves sampIdx empty here, "
              "Deformation.cpp:184)");
    return MVS_ERR_STATE;
  }
  MVS_TRY(ensure_graph(h));
  MVS_TRY(c.reserve((size_t)h->m * 3));
  MVS_TRY(v.reserve((size_t)h->m));
  ProfScope ps(ctx, "node_targets_threshold");
  k_node_targets_threshold<<<div_up(h->m, 64), 64, 0, ctx->stream>>>(
      h->tgt_index->pyr, h->g.p, h->gnrm.p, h->m, threshold, c.p, v.p);
  MVS_CUDA(cudaGetLastError());
  return MVS_OK;
}

int mvs_deform_node_targets_threshold(mvs_deform* h, double threshold, double* controls_xyz,
                                      int32_t* valid) {
  MVS_REQUIRE(h, "null handle");
  MVS_TRY(ctx_ensure_device(h->ctx));
  DevBuf<double> c;
  DevBuf<int> v;
  MVS_TRY(node_targets_threshold_impl(h, threshold, c, v));
  if (controls_xyz)
    MVS_CUDA(cudaMemcpyAsync(controls_xyz, c.p, sizeof(double) * 3 * h->m, cudaMemcpyDeviceToHost,
                             h->ctx->stream));
  if (valid)
    MVS_CUDA(cudaMemcpyAsync(valid, v.p, sizeof(int) * h->m, cudaMemcpyDeviceToHost,
                             h->ctx->stream));
  MVS_CUDA(cudaStreamSynchronize(h->ctx->stream));
  return MVS_OK;
}

int mvs_deform_run_threshold(mvs_deform* h, double threshold, const mvs_gn_params* params) {
  MVS_REQUIRE(h, "null handle");
  mvs_gn_params def;
  if (!params) {
    mvs_gn_default_params(&def);
    params = &def;
  }
  MVS_TRY(validate_params(params));
  mvs_ctx* ctx = h->ctx;
  MVS_TRY(ctx_ensure_device(ctx));
  DevBuf<double> c;
  DevBuf<int> v;
  MVS_TRY(node_targets_threshold_impl(h, threshold, c, v));
  MVS_TRY(set_node_targets(h, c, v, 1));
  mvs_gn_params p = *params;
  p.corr_mode = 2;
  MVS_TRY(gn_begin_impl(h, &p));
  for (int it = 0; it < p.gn_iters; ++it) MVS_TRY(gn_step_impl(h));
  MVS_TRY(gn_final_vertices(h));
  h->gn_active = false;
  MVS_TRY(bake(h));
  MVS_CUDA(cudaStreamSynchronize(ctx->stream));
  return MVS_OK;
}

int mvs_arap_default_params(mvs_arap_params* p) {
  if (!p) return MVS_ERR_INVALID_ARG;
  p->iters = 5;       // Deformation.cpp:398 deform(5, 1e-4)
  p->cg_iters = 400;
  p->tol = 1e-4;
  p->cg_tol = 1e-6;
  return MVS_OK;
}

static int validate_arap(const mvs_arap_params* p) {
  if (p->iters < 1 || p->cg_iters < 1 || !(p->tol >= 0.0) || !(p->cg_tol >= 0.0)) {
    set_error("arap params: iters and cg_iters must be >= 1, tolerances >= 0");
    return MVS_ERR_INVALID_ARG;
  }
  return MVS_OK;
}

int mvs_deform_arap(mvs_deform* h, const double* ctrl_xyz, const mvs_arap_params* params,
                    double* energy, int32_t* iterations_run) {
  MVS_REQUIRE(h, "null handle");
  mvs_arap_params def;
  if (!params) {
    mvs_arap_default_params(&def);
    params = &def;
  }
  MVS_TRY(validate_arap(params));
  MVS_TRY(ctx_ensure_device(h->ctx));
  if (h->m <= 0 || !ctrl_xyz) {
    set_error("arap: no control vertices (sample nodes first and pass their target positions)");
    return MVS_ERR_STATE;
  }
  int it = 0;
  MVS_TRY(arap_solve(h, ctrl_xyz, h->node_ids.data(), h->m, params, energy, &it));
  if (iterations_run) *iterations_run = it;
  MVS_CUDA(cudaStreamSynchronize(h->ctx->stream));
  return MVS_OK;
}

int mvs_deform_run_arap(mvs_deform* h, double proj_len_err, double proj_dist_err,
                        const mvs_arap_params* params) {
  MVS_REQUIRE(h, "null handle");
  mvs_arap_params def;
  if (!params) {
    mvs_arap_default_params(&def);
    params = &def;
  }
  MVS_TRY(validate_arap(params));
  mvs_ctx* ctx = h->ctx;
  MVS_TRY(ctx_ensure_device(ctx));
  MVS_TRY(ensure_graph(h));
  DevBuf<double> c;
  DevBuf<int> v;
  // Deformation.cpp:266-381; every node becomes a control vertex, valid or not (:384 commented)
  MVS_TRY(deform_node_targets(h, proj_len_err, proj_dist_err, 2, c, v));
  std::vector<double> ctrl((size_t)h->m * 3);
  MVS_TRY(d2h_sync(ctx, ctrl.data(), c.p, sizeof(double) * 3 * h->m));
  MVS_TRY(arap_solve(h, ctrl.data(), h->node_ids.data(), h->m, params, nullptr, nullptr));
  MVS_TRY(bake(h));
  MVS_CUDA(cudaStreamSynchronize(ctx->stream));
  return MVS_OK;
}

int mvs_deform_get_energy(mvs_deform* h, double energy[3]) {
  MVS_REQUIRE(h && energy, "null argument");
  MVS_TRY(ctx_ensure_device(h->ctx));
  if (!h->graph_built) {
    set_error("get_energy: no solver state");
    return MVS_ERR_STATE;
  }
  MVS_TRY(gn_deform_and_correspond(h));
  return gn_energy(h, energy);
}

int mvs_deform_get_node_state(mvs_deform* h, float* A_t) {
  MVS_REQUIRE(h && A_t, "null argument");
  if (!h->graph_built) {
    set_error("get_node_state: graph not built");
    return MVS_ERR_STATE;
  }
  return d2h_sync(h->ctx, A_t, h->x.p, sizeof(float) * 12 * h->m);
}

int mvs_deform_set_node_state(mvs_deform* h, const float* A_t) {
  MVS_REQUIRE(h && A_t, "null argument");
  MVS_TRY(ensure_graph(h));
  MVS_TRY(h2d(h->ctx, h->x.p, A_t, sizeof(float) * 12 * h->m));
  MVS_CUDA(cudaStreamSynchronize(h->ctx->stream));
  h->cur_valid = false;
  return MVS_OK;
}

static int get_f4_as_xyz(mvs_deform* h, const float4* src, int64_t n, double* out64, float* out32) {
  mvs_ctx* ctx = h->ctx;
  size_t bytes = (out64 ? sizeof(double) : sizeof(float)) * 3 * n;
  DevBuf<char> tmp_buf;
  MVS_TRY(tmp_buf.reserve(bytes));
  void* tmp = tmp_buf.p;
  {
    ProfScope ps(ctx, "unpack_xyz");
    if (out64)
      k_f4_to_f64xyz<<<div_up(n, 256), 256, 0, ctx->stream>>>(src, (double*)tmp, n);
    else
      k_f4_to_f32xyz<<<div_up(n, 256), 256, 0, ctx->stream>>>(src, (float*)tmp, n);
  }
  int rc = d2h_async(ctx, out64 ? (void*)out64 : (void*)out32, tmp, bytes);
  if (rc != MVS_OK) return rc;
  MVS_CUDA(cudaStreamSynchronize(ctx->stream));
  return MVS_OK;
}

int mvs_deform_get_vertices(mvs_deform* h, double* pts) {
  MVS_REQUIRE(h && pts, "null argument");
  MVS_TRY(ctx_ensure_device(h->ctx));
  return get_f4_as_xyz(h, h->cur_valid ? h->cur0.p : h->v0.p, h->n, pts, nullptr);
}

int mvs_deform_get_vertices_f32(mvs_deform* h, float* pts) {
  MVS_REQUIRE(h && pts, "null argument");
  MVS_TRY(ctx_ensure_device(h->ctx));
  return get_f4_as_xyz(h, h->cur_valid ? h->cur0.p : h->v0.p, h->n, nullptr, pts);
}

int mvs_deform_get_normals(mvs_deform* h, double* normals) {
  MVS_REQUIRE(h && normals, "null argument");
  MVS_TRY(ctx_ensure_device(h->ctx));
  return get_f4_as_xyz(h, (h->cur_valid && h->nf > 0) ? h->curn0.p : h->nrm0.p, h->n, normals,
                       nullptr);
}

int mvs_deform_get_facets(mvs_deform* h, int32_t* facets, int64_t* n_facets) {
  MVS_REQUIRE(h, "null handle");
  if (n_facets) *n_facets = h->nf;
  if (facets && h->nf > 0) return d2h_sync(h->ctx, facets, h->facets.p, sizeof(int) * 3 * h->nf);
  return MVS_OK;
}

int mvs_deform_get_weights(mvs_deform* h, int32_t* node_idx, float* weights) {
  MVS_REQUIRE(h, "null handle");
  MVS_TRY(ctx_ensure_device(h->ctx));
  MVS_TRY(ensure_graph(h));
  if (node_idx)
    MVS_CUDA(cudaMemcpyAsync(node_idx, h->vidx_knn.p, sizeof(int4) * h->n, cudaMemcpyDeviceToHost,
                             h->ctx->stream));
  if (weights)
    MVS_CUDA(cudaMemcpyAsync(weights, h->vw_knn.p, sizeof(float4) * h->n, cudaMemcpyDeviceToHost,
                             h->ctx->stream));
  MVS_CUDA(cudaStreamSynchronize(h->ctx->stream));
  return MVS_OK;
}

int mvs_deform_get_graph(mvs_deform* h, int32_t* nbr_idx) {
  MVS_REQUIRE(h && nbr_idx, "null argument");
  MVS_TRY(ctx_ensure_device(h->ctx));
  MVS_TRY(ensure_graph(h));
  return d2h_sync(h->ctx, nbr_idx, h->graph.p, sizeof(int) * h->m * (h->graph_k + 1));
}

int mvs_deform_get_system(mvs_deform* h, int32_t* row_ptr, int32_t* col, float* blocks,
                          float* rot, float* rhs, int64_t* nnz) {
  MVS_REQUIRE(h, "null handle");
  if (!h->graph_built) {
    set_error("get_system: graph not built");
    return MVS_ERR_STATE;
  }
  mvs_ctx* ctx = h->ctx;
  MVS_TRY(ctx_ensure_device(ctx));
  if (nnz) *nnz = h->nnz;
  cudaStream_t st = ctx->stream;
  if (row_ptr)
    MVS_TRY(d2h_async(ctx, row_ptr, h->row_ptr.p, sizeof(int) * (h->m + 1)));
  if (col)
    MVS_TRY(d2h_async(ctx, col, h->col.p, sizeof(int) * h->nnz));
  if (blocks)
    MVS_TRY(d2h_async(ctx, blocks, h->B(), sizeof(float) * 16 * h->nnz));
  if (rot)
    MVS_TRY(d2h_async(ctx, rot, h->Rot(), sizeof(float) * 81 * h->m));
  if (rhs)
    MVS_TRY(d2h_async(ctx, rhs, h->rhs(), sizeof(float) * 12 * h->m));
  MVS_CUDA(cudaStreamSynchronize(st));
  return MVS_OK;
}

int mvs_deform_get_stats(mvs_deform* h, int64_t stats[6]) {
  MVS_REQUIRE(h && stats, "null argument");
  stats[0] = h->n;
  stats[1] = h->m;
  stats[2] = h->graph_built ? h->nnz : 0;
  stats[3] = h->graph_built ? h->nruns : 0;
  stats[4] = h->nf;
  stats[5] = h->nt;
  return MVS_OK;
}

int mvs_deform_set_shard(mvs_deform* h, int64_t v_begin, int64_t v_end) {
  MVS_REQUIRE(h, "null handle");
  if (v_begin < 0 || (v_end >= 0 && (v_end < v_begin || v_end > h->n))) {
    set_error("set_shard: bad range");
    return MVS_ERR_INVALID_ARG;
  }
  h->v_begin = v_begin;
  h->v_end = v_end;
  h->il_nranks = 0;
  return MVS_OK;
}

int mvs_deform_set_shard_interleaved(mvs_deform* h, int rank, int nranks, int chunks_per_rank) {
  MVS_REQUIRE(h, "null handle");
  if (nranks < 0 || (nranks > 0 && (rank < 0 || rank >= nranks)) || chunks_per_rank < 1 ||
      chunks_per_rank > 4096) {
    set_error("set_shard_interleaved: bad argument");
    return MVS_ERR_INVALID_ARG;
  }
  h->il_rank = rank;
  h->il_nranks = nranks;
  h->il_chunks = chunks_per_rank;
  return MVS_OK;
}

int mvs_deform_set_shard_parts(mvs_deform* h, int parts) {
  MVS_REQUIRE(h, "null handle");
  if (parts < 1 || parts > 1024) {
    set_error("set_shard_parts: parts must be in [1, 1024]");
    return MVS_ERR_INVALID_ARG;
  }
  h->shard_parts = parts;
  return MVS_OK;
}

// Deformation/Deformation.h:174-221 exportOBJ -> PlyObj/PlyObj.cpp:77-137 WriteObjCore format
int mvs_deform_export_obj(mvs_deform* h, const char* filename) {
  MVS_REQUIRE(h && filename, "null argument");
  int64_t n = h->n, nf = h->nf;
  std::vector<float> p((size_t)3 * n);
  std::vector<double> nr((size_t)3 * n);
  std::vector<int32_t> f((size_t)3 * nf);
  MVS_TRY(mvs_deform_get_vertices_f32(h, p.data()));
  MVS_TRY(mvs_deform_get_normals(h, nr.data()));
  MVS_TRY(mvs_deform_get_facets(h, f.data(), nullptr));
  FILE* fp = fopen(filename, "w");
  if (!fp) {
    set_error("exportOBJ: Error exporting %s", filename);
    return MVS_ERR_INVALID_ARG;
  }
  fprintf(fp, "####\n#\n# OBJ File Generated by MultiviewStitch Program\n#\n####\n");
  fprintf(fp, "# Object %s\n#\n# Vertices: %lld\n# Faces: %lld\n#\n####\n", filename, (long long)n,
          (long long)nf);
  for (int64_t i = 0; i < n; ++i) {
    fprintf(fp, "vn %g %g %g\n", (float)nr[3 * i], (float)nr[3 * i + 1], (float)nr[3 * i + 2]);
    fprintf(fp, "v %g %g %g\n", p[3 * i], p[3 * i + 1], p[3 * i + 2]);
  }
  fprintf(fp, "# %lld vertices, %lld vertices normals\n\n", (long long)n, (long long)n);
  for (int64_t i = 0; i < nf; ++i)
    fprintf(fp, "f %d//%d %d//%d %d//%d\n", f[3 * i] + 1, f[3 * i] + 1, f[3 * i + 1] + 1,
            f[3 * i + 1] + 1, f[3 * i + 2] + 1, f[3 * i + 2] + 1);
  fclose(fp);
  return MVS_OK;
}

}  // extern "C"
