// pipeline.cu — the callers and data formats either side of the solver kernels:
//   Processor::RemoveOutliers round loop          Processor/Processor.cpp:177-269
//   Processor::CheckConsistencyCore               Processor/Processor.cpp:66-115
//   Depth2Model::SaveModel (depth map -> mesh)    Depth2Model/Depth2Model.cpp:7-82
#include <cmath>
#include <cstring>
#include <vector>

#include "nn_grid.cuh"

namespace mvs {

struct CamK {
  double fx, cx, fy, cy;
  double R[9], t[3];
  int w, h;
};

static CamK to_camk(const mvs_camera& c) {
  CamK k;
  k.fx = c.K[0];
  k.cx = c.K[2];
  k.fy = c.K[4];
  k.cy = c.K[5];
  for (int i = 0; i < 9; ++i) k.R[i] = c.R[i];
  for (int i = 0; i < 3; ++i) k.t[i] = c.t[i];
  k.w = c.w;
  k.h = c.h;
  return k;
}

// Camera/Camera.cpp:40-44, 61-67 (FMA-free double, as the host oracle)
__device__ __forceinline__ void unproject_rn(const CamK& c, int u, int v, double d, double out[3]) {
  double c0 = __dsub_rn(__ddiv_rn(__dmul_rn(__dsub_rn((double)u, c.cx), d), c.fx), c.t[0]);
  double c1 = __dsub_rn(__ddiv_rn(__dmul_rn(__dsub_rn((double)v, c.cy), d), c.fy), c.t[1]);
  double c2 = __dsub_rn(d, c.t[2]);
#pragma unroll
  for (int a = 0; a < 3; ++a)
    out[a] = __dadd_rn(__dadd_rn(__dmul_rn(c.R[a], c0), __dmul_rn(c.R[3 + a], c1)),
                       __dmul_rn(c.R[6 + a], c2));
}

// Camera/Camera.cpp:68-72, 45-48
__device__ __forceinline__ void project_k(const CamK& c, const double p[3], int& u, int& v) {
  double x = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(c.R[0], p[0]), __dmul_rn(c.R[1], p[1])),
                                 __dmul_rn(c.R[2], p[2])),
                       c.t[0]);
  double y = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(c.R[3], p[0]), __dmul_rn(c.R[4], p[1])),
                                 __dmul_rn(c.R[5], p[2])),
                       c.t[1]);
  double z = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(c.R[6], p[0]), __dmul_rn(c.R[7], p[1])),
                                 __dmul_rn(c.R[8], p[2])),
                       c.t[2]);
  u = __double2int_rz(__dadd_rn(__dadd_rn(__ddiv_rn(__dmul_rn(c.fx, x), z), c.cx), 0.5));
  v = __double2int_rz(__dadd_rn(__dadd_rn(__ddiv_rn(__dmul_rn(c.fy, y), z), c.cy), 0.5));
}

// ------------------------------------------------------------------ Depth2Model
__global__ void k_d2m_valid(const float* __restrict__ dsp, int64_t npx, double mn, double mx,
                            int* __restrict__ flag) {
  int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= npx) return;
  double d = (double)dsp[k];
  flag[k] = (d > 0.0 && !(d > mx || d < mn)) ? 1 : 0;  // Depth2Model.cpp:31-32
}

// tab[k] = exclusive scan of the valid flags: vertex id of pixel k when valid
__global__ void k_d2m_points(const float* __restrict__ dsp, const int* __restrict__ tab,
                             int64_t npx, CamK cam, double mn, double mx, double* __restrict__ xyz,
                             int* __restrict__ tex) {
  int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= npx) return;
  double d = (double)dsp[k];
  if (!(d > 0.0 && !(d > mx || d < mn))) return;
  int id = tab[k];
  if (tex) tex[id] = (int)k;
  if (xyz) {
    double p[3];
    unproject_rn(cam, (int)(k % cam.w), (int)(k / cam.w), __ddiv_rn(1.0, d), p);
    xyz[3 * (size_t)id] = p[0];
    xyz[3 * (size_t)id + 1] = p[1];
    xyz[3 * (size_t)id + 2] = p[2];
  }
}

struct QuadTris {
  bool t0, t1;
};
__device__ __forceinline__ QuadTris quad_tris(const float* __restrict__ dsp, int w, int x, int y,
                                              double mn, double mx, float thr) {
  auto ok = [&](int xx, int yy) {
    double d = (double)dsp[(size_t)yy * w + xx];
    return d > 0.0 && !(d > mx || d < mn);
  };
  QuadTris q{false, false};
  if (!ok(x, y) || !ok(x + 1, y + 1)) return q;
  double d00 = dsp[(size_t)y * w + x], d10 = dsp[(size_t)(y + 1) * w + x];
  double d11 = dsp[(size_t)(y + 1) * w + x + 1], d01 = dsp[(size_t)y * w + x + 1];
  double t = (double)thr;  // float threshold promoted in the comparisons (Depth2Model.cpp:45)
  q.t0 = ok(x, y + 1) && fabs(d00 - d10) <= t && fabs(d11 - d10) <= t && fabs(d00 - d11) <= t;
  q.t1 = ok(x + 1, y) && fabs(d00 - d01) <= t && fabs(d11 - d01) <= t && fabs(d11 - d00) <= t;
  return q;
}

__global__ void k_d2m_count(const float* __restrict__ dsp, int w, int h, double mn, double mx,
                            float thr, int* __restrict__ cnt) {
  int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t nq = (int64_t)(w - 1) * (h - 1);
  if (k >= nq) return;
  int x = (int)(k % (w - 1)), y = (int)(k / (w - 1));
  QuadTris q = quad_tris(dsp, w, x, y, mn, mx, thr);
  cnt[k] = (q.t0 ? 1 : 0) + (q.t1 ? 1 : 0);
}

__global__ void k_d2m_facets(const float* __restrict__ dsp, int w, int h, double mn, double mx,
                             float thr, const int* __restrict__ tab, const int* __restrict__ off,
                             int* __restrict__ facets) {
  int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t nq = (int64_t)(w - 1) * (h - 1);
  if (k >= nq) return;
  int x = (int)(k % (w - 1)), y = (int)(k / (w - 1));
  QuadTris q = quad_tris(dsp, w, x, y, mn, mx, thr);
  int o = off[k];
  int a = tab[(size_t)y * w + x], b = tab[(size_t)(y + 1) * w + x];
  int c = tab[(size_t)(y + 1) * w + x + 1], d = tab[(size_t)y * w + x + 1];
  if (q.t0) {
    facets[3 * (size_t)o] = a;
    facets[3 * (size_t)o + 1] = b;
    facets[3 * (size_t)o + 2] = c;
    ++o;
  }
  if (q.t1) {
    facets[3 * (size_t)o] = a;
    facets[3 * (size_t)o + 1] = c;
    facets[3 * (size_t)o + 2] = d;
  }
}

// ------------------------------------------------------------------ CheckConsistencyCore
__global__ void k_check_consistency(float* __restrict__ depth, CamK cur,
                                    const float* __restrict__ refd, const CamK* __restrict__ refc,
                                    int nref, double mn, double mx, double reproj) {
  int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int w = cur.w, h = cur.h;
  if (k >= (int64_t)w * h) return;
  int i = (int)(k % w), j = (int)(k / w);
  double dp = (double)depth[k];
  if (!(dp >= mn && dp <= mx)) {
    depth[k] = 0.f;
    return;
  }
  double p3d[3];
  unproject_rn(cur, i, j, __ddiv_rn(1.0, dp), p3d);
  for (int r = 0; r < nref; ++r) {
    CamK rc = refc[r];
    int u, v;
    project_k(rc, p3d, u, v);
    if (!(u >= 0 && u < rc.w && v >= 0 && v < rc.h)) {
      depth[k] = 0.f;
      return;
    }
    // the reference indexes the neighbour's map with the CURRENT frame's width (Processor.cpp:84)
    double rd = (double)refd[(size_t)r * w * h + (size_t)v * w + u];
    if (!(rd >= mn && rd <= mx)) {
      depth[k] = 0.f;
      return;
    }
    double q3d[3];
    unproject_rn(rc, u, v, __ddiv_rn(1.0, rd), q3d);
    project_k(cur, q3d, u, v);
    if (!(u >= 0 && u < w && v >= 0 && v < h)) {
      depth[k] = 0.f;
      return;
    }
    double e = sqrt((double)((i - u) * (i - u) + (j - v) * (j - v)));
    if (e > reproj) {
      depth[k] = 0.f;
      return;
    }
  }
}

// Depth2Model/Depth2Model.cpp:7-82 with the results left on the device (the public entry point
// copies them out; mvs_stitch_pairs keeps them there).  Null output buffers = counts only.
int depth2model_dev(mvs_ctx* ctx, const float* inv_depth, const mvs_camera* cam, double min_dsp,
                    double max_dsp, double smooth_threshold, DevBuf<double>* dxyz, DevBuf<int>* dtex,
                    DevBuf<int>* dfac, int64_t* n_points, int64_t* n_facets) {
  cudaStream_t st = ctx->stream;
  int w = cam->w, h = cam->h;
  int64_t npx = (int64_t)w * h, nq = (int64_t)(w - 1) * (h - 1);
  DevBuf<float> d;
  DevBuf<int> tab, cnt, scan_tmp;
  MVS_TRY(d.reserve((size_t)npx));
  MVS_TRY(tab.reserve((size_t)npx + 2));
  MVS_TRY(cnt.reserve((size_t)(nq > 0 ? nq : 0) + 2));
  MVS_TRY(h2d(ctx, d.p, inv_depth, sizeof(float) * npx));
  // Depth2Model.cpp:45: float threshold from double operands
  float thr = (float)(smooth_threshold * (max_dsp - min_dsp) / 100);
  {
    ProfScope ps(ctx, "d2m_valid");
    k_d2m_valid<<<div_up(npx, 256), 256, 0, st>>>(d.p, npx, min_dsp, max_dsp, tab.p);
  }
  MVS_CUDA(cudaMemsetAsync(tab.p + npx, 0, sizeof(int) * 2, st));
  MVS_TRY(exclusive_scan_inplace(ctx, tab.p, npx, scan_tmp));
  if (nq > 0) {
    {
      ProfScope ps(ctx, "d2m_count");
      k_d2m_count<<<div_up(nq, 256), 256, 0, st>>>(d.p, w, h, min_dsp, max_dsp, thr, cnt.p);
    }
    MVS_CUDA(cudaMemsetAsync(cnt.p + nq, 0, sizeof(int) * 2, st));
    MVS_TRY(exclusive_scan_inplace(ctx, cnt.p, nq, scan_tmp));
  }
  // both counts in one round trip
  int np = 0, nf = 0;
  MVS_CUDA(cudaMemcpyAsync(&np, tab.p + npx, sizeof(int), cudaMemcpyDeviceToHost, st));
  if (nq > 0) MVS_CUDA(cudaMemcpyAsync(&nf, cnt.p + nq, sizeof(int), cudaMemcpyDeviceToHost, st));
  MVS_CUDA(cudaStreamSynchronize(st));
  if (n_points) *n_points = np;
  if (n_facets) *n_facets = nf;
  if (np > 0 && (dxyz || dtex)) {
    if (dxyz) MVS_TRY(dxyz->reserve((size_t)np * 3));
    if (dtex) MVS_TRY(dtex->reserve((size_t)np));
    ProfScope ps(ctx, "d2m_points");
    k_d2m_points<<<div_up(npx, 256), 256, 0, st>>>(d.p, tab.p, npx, to_camk(*cam), min_dsp, max_dsp,
                                                   dxyz ? dxyz->p : nullptr, dtex ? dtex->p : nullptr);
  }
  if (nf > 0 && dfac) {
    MVS_TRY(dfac->reserve((size_t)nf * 3));
    ProfScope ps(ctx, "d2m_facets");
    k_d2m_facets<<<div_up(nq, 256), 256, 0, st>>>(d.p, w, h, min_dsp, max_dsp, thr, tab.p, cnt.p,
                                                  dfac->p);
  }
  MVS_CUDA(cudaGetLastError());
  // d, tab, cnt are released in stream order behind the kernels that read them
  return MVS_OK;
}

static int resolve(mvs_ctx*& ctx) {
  if (!ctx) ctx = default_ctx();
  if (!ctx) return MVS_ERR_NO_DEVICE;
  return ctx_ensure_device(ctx);
}

}  // namespace mvs

using namespace mvs;

extern "C" {

int mvs_outlier_default_params(mvs_outlier_params* p) {
  if (!p) return MVS_ERR_INVALID_ARG;
  p->rounds = 3;          // Processor.cpp:198
  p->iter_num = 200;      // Processor.cpp:202
  p->pixel_err = 55.0;    // Parameter/ParamParser.cpp:19
  p->adapt_ratio = 0.6;   // Parameter/ParamParser.cpp:20
  p->seed = 0;
  return MVS_OK;
}

int mvs_remove_outliers(mvs_ctx* ctx, const double* p_xyz, const double* q_xyz, int64_t n,
                        const mvs_camera* cam1, const mvs_camera* cam2,
                        const mvs_outlier_params* params, int32_t* keep_idx, int64_t* n_keep,
                        double* err, double* inlier_ratio) {
  return mvs_remove_outliers_samples(ctx, p_xyz, q_xyz, n, cam1, cam2, params, nullptr, nullptr,
                                     keep_idx, n_keep, err, inlier_ratio);
}

int mvs_remove_outliers_samples(mvs_ctx* ctx, const double* p_xyz, const double* q_xyz, int64_t n,
                                const mvs_camera* cam1, const mvs_camera* cam2,
                                const mvs_outlier_params* params, mvs_triples_fn draw, void* user,
                                int32_t* keep_idx, int64_t* n_keep, double* err,
                                double* inlier_ratio) {
  MVS_TRY(resolve(ctx));
  if (!p_xyz || !q_xyz || n < 0 || !cam1 || !cam2 || !keep_idx || !n_keep) {
    set_error("remove_outliers: bad argument");
    return MVS_ERR_INVALID_ARG;
  }
  mvs_outlier_params def;
  if (!params) {
    mvs_outlier_default_params(&def);
    params = &def;
  }
  std::vector<double> p(p_xyz, p_xyz + 3 * n), q(q_xyz, q_xyz + 3 * n);
  std::vector<int32_t> idx((size_t)n);
  for (int64_t i = 0; i < n; ++i) idx[i] = (int32_t)i;
  int64_t size = n;
  double e = HUGE_VAL, ratio = 1.0;
  std::vector<uint8_t> flags;
  std::vector<int32_t> triples;
  // one solver for all rounds (the reference constructs one per round, Processor.cpp:202; SetInput
  // replaces the matches): its device buffers are reused and destroyed — one stream wait — once
  mvs_srt* s = nullptr;
  MVS_TRY(mvs_srt_create(ctx, &s));
  struct Guard {
    mvs_srt* s;
    ~Guard() { mvs_srt_destroy(s); }
  } guard{s};
  for (int k = 0; k < params->rounds; ++k) {
    if (size < 3) break;  // the reference would index an empty match list (:131-146)
    int rc = mvs_srt_set_iteration_num(s, params->iter_num);
    if (rc == MVS_OK) rc = mvs_srt_set_print(s, 0);
    if (rc == MVS_OK) rc = mvs_srt_set_input(s, p.data(), q.data(), size, cam1, cam2);
    double sc = 1.0, R[9], t[3];
    const int32_t* samples = nullptr;  // null: the solver draws them from seed + round
    if (draw && params->iter_num > 0) {
      triples.resize((size_t)params->iter_num * 3);
      draw(size, params->iter_num, triples.data(), user);
      samples = triples.data();
    }
    if (rc == MVS_OK)
      rc = mvs_srt_estimate_ransac(s, samples, params->seed + (uint64_t)k, &sc, R, t, nullptr, nullptr);
    flags.resize((size_t)size);
    double err_all = 0.0;
    int64_t nin = 0;
    if (rc == MVS_OK)
      rc = mvs_srt_inliers(s, sc, R, t, params->pixel_err * ratio, flags.data(), &err_all, &nin);
    MVS_TRY(rc);
    int64_t m = 0;
    for (int64_t i = 0; i < size; ++i) {
      if (!flags[i]) continue;
      if (m != i) {
        memcpy(&p[3 * m], &p[3 * i], sizeof(double) * 3);
        memcpy(&q[3 * m], &q[3 * i], sizeof(double) * 3);
        idx[m] = idx[i];
      }
      ++m;
    }
    ratio *= params->adapt_ratio;
    e = err_all / (double)size;  // :241
    size = m;
    if (m < 3) break;
  }
  memcpy(keep_idx, idx.data(), sizeof(int32_t) * (size_t)size);
  *n_keep = size;
  if (err) *err = e;
  if (inlier_ratio) *inlier_ratio = 0.0;  // never updated in the reference (:243-253 commented out)
  return MVS_OK;
}

int mvs_depth2model(mvs_ctx* ctx, const float* inv_depth, const mvs_camera* cam, double min_dsp,
                    double max_dsp, double smooth_threshold, double* out_xyz, int32_t* tex_index,
                    int32_t* facets, int64_t* n_points, int64_t* n_facets) {
  MVS_TRY(resolve(ctx));
  if (!inv_depth || !cam || cam->w <= 0 || cam->h <= 0) {
    set_error("depth2model: bad argument");
    return MVS_ERR_INVALID_ARG;
  }
  DevBuf<double> dxyz;
  DevBuf<int> dtex, dfac;
  int64_t np = 0, nf = 0;
  MVS_TRY(depth2model_dev(ctx, inv_depth, cam, min_dsp, max_dsp, smooth_threshold,
                          out_xyz ? &dxyz : nullptr, tex_index ? &dtex : nullptr,
                          facets ? &dfac : nullptr, &np, &nf));
  if (n_points) *n_points = np;
  if (n_facets) *n_facets = nf;
  if (np > 0 && out_xyz) MVS_TRY(d2h_async(ctx, out_xyz, dxyz.p, sizeof(double) * 3 * np));
  if (np > 0 && tex_index) MVS_TRY(d2h_async(ctx, tex_index, dtex.p, sizeof(int) * np));
  if (nf > 0 && facets) MVS_TRY(d2h_async(ctx, facets, dfac.p, sizeof(int) * 3 * nf));
  MVS_CUDA(cudaStreamSynchronize(ctx->stream));
  return MVS_OK;
}

int mvs_check_consistency(mvs_ctx* ctx, float* depth, const mvs_camera* cam, const float* ref_depths,
                          const mvs_camera* ref_cams, int nref, double min_dsp, double max_dsp,
                          double reproj_err) {
  MVS_TRY(resolve(ctx));
  if (!depth || !cam || cam->w <= 0 || cam->h <= 0 || nref < 0 || (nref > 0 && (!ref_depths || !ref_cams))) {
    set_error("check_consistency: bad argument");
    return MVS_ERR_INVALID_ARG;
  }
  for (int r = 0; r < nref; ++r)
    if (ref_cams[r].w != cam->w || ref_cams[r].h != cam->h) {
      set_error("check_consistency: reference frames must share the frame size");
      return MVS_ERR_INVALID_ARG;
    }
  cudaStream_t st = ctx->stream;
  int64_t npx = (int64_t)cam->w * cam->h;
  DevBuf<float> d, rd;
  DevBuf<CamK> rc;
  MVS_TRY(d.reserve((size_t)npx));
  MVS_TRY(rd.reserve((size_t)npx * (nref > 0 ? nref : 1)));
  MVS_TRY(rc.reserve((size_t)(nref > 0 ? nref : 1)));
  std::vector<CamK> hk((size_t)nref);
  for (int r = 0; r < nref; ++r) hk[r] = to_camk(ref_cams[r]);
  MVS_TRY(h2d(ctx, d.p, depth, sizeof(float) * npx));
  if (nref > 0) {
    MVS_TRY(h2d(ctx, rd.p, ref_depths, sizeof(float) * npx * nref));
    MVS_TRY(h2d(ctx, rc.p, hk.data(), sizeof(CamK) * nref));
  }
  {
    ProfScope ps(ctx, "check_consistency");
    k_check_consistency<<<div_up(npx, 256), 256, 0, st>>>(d.p, to_camk(*cam), rd.p, rc.p, nref,
                                                          min_dsp, max_dsp, reproj_err);
  }
  MVS_CUDA(cudaGetLastError());
  return d2h_sync(ctx, depth, d.p, sizeof(float) * npx);
}

}  // extern "C"
