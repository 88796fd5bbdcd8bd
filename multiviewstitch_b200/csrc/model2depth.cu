// model2depth.cu — CUDA z-buffer re-projection replacing Model2Depth's fixed-function OpenGL pass
// (Model2Depth/Model2Depth.cpp:58-156, Camera/Camera.cpp:6-38), the depth .raw wire format
// (Common/Utils.h:166-185) and the depth-map unprojection (Image3D/Image3D.cpp:92-106).
//
// One kernel per camera: a thread per triangle gathers its three vertices, transforms and projects
// them (float64, FMA-free, so coverage decisions are bit-reproducible on the host), snaps them to
// 1/256 pixel, walks the pixel centres of its bounding box with exact int64 edge functions and
// resolves visibility with atomicMax on the float32 bit pattern of 1/Z (positive floats order like
// unsigned integers; the cleared value 0 is "empty"), which is directly the output image — no
// separate resolve pass.  Triangles whose bounding box exceeds kBigBox pixels are queued and
// rasterised by one CTA each in a second kernel.
// Algorithmic bytes per view: 12 B/triangle (indices) + 16 B/vertex (each vertex is gathered ~6x but
// served from L2) + 8 B/pixel (clear + final image).
#include <stdio.h>

#include <cmath>
#include <vector>

#include "nn_grid.cuh"

namespace mvs {

struct CamP {
  double R[9], t[3];
  double fx, cx, fy, cy;
  int w, h;
  float lo, hi;  // 1/zfar, 1/znear
  double znear;
};

static CamP make_camp(const mvs_camera& c, float znear, float zfar) {
  CamP p;
  // Camera.cpp:6-13 / Model2Depth.h:35-39: GL sees float matrices
  for (int i = 0; i < 9; ++i) p.R[i] = (double)(float)c.R[i];
  for (int i = 0; i < 3; ++i) p.t[i] = (double)(float)c.t[i];
  p.fx = (double)(float)c.K[0];
  p.cx = (double)(float)c.K[2];
  p.fy = (double)(float)c.K[4];
  p.cy = (double)(float)c.K[5];
  p.w = c.w;
  p.h = c.h;
  p.lo = 1.0f / zfar;
  p.hi = 1.0f / znear;
  p.znear = (double)znear;
  return p;
}

struct SVert {
  long long x, y;
  double iz;
  bool ok;
};

__device__ __forceinline__ double dot3_rn(double a0, double b0, double a1, double b1, double a2,
                                          double b2, double c) {
  return __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(a0, b0), __dmul_rn(a1, b1)), __dmul_rn(a2, b2)), c);
}

__device__ __forceinline__ SVert project_vertex(const CamP& c, float4 p) {
  double x = p.x, y = p.y, z = p.z;
  double X = dot3_rn(c.R[0], x, c.R[1], y, c.R[2], z, c.t[0]);
  double Y = dot3_rn(c.R[3], x, c.R[4], y, c.R[5], z, c.t[1]);
  double Z = dot3_rn(c.R[6], x, c.R[7], y, c.R[8], z, c.t[2]);
  SVert s;
  s.x = s.y = 0;
  s.iz = 0.0;
  s.ok = Z >= c.znear;
  if (!s.ok) return s;
  double u = __dadd_rn(__ddiv_rn(__dmul_rn(c.fx, X), Z), c.cx);
  double v = __dadd_rn(__ddiv_rn(__dmul_rn(c.fy, Y), Z), c.cy);
  double fxp = rint(__dmul_rn(u, 256.0)), fyp = rint(__dmul_rn(v, 256.0));
  const double lim = 536870912.0;  // 2^29
  if (!(fabs(fxp) < lim) || !(fabs(fyp) < lim)) {
    s.ok = false;
    return s;
  }
  s.x = (long long)fxp;
  s.y = (long long)fyp;
  s.iz = __ddiv_rn(1.0, Z);
  return s;
}

__device__ __forceinline__ long long ceil_div256(long long a) {
  return a >= 0 ? (a + 255) / 256 : -((-a) / 256);
}
__device__ __forceinline__ long long floor_div256(long long a) {
  return a >= 0 ? a / 256 : -((-a + 255) / 256);
}

struct Tri {
  SVert a, b, d;
  long long area;
  int i0, i1, j0, j1;
};

__device__ __forceinline__ bool setup_tri(const CamP& c, const float4* __restrict__ pts,
                                          const int* __restrict__ facets, int64_t f, Tri& t) {
  int ia = __ldg(facets + 3 * f), ib = __ldg(facets + 3 * f + 1), id = __ldg(facets + 3 * f + 2);
  t.a = project_vertex(c, __ldg(pts + ia));
  t.b = project_vertex(c, __ldg(pts + ib));
  t.d = project_vertex(c, __ldg(pts + id));
  if (!t.a.ok || !t.b.ok || !t.d.ok) return false;
  t.area = (t.b.x - t.a.x) * (t.d.y - t.a.y) - (t.b.y - t.a.y) * (t.d.x - t.a.x);
  if (t.area == 0) return false;
  if (t.area < 0) {
    SVert tmp = t.b;
    t.b = t.d;
    t.d = tmp;
    t.area = -t.area;
  }
  long long minx = min(t.a.x, min(t.b.x, t.d.x)), maxx = max(t.a.x, max(t.b.x, t.d.x));
  long long miny = min(t.a.y, min(t.b.y, t.d.y)), maxy = max(t.a.y, max(t.b.y, t.d.y));
  long long i0 = max(0ll, ceil_div256(minx - 128)), i1 = min((long long)c.w - 1, floor_div256(maxx - 128));
  long long j0 = max(0ll, ceil_div256(miny - 128)), j1 = min((long long)c.h - 1, floor_div256(maxy - 128));
  if (i0 > i1 || j0 > j1) return false;
  t.i0 = (int)i0; t.i1 = (int)i1; t.j0 = (int)j0; t.j1 = (int)j1;
  return true;
}

__device__ __forceinline__ void shade(const CamP& c, const Tri& t, int i, int j,
                                      unsigned* __restrict__ fb) {
  long long px = (long long)i * 256 + 128, py = (long long)j * 256 + 128;
  long long e0 = (t.d.x - t.b.x) * (py - t.b.y) - (t.d.y - t.b.y) * (px - t.b.x);
  long long e1 = (t.a.x - t.d.x) * (py - t.d.y) - (t.a.y - t.d.y) * (px - t.d.x);
  long long e2 = (t.b.x - t.a.x) * (py - t.a.y) - (t.b.y - t.a.y) * (px - t.a.x);
  if (e0 < 0 || e1 < 0 || e2 < 0) return;
  double iz = __ddiv_rn(
      __dadd_rn(__dadd_rn(__dmul_rn((double)e0, t.a.iz), __dmul_rn((double)e1, t.b.iz)),
                __dmul_rn((double)e2, t.d.iz)),
      (double)t.area);
  float fz = __double2float_rn(iz);
  if (!(fz > c.lo && fz < c.hi)) return;
  atomicMax(fb + (size_t)j * c.w + i, __float_as_uint(fz));
}

constexpr int kBigBox = 64;

__global__ void __launch_bounds__(256)
    k_raster_small(CamP c, const float4* __restrict__ pts, const int* __restrict__ facets,
                   int64_t nf, unsigned* __restrict__ fb, int* __restrict__ big_list,
                   int* __restrict__ big_count) {
  int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; f < nf; f += stride) {
    Tri t;
    if (!setup_tri(c, pts, facets, f, t)) continue;
    int bw = t.i1 - t.i0 + 1, bh = t.j1 - t.j0 + 1;
    if ((long long)bw * bh > kBigBox) {
      big_list[atomicAdd(big_count, 1)] = (int)f;
      continue;
    }
    for (int j = t.j0; j <= t.j1; ++j)
      for (int i = t.i0; i <= t.i1; ++i) shade(c, t, i, j, fb);
  }
}

__global__ void __launch_bounds__(256)
    k_raster_big(CamP c, const float4* __restrict__ pts, const int* __restrict__ facets,
                 const int* __restrict__ big_list, const int* __restrict__ big_count,
                 unsigned* __restrict__ fb) {
  int nb = *big_count;
  for (int k = blockIdx.x; k < nb; k += gridDim.x) {
    Tri t;
    if (!setup_tri(c, pts, facets, big_list[k], t)) continue;
    int bw = t.i1 - t.i0 + 1, bh = t.j1 - t.j0 + 1;
    long long np = (long long)bw * bh;
    for (long long q = threadIdx.x; q < np; q += blockDim.x)
      shade(c, t, t.i0 + (int)(q % bw), t.j0 + (int)(q / bw), fb);
  }
}

static int render_views(mvs_ctx* ctx, const float4* pts, int64_t n, const int* facets, int64_t nf,
                        const mvs_camera* cams, int ncams, float znear, float zfar,
                        float* out_dev) {
  if (!(znear > 0.f)) znear = 0.01f;   // Model2Depth.cpp:19
  if (!(zfar > 0.f)) zfar = 2000.0f;   // Model2Depth.cpp:20
  cudaStream_t st = ctx->stream;
  DevBuf<int> big, cnt;
  MVS_TRY(big.reserve((size_t)(nf > 0 ? nf : 1)));
  MVS_TRY(cnt.reserve(1));
  for (int ci = 0; ci < ncams; ++ci) {
    const mvs_camera& c = cams[ci];
    size_t px = (size_t)c.w * c.h;
    unsigned* fb = reinterpret_cast<unsigned*>(out_dev + (size_t)ci * px);
    MVS_CUDA(cudaMemsetAsync(fb, 0, sizeof(float) * px, st));
    MVS_CUDA(cudaMemsetAsync(cnt.p, 0, sizeof(int), st));
    if (nf <= 0) continue;
    CamP cp = make_camp(c, znear, zfar);
    ProfScope ps(ctx, "model2depth", 2);
    k_raster_small<<<grid_for(ctx, nf, 256, 8), 256, 0, st>>>(cp, pts, facets, nf, fb, big.p, cnt.p);
    k_raster_big<<<ctx->num_sms * 4, 256, 0, st>>>(cp, pts, facets, big.p, cnt.p, fb);
  }
  MVS_CUDA(cudaGetLastError());
  return MVS_OK;
}

static int check_cams(const mvs_camera* cams, int ncams) {
  if (!cams || ncams <= 0) {
    set_error("model2depth: no cameras");
    return MVS_ERR_INVALID_ARG;
  }
  for (int i = 0; i < ncams; ++i)
    if (cams[i].w <= 0 || cams[i].h <= 0 || cams[i].w != cams[0].w || cams[i].h != cams[0].h) {
      set_error("model2depth: cameras must share a positive w x h (Model2Depth.h:47-50)");
      return MVS_ERR_INVALID_ARG;
    }
  return MVS_OK;
}

static int check_facets(const int32_t* facets, int64_t nf, int64_t n) {
  for (int64_t i = 0; i < 3 * nf; ++i)
    if (facets[i] < 0 || facets[i] >= n) {
      set_error("model2depth: facet index %d out of range", facets[i]);
      return MVS_ERR_INVALID_ARG;
    }
  return MVS_OK;
}

template <typename T>
static int model2depth_host(mvs_ctx* ctx, const T* pts, int64_t n, const int32_t* facets,
                            int64_t nf, const mvs_camera* cams, int ncams, float znear, float zfar,
                            float* inv_depth) {
  if (!ctx) ctx = default_ctx();
  if (!ctx) return MVS_ERR_NO_DEVICE;
  MVS_TRY(ctx_ensure_device(ctx));
  MVS_TRY(check_cams(cams, ncams));
  if (!inv_depth || n < 0 || nf < 0 || (n > 0 && !pts) || (nf > 0 && !facets)) {
    set_error("model2depth: bad argument");
    return MVS_ERR_INVALID_ARG;
  }
  MVS_TRY(check_facets(facets, nf, n));
  size_t px = (size_t)cams[0].w * cams[0].h;
  DevBuf<float4> dp;
  DevBuf<int> df;
  DevBuf<float> dout;
  if (n > 0) {
    if (sizeof(T) == 8)
      MVS_TRY(upload_xyz_f64(ctx, (const double*)pts, n, dp));
    else
      MVS_TRY(upload_xyz_f32(ctx, (const float*)pts, n, dp));
  }
  MVS_TRY(df.reserve((size_t)3 * nf + 1));
  if (nf > 0) MVS_TRY(h2d(ctx, df.p, facets, sizeof(int) * 3 * nf));
  MVS_TRY(dout.reserve(px * ncams));
  MVS_TRY(render_views(ctx, dp.p, n, df.p, nf, cams, ncams, znear, zfar, dout.p));
  return d2h_sync(ctx, inv_depth, dout.p, sizeof(float) * px * ncams);
}

// Image3D/Image3D.cpp:92-106 + Camera.cpp:40-44,61-67 (FMA-free double, as the host oracle)
__global__ void k_unproject(const float* __restrict__ dsp, int w, int h, double fx, double cx,
                            double fy, double cy, const double* __restrict__ Rt /*R(9) t(3)*/,
                            double min_dsp, double max_dsp, double* __restrict__ out,
                            int* __restrict__ valid) {
  int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= (int64_t)w * h) return;
  int i = (int)(k % w), j = (int)(k / w);
  double d = (double)dsp[k];
  if (d < min_dsp || d > max_dsp) {
    valid[k] = 0;
    out[3 * k] = out[3 * k + 1] = out[3 * k + 2] = 0.0;
    return;
  }
  valid[k] = 1;
  double z = __ddiv_rn(1.0, d);
  double c0 = __dsub_rn(__ddiv_rn(__dmul_rn(__dsub_rn((double)i, cx), z), fx), Rt[9]);
  double c1 = __dsub_rn(__ddiv_rn(__dmul_rn(__dsub_rn((double)j, cy), z), fy), Rt[10]);
  double c2 = __dsub_rn(z, Rt[11]);
#pragma unroll
  for (int a = 0; a < 3; ++a)
    out[3 * k + a] = __dadd_rn(__dadd_rn(__dmul_rn(Rt[a], c0), __dmul_rn(Rt[3 + a], c1)),
                               __dmul_rn(Rt[6 + a], c2));
}

// the same arithmetic for a list of pixels (u, v): out n x 3, valid n
__global__ void k_unproject_pixels(const float* __restrict__ dsp_at, const int* __restrict__ uv,
                                   int64_t n, double fx, double cx, double fy, double cy,
                                   const double* __restrict__ Rt, double min_dsp, double max_dsp,
                                   double* __restrict__ out, int* __restrict__ valid) {
  int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= n) return;
  int i = uv[2 * k], j = uv[2 * k + 1];
  double d = (double)dsp_at[k];
  if (d < min_dsp || d > max_dsp) {
    valid[k] = 0;
    out[3 * k] = out[3 * k + 1] = out[3 * k + 2] = 0.0;
    return;
  }
  valid[k] = 1;
  double z = __ddiv_rn(1.0, d);
  double c0 = __dsub_rn(__ddiv_rn(__dmul_rn(__dsub_rn((double)i, cx), z), fx), Rt[9]);
  double c1 = __dsub_rn(__ddiv_rn(__dmul_rn(__dsub_rn((double)j, cy), z), fy), Rt[10]);
  double c2 = __dsub_rn(z, Rt[11]);
#pragma unroll
  for (int a = 0; a < 3; ++a)
    out[3 * k + a] = __dadd_rn(__dadd_rn(__dmul_rn(Rt[a], c0), __dmul_rn(Rt[3 + a], c1)),
                               __dmul_rn(Rt[6 + a], c2));
}

// valid points in pixel order -> float4 (the rounding of k_xyz_f64_to_f4); prefix = exclusive scan of
// the valid flags with the total at prefix[n]
__global__ void k_compact_valid_f4(const double* __restrict__ xyz, const int* __restrict__ prefix,
                                   int64_t n, float4* __restrict__ out) {
  int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= n) return;
  int at = prefix[k];
  if (prefix[k + 1] == at) return;
  out[at] = make_float4((float)xyz[3 * k], (float)xyz[3 * k + 1], (float)xyz[3 * k + 2], 0.f);
}

// Inverse-depth image (host) -> the valid world points, compacted in pixel order, as float4 on the
// device: the unprojection of mvs_unproject_depth followed by what a caller does with its result
// before Deformation::Deform (keep the valid points, Processor.cpp:1122-1137) without the trip
// through host vectors.  *count = number of valid points.
int unproject_compact_f4(mvs_ctx* ctx, const float* inv_depth, const mvs_camera* cam, double min_dsp,
                         double max_dsp, DevBuf<float4>& out, DevBuf<int>& scan_tmp, int64_t* count) {
  size_t px = (size_t)cam->w * cam->h;
  DevBuf<float> din;
  DevBuf<double> dout, drt;
  DevBuf<int> dv;
  MVS_TRY(din.reserve(px));
  MVS_TRY(dout.reserve(px * 3));
  MVS_TRY(dv.reserve(px + 2));
  MVS_TRY(drt.reserve(12));
  MVS_TRY(out.reserve(px));
  double rt[12];
  for (int i = 0; i < 9; ++i) rt[i] = cam->R[i];
  for (int i = 0; i < 3; ++i) rt[9 + i] = cam->t[i];
  MVS_TRY(h2d(ctx, din.p, inv_depth, sizeof(float) * px));
  MVS_TRY(h2d(ctx, drt.p, rt, sizeof(rt)));
  {
    ProfScope ps(ctx, "unproject_depth");
    k_unproject<<<div_up((int64_t)px, 256), 256, 0, ctx->stream>>>(
        din.p, cam->w, cam->h, cam->K[0], cam->K[2], cam->K[4], cam->K[5], drt.p, min_dsp, max_dsp,
        dout.p, dv.p);
  }
  MVS_CUDA(cudaMemsetAsync(dv.p + px, 0, sizeof(int) * 2, ctx->stream));
  MVS_TRY(exclusive_scan_inplace(ctx, dv.p, (int64_t)px, scan_tmp));
  {
    ProfScope ps(ctx, "compact_valid");
    k_compact_valid_f4<<<div_up((int64_t)px, 256), 256, 0, ctx->stream>>>(dout.p, dv.p, (int64_t)px,
                                                                           out.p);
  }
  MVS_CUDA(cudaGetLastError());
  int total = 0;
  MVS_TRY(d2h_sync(ctx, &total, dv.p + px, sizeof(int)));  // rt[] and the temporaries stay alive
  *count = total;
  return MVS_OK;
}

}  // namespace mvs

using namespace mvs;

extern "C" {

int mvs_unproject_pixels(mvs_ctx* ctx, const float* inv_depth, const mvs_camera* cam, double min_dsp,
                         double max_dsp, const int32_t* uv, int64_t n, double* out_xyz,
                         int32_t* valid) {
  if (!ctx) ctx = default_ctx();
  if (!ctx) return MVS_ERR_NO_DEVICE;
  MVS_TRY(ctx_ensure_device(ctx));
  if (!inv_depth || !cam || !out_xyz || !valid || (n > 0 && !uv) || n < 0 || cam->w <= 0 ||
      cam->h <= 0) {
    set_error("unproject_pixels: bad argument");
    return MVS_ERR_INVALID_ARG;
  }
  if (n == 0) return MVS_OK;
  std::vector<float> at((size_t)n);
  for (int64_t k = 0; k < n; ++k) {
    int u = uv[2 * k], v = uv[2 * k + 1];
    if (u < 0 || v < 0 || u >= cam->w || v >= cam->h) {
      set_error("unproject_pixels: pixel %lld = (%d, %d) outside the %d x %d image", (long long)k, u,
                v, cam->w, cam->h);
      return MVS_ERR_INVALID_ARG;
    }
    at[(size_t)k] = inv_depth[(size_t)v * cam->w + u];
  }
  DevBuf<float> dat;
  DevBuf<int> duv, dv;
  DevBuf<double> dout, drt;
  MVS_TRY(dat.reserve((size_t)n));
  MVS_TRY(duv.reserve((size_t)n * 2));
  MVS_TRY(dv.reserve((size_t)n));
  MVS_TRY(dout.reserve((size_t)n * 3));
  MVS_TRY(drt.reserve(12));
  double rt[12];
  for (int i = 0; i < 9; ++i) rt[i] = cam->R[i];
  for (int i = 0; i < 3; ++i) rt[9 + i] = cam->t[i];
  MVS_TRY(h2d(ctx, dat.p, at.data(), sizeof(float) * n));
  MVS_TRY(h2d(ctx, duv.p, uv, sizeof(int) * 2 * n));
  MVS_TRY(h2d(ctx, drt.p, rt, sizeof(rt)));
  {
    ProfScope ps(ctx, "unproject_pixels");
    k_unproject_pixels<<<div_up(n, 128), 128, 0, ctx->stream>>>(
        dat.p, duv.p, n, cam->K[0], cam->K[2], cam->K[4], cam->K[5], drt.p, min_dsp, max_dsp, dout.p,
        dv.p);
  }
  MVS_CUDA(cudaGetLastError());
  MVS_TRY(d2h_async(ctx, out_xyz, dout.p, sizeof(double) * 3 * n));
  return d2h_sync(ctx, valid, dv.p, sizeof(int) * n);
}

int mvs_model2depth(mvs_ctx* ctx, const float* pts_xyz, int64_t n, const int32_t* facets,
                    int64_t n_facets, const mvs_camera* cams, int ncams, float znear, float zfar,
                    float* inv_depth) {
  return model2depth_host<float>(ctx, pts_xyz, n, facets, n_facets, cams, ncams, znear, zfar,
                                 inv_depth);
}

int mvs_model2depth_f64(mvs_ctx* ctx, const double* pts_xyz, int64_t n, const int32_t* facets,
                        int64_t n_facets, const mvs_camera* cams, int ncams, float znear,
                        float zfar, float* inv_depth) {
  return model2depth_host<double>(ctx, pts_xyz, n, facets, n_facets, cams, ncams, znear, zfar,
                                  inv_depth);
}

int mvs_model2depth_dev(mvs_ctx* ctx, const float* pts_xyzw_dev, int64_t n,
                        const int32_t* facets_dev, int64_t n_facets, const mvs_camera* cams,
                        int ncams, float znear, float zfar, float* inv_depth_dev) {
  if (!ctx) ctx = default_ctx();
  if (!ctx) return MVS_ERR_NO_DEVICE;
  MVS_TRY(ctx_ensure_device(ctx));
  MVS_TRY(check_cams(cams, ncams));
  if (!inv_depth_dev || n < 0 || n_facets < 0 || (n_facets > 0 && (!facets_dev || !pts_xyzw_dev))) {
    set_error("model2depth_dev: bad argument");
    return MVS_ERR_INVALID_ARG;
  }
  return render_views(ctx, reinterpret_cast<const float4*>(pts_xyzw_dev), n, facets_dev, n_facets,
                      cams, ncams, znear, zfar, inv_depth_dev);
}

// Common/Utils.h:177-185 SaveDepth: w*h little-endian float32, row-major, top-left origin
int mvs_save_depth_raw(const char* filename, const float* inv_depth, int w, int h) {
  if (!filename || !inv_depth || w <= 0 || h <= 0) {
    set_error("save_depth_raw: bad argument");
    return MVS_ERR_INVALID_ARG;
  }
  FILE* fp = fopen(filename, "wb");
  if (!fp) {
    set_error("save_depth_raw: cannot open %s", filename);
    return MVS_ERR_INVALID_ARG;
  }
  size_t nw = fwrite(inv_depth, sizeof(float), (size_t)w * h, fp);
  fclose(fp);
  if (nw != (size_t)w * h) {
    set_error("save_depth_raw: short write to %s", filename);
    return MVS_ERR_INVALID_ARG;
  }
  return MVS_OK;
}

// Common/Utils.h:166-176 LoadDepth
int mvs_load_depth_raw(const char* filename, float* inv_depth, int w, int h) {
  if (!filename || !inv_depth || w <= 0 || h <= 0) {
    set_error("load_depth_raw: bad argument");
    return MVS_ERR_INVALID_ARG;
  }
  FILE* fp = fopen(filename, "rb");
  if (!fp) {
    set_error("load_depth_raw: cannot open %s", filename);
    return MVS_ERR_INVALID_ARG;
  }
  size_t nr = fread(inv_depth, sizeof(float), (size_t)w * h, fp);
  fclose(fp);
  if (nr != (size_t)w * h) {
    set_error("load_depth_raw: %s holds %zu of %zu floats", filename, nr, (size_t)w * h);
    return MVS_ERR_INVALID_ARG;
  }
  return MVS_OK;
}

int mvs_unproject_depth(mvs_ctx* ctx, const float* inv_depth, const mvs_camera* cam,
                        double min_dsp, double max_dsp, double* out_xyz, int32_t* valid) {
  if (!ctx) ctx = default_ctx();
  if (!ctx) return MVS_ERR_NO_DEVICE;
  MVS_TRY(ctx_ensure_device(ctx));
  if (!inv_depth || !cam || !out_xyz || !valid || cam->w <= 0 || cam->h <= 0) {
    set_error("unproject_depth: bad argument");
    return MVS_ERR_INVALID_ARG;
  }
  size_t px = (size_t)cam->w * cam->h;
  DevBuf<float> din;
  DevBuf<double> dout, drt;
  DevBuf<int> dv;
  MVS_TRY(din.reserve(px));
  MVS_TRY(dout.reserve(px * 3));
  MVS_TRY(dv.reserve(px));
  MVS_TRY(drt.reserve(12));
  double rt[12];
  for (int i = 0; i < 9; ++i) rt[i] = cam->R[i];
  for (int i = 0; i < 3; ++i) rt[9 + i] = cam->t[i];
  MVS_TRY(h2d(ctx, din.p, inv_depth, sizeof(float) * px));
  MVS_TRY(h2d(ctx, drt.p, rt, sizeof(rt)));
  {
    ProfScope ps(ctx, "unproject_depth");
    k_unproject<<<div_up((int64_t)px, 256), 256, 0, ctx->stream>>>(
        din.p, cam->w, cam->h, cam->K[0], cam->K[2], cam->K[4], cam->K[5], drt.p, min_dsp, max_dsp,
        dout.p, dv.p);
  }
  MVS_CUDA(cudaGetLastError());
  MVS_TRY(d2h_async(ctx, out_xyz, dout.p, sizeof(double) * 3 * px));
  return d2h_sync(ctx, valid, dv.p, sizeof(int) * px);
}

}  // extern "C"
