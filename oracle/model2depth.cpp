// oracle/model2depth.cpp — TEST INFRASTRUCTURE ONLY (see oracle.h).
//
// Software restatement of Model2Depth (Model2Depth/Model2Depth.cpp:58-156,186-191 with
// Camera/Camera.cpp:6-38): the reference draws the mesh with fixed-function OpenGL into a w x h
// window, reads the depth buffer back and stores 1/z_eye per pixel, flipped to a top-left origin.
//
// What GL computes, restated without GL (derivation in DESIGN.md §Model2Depth):
//   * modelview = diag(1,-1,-1)[R|t] (Camera.cpp:6-13) and the frustum of Camera.cpp:15-38 map a
//     camera-space point (X,Y,Z) to window coordinates  x_w = fx X/Z + cx,  y_w = h - (fy Y/Z + cy);
//     after the vertical flip of Model2Depth.cpp:134 output pixel (i,j) samples the continuous image
//     position (u,v) = (i + 0.5, j + 0.5) — GL pixel centres — with u = fx X/Z + cx, v = fy Y/Z + cy.
//   * window depth is affine in 1/Z, and 1/Z is linear in screen space over a planar triangle, so the
//     stored value raw = 1/z_e (Model2Depth.cpp:135-139) is the screen-space barycentric
//     interpolation of the vertices' 1/Z; GL_LEQUAL with clear depth 1 keeps the nearest fragment =
//     the largest 1/Z; fragments outside (znear, zfar) are clipped (z_b <= 0 or >= 1, :134).
// Rasterisation rules fixed here (GL leaves them to the driver — "parity unpinned" vs a real GL):
//   vertex positions snapped to 1/256 pixel (round to nearest even), edge functions exact in int64,
//   a pixel centre is covered when all three edge functions are >= 0 (edges inclusive), both
//   windings drawn (no culling is enabled in InitGL), zero-area triangles dropped, triangles with a
//   vertex at Z < znear dropped instead of clipped, 1/Z kept in float32 (no 24-bit quantisation).
#include <cmath>
#include <cstring>
#include <vector>

#include "oracle.h"

namespace {

struct SV {
  int64_t x, y;  // 1/256 pixel
  double iz;
  bool ok;
};

const int64_t kLimit = (int64_t)1 << 29;

SV project_vertex(const orc_camera& c, const float* p, float znear) {
  // Model2Depth.h:35-39 casts the mesh to float; Camera.cpp:6-13 casts R,t to float
  double X = (double)(float)c.R[0] * p[0] + (double)(float)c.R[1] * p[1] + (double)(float)c.R[2] * p[2] + (double)(float)c.t[0];
  double Y = (double)(float)c.R[3] * p[0] + (double)(float)c.R[4] * p[1] + (double)(float)c.R[5] * p[2] + (double)(float)c.t[1];
  double Z = (double)(float)c.R[6] * p[0] + (double)(float)c.R[7] * p[1] + (double)(float)c.R[8] * p[2] + (double)(float)c.t[2];
  SV s;
  s.ok = Z >= (double)znear;
  if (!s.ok) {
    s.x = s.y = 0;
    s.iz = 0;
    return s;
  }
  double fx = (double)(float)c.K[0], cx = (double)(float)c.K[2];
  double fy = (double)(float)c.K[4], cy = (double)(float)c.K[5];
  double u = fx * X / Z + cx, v = fy * Y / Z + cy;
  double fxp = std::nearbyint(u * 256.0), fyp = std::nearbyint(v * 256.0);
  if (!(std::fabs(fxp) < (double)kLimit) || !(std::fabs(fyp) < (double)kLimit)) {
    s.ok = false;
    s.x = s.y = 0;
    s.iz = 0;
    return s;
  }
  s.x = (int64_t)fxp;
  s.y = (int64_t)fyp;
  s.iz = 1.0 / Z;
  return s;
}

int64_t ceil_div(int64_t a, int64_t b) { return a >= 0 ? (a + b - 1) / b : -((-a) / b); }
int64_t floor_div(int64_t a, int64_t b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }

}  // namespace

extern "C" {

void orc_model2depth(const float* pts, int64_t n, const int32_t* facets, int64_t nf,
                     const orc_camera* cams, int ncams, float znear, float zfar,
                     float* inv_depth) {
  if (!(znear > 0)) znear = 0.01f;   // Model2Depth.cpp:19
  if (!(zfar > 0)) zfar = 2000.0f;   // Model2Depth.cpp:20
  for (int ci = 0; ci < ncams; ++ci) {
    const orc_camera& c = cams[ci];
    int w = c.w, h = c.h;
    float* out = inv_depth + (size_t)ci * w * h;
    std::memset(out, 0, sizeof(float) * (size_t)w * h);
    std::vector<SV> sv((size_t)n);
    for (int64_t i = 0; i < n; ++i) sv[i] = project_vertex(c, pts + 3 * i, znear);
    float lo = 1.0f / zfar, hi = 1.0f / znear;
    for (int64_t f = 0; f < nf; ++f) {
      SV a = sv[facets[3 * f]], b = sv[facets[3 * f + 1]], d = sv[facets[3 * f + 2]];
      if (!a.ok || !b.ok || !d.ok) continue;
      int64_t area = (b.x - a.x) * (d.y - a.y) - (b.y - a.y) * (d.x - a.x);
      if (area == 0) continue;
      if (area < 0) {
        std::swap(b, d);
        area = -area;
      }
      int64_t minx = std::min(a.x, std::min(b.x, d.x)), maxx = std::max(a.x, std::max(b.x, d.x));
      int64_t miny = std::min(a.y, std::min(b.y, d.y)), maxy = std::max(a.y, std::max(b.y, d.y));
      int64_t i0 = std::max<int64_t>(0, ceil_div(minx - 128, 256));
      int64_t i1 = std::min<int64_t>(w - 1, floor_div(maxx - 128, 256));
      int64_t j0 = std::max<int64_t>(0, ceil_div(miny - 128, 256));
      int64_t j1 = std::min<int64_t>(h - 1, floor_div(maxy - 128, 256));
      for (int64_t j = j0; j <= j1; ++j)
        for (int64_t i = i0; i <= i1; ++i) {
          int64_t px = i * 256 + 128, py = j * 256 + 128;
          int64_t e0 = (d.x - b.x) * (py - b.y) - (d.y - b.y) * (px - b.x);  // opposite a
          int64_t e1 = (a.x - d.x) * (py - d.y) - (a.y - d.y) * (px - d.x);  // opposite b
          int64_t e2 = (b.x - a.x) * (py - a.y) - (b.y - a.y) * (px - a.x);  // opposite d
          if (e0 < 0 || e1 < 0 || e2 < 0) continue;
          double iz = ((double)e0 * a.iz + (double)e1 * b.iz + (double)e2 * d.iz) / (double)area;
          float fz = (float)iz;
          if (!(fz > lo && fz < hi)) continue;
          float& dst = out[(size_t)j * w + i];
          if (fz > dst) dst = fz;
        }
    }
  }
}

// Image3D/Image3D.cpp:92-106 with Camera/Camera.cpp:40-44,50-54,61-67
void orc_unproject_depth(const float* inv_depth, const orc_camera* cam, double min_dsp,
                         double max_dsp, double* out_xyz, int32_t* valid) {
  int w = cam->w, h = cam->h;
  for (int i = 0; i < w; ++i)
    for (int j = 0; j < h; ++j) {
      size_t k = (size_t)j * w + i;
      double dsp = (double)inv_depth[k];  // LoadDepth widens the float .raw (Common/Utils.h:170-173)
      if (dsp < min_dsp || dsp > max_dsp) {
        valid[k] = 0;
        out_xyz[3 * k] = out_xyz[3 * k + 1] = out_xyz[3 * k + 2] = 0.0;
        continue;
      }
      valid[k] = 1;
      orc_unproject(cam, i, j, 1.0 / dsp, out_xyz + 3 * k);
    }
}

}  // extern "C"
