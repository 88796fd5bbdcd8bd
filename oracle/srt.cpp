// oracle/srt.cpp — TEST INFRASTRUCTURE ONLY (see oracle.h).
//
// Restatement of Solver/SRTSolver.cpp and the Camera projection it uses, in the reference's own
// evaluation order (sequential double sums, left-to-right products).
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

#include "linalg.h"
#include "oracle.h"

using namespace orc;

namespace {

// (int)x with CUDA's saturating semantics made explicit (the C cast is UB outside int range).
inline int32_t trunc_to_int(double x) {
  if (x != x) return 0;
  if (x >= 2147483647.0) return 2147483647;
  if (x <= -2147483648.0) return (-2147483647 - 1);
  return (int32_t)x;
}

// Camera/Camera.cpp:68-72 GetCamCoordFromWorld
inline void cam_from_world(const orc_camera* c, const double p[3], double out[3]) {
  for (int i = 0; i < 3; ++i)
    out[i] = c->R[3 * i + 0] * p[0] + c->R[3 * i + 1] * p[1] + c->R[3 * i + 2] * p[2] + c->t[i];
}

// Camera/Camera.cpp:45-48 GetImgCoordFromCam: u = (int)(fx*x/z + cx + 0.5) (truncation)
inline void img_from_cam(const orc_camera* c, const double p[3], int32_t* u, int32_t* v) {
  *u = trunc_to_int(c->K[0] * p[0] / p[2] + c->K[2] + 0.5);
  *v = trunc_to_int(c->K[4] * p[1] / p[2] + c->K[5] + 0.5);
}

inline void centroids(const double* p, const double* q, int64_t n, double c1[3], double c2[3]) {
  for (int a = 0; a < 3; ++a) c1[a] = c2[a] = 0.0;
  for (int64_t i = 0; i < n; ++i)
    for (int a = 0; a < 3; ++a) {
      c1[a] = c1[a] + p[3 * i + a];
      c2[a] = c2[a] + q[3 * i + a];
    }
  for (int a = 0; a < 3; ++a) {
    c1[a] = c1[a] / (double)n;
    c2[a] = c2[a] / (double)n;
  }
}

// R = V U^T with the reflection fix of SRTSolver.cpp:114-119, t = c2 - s R c1 (:120)
inline void kabsch_from_S(const M3& S, double scale, const double c1[3], const double c2[3],
                          double R[9], double t[3]) {
  M3 U, V;
  double sig[3];
  svd3(S, U, sig, V);
  M3 Rm = m3_mul(V, m3_transpose(U));
  if (std::fabs(m3_det(Rm) + 1.0) <= 1e-9) {
    M3 Sg = m3_identity();
    Sg.a[2][2] = -1.0;
    Rm = m3_mul(m3_mul(V, Sg), m3_transpose(U));
  }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R[3 * i + j] = Rm.a[i][j];
  for (int i = 0; i < 3; ++i)
    t[i] = c2[i] - scale * (R[3 * i] * c1[0] + R[3 * i + 1] * c1[1] + R[3 * i + 2] * c1[2]);
}

inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

}  // namespace

extern "C" {

void orc_project(const orc_camera* cam, const double p[3], int32_t* u, int32_t* v) {
  double pc[3];
  cam_from_world(cam, p, pc);
  img_from_cam(cam, pc, u, v);
}

// Camera/Camera.cpp:40-44 GetCamCoordFromImg, :61-67 GetWorldCoordFromCam
void orc_unproject(const orc_camera* c, int u, int v, double d, double out[3]) {
  double pc[3];
  pc[0] = (u - c->K[2]) * d / c->K[0];
  pc[1] = (v - c->K[5]) * d / c->K[4];
  pc[2] = d;
  double tmp[3] = {pc[0] - c->t[0], pc[1] - c->t[1], pc[2] - c->t[2]};
  for (int j = 0; j < 3; ++j)
    out[j] = c->R[0 + j] * tmp[0] + c->R[3 + j] * tmp[1] + c->R[6 + j] * tmp[2];
}

// Solver/SRTSolver.cpp:6-29 ResidualError
double orc_srt_residual(const double* p, const double* q, int64_t n, const orc_camera* c1,
                        const orc_camera* c2, double scale, const double R[9],
                        const double t[3]) {
  double sR[9], iR[9];
  double inv_s = 1.0 / scale;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      sR[3 * i + j] = scale * R[3 * i + j];
      iR[3 * i + j] = inv_s * R[3 * j + i];
    }
  double err = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    const double* p1 = p + 3 * i;
    const double* p2 = q + 3 * i;
    double tp[3], tp_[3], d[3];
    for (int a = 0; a < 3; ++a)
      tp[a] = sR[3 * a] * p1[0] + sR[3 * a + 1] * p1[1] + sR[3 * a + 2] * p1[2] + t[a];
    int32_t u1, v1, u2, v2;
    orc_project(c2, tp, &u1, &v1);
    orc_project(c2, p2, &u2, &v2);
    for (int a = 0; a < 3; ++a) d[a] = p2[a] - t[a];
    for (int a = 0; a < 3; ++a)
      tp_[a] = iR[3 * a] * d[0] + iR[3 * a + 1] * d[1] + iR[3 * a + 2] * d[2];
    int32_t u1_, v1_, u2_, v2_;
    orc_project(c1, tp_, &u2_, &v2_);
    orc_project(c1, p1, &u1_, &v1_);
    double du = (double)u1 - (double)u2, dv = (double)v1 - (double)v2;
    double du_ = (double)u1_ - (double)u2_, dv_ = (double)v1_ - (double)v2_;
    double err1 = std::sqrt(du * du + dv * dv);
    double err2 = std::sqrt(du_ * du_ + dv_ * dv_);
    err = err + (err1 + err2) * 0.5;
  }
  return err / (double)n;
}

// Solver/SRTSolver.cpp:31-46 EstimateScale: mean of ratios of centroid distances
double orc_srt_scale(const double* p, const double* q, int64_t n) {
  double c1[3], c2[3];
  centroids(p, q, n, c1, c2);
  double scale = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    double a[3], b[3];
    for (int k = 0; k < 3; ++k) {
      a[k] = p[3 * i + k] - c1[k];
      b[k] = q[3 * i + k] - c2[k];
    }
    double na = std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    double nb = std::sqrt(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]);
    scale += nb / na;
  }
  return scale / (double)n;
}

// Solver/SRTSolver.cpp:65-129 EstimateRT
void orc_srt_rt(const double* p, const double* q, int64_t n, double scale, double R[9],
                double t[3]) {
  double c1[3], c2[3];
  centroids(p, q, n, c1, c2);
  M3 S = m3_zero();
  for (int64_t i = 0; i < n; ++i) {
    double x[3], y[3];
    for (int k = 0; k < 3; ++k) {
      x[k] = (p[3 * i + k] - c1[k]) * scale;
      y[k] = q[3 * i + k] - c2[k];
    }
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) S.a[a][b] += x[a] * y[b];
  }
  kabsch_from_S(S, scale, c1, c2, R, t);
}

// Solver/SRTSolver.cpp:272-275
void orc_srt_estimate(const double* p, const double* q, int64_t n, double* scale, double R[9],
                      double t[3]) {
  *scale = orc_srt_scale(p, q, n);
  orc_srt_rt(p, q, n, *scale, R, t);
}

// Solver/SRTSolver.cpp:277-280 + :131-185 EstimateRTRansac with injected sample triples
double orc_srt_ransac(const double* p, const double* q, int64_t n, const orc_camera* cam1,
                      const orc_camera* cam2, const int32_t* triples, int iter_num, double* scale,
                      double R[9], double t[3], int32_t* best_iter, double* all_errs) {
  double s = orc_srt_scale(p, q, n);
  *scale = s;
  double c1[3], c2[3];
  centroids(p, q, n, c1, c2);
  double err = HUGE_VAL;
  int32_t best = -1;
  for (int k = 0; k < iter_num; ++k) {
    M3 S = m3_zero();
    for (int j = 0; j < 3; ++j) {
      int64_t id = triples[3 * k + j];
      double x[3], y[3];
      for (int a = 0; a < 3; ++a) {
        x[a] = (p[3 * id + a] - c1[a]) * s;
        y[a] = q[3 * id + a] - c2[a];
      }
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) S.a[a][b] += x[a] * y[b];
    }
    double R_[9], t_[3];
    kabsch_from_S(S, s, c1, c2, R_, t_);
    double e = orc_srt_residual(p, q, n, cam1, cam2, s, R_, t_);
    if (all_errs) all_errs[k] = e;
    if (e < err) {
      err = e;
      best = k;
      std::memcpy(R, R_, sizeof(R_));
      std::memcpy(t, t_, sizeof(t_));
    }
  }
  if (best_iter) *best_iter = best;
  return err;
}

// Common/Utils.h:25-34 Shuffle(k, n, 3): m distinct indices out of n, kept sorted; rand() replaced
// by splitmix64(seed, iteration, draw).
void orc_draw_triples(int64_t n, int iter_num, uint64_t seed, int32_t* triples) {
  for (int it = 0; it < iter_num; ++it) {
    int32_t* k = triples + 3 * it;
    for (int i = 0; i < 3; i++) {
      uint64_t rnd = splitmix64(seed ^ splitmix64((uint64_t)it * 3u + (uint64_t)i + 1u));
      int64_t r = (int64_t)(rnd % (uint64_t)(n - i));
      int j;
      for (j = 0; j < i && r >= k[j]; j++) r++;
      int j0 = j;
      for (j = i; j > j0; j--) k[j] = k[j - 1];
      k[j0] = (int32_t)r;
    }
  }
}

void orc_svd3(const double S[9], double U[9], double sig[3], double V[9]) {
  M3 s, u, v;
  std::memcpy(s.a, S, sizeof(s.a));
  svd3(s, u, sig, v);
  std::memcpy(U, u.a, sizeof(u.a));
  std::memcpy(V, v.a, sizeof(v.a));
}

void orc_sym_eig3(const double A[9], double w[3], double Q[9]) {
  M3 a, q;
  std::memcpy(a.a, A, sizeof(a.a));
  sym_eig3(a, w, q);
  std::memcpy(Q, q.a, sizeof(q.a));
}

}  // extern "C"
