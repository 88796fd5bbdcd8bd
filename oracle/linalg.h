// oracle/linalg.h — TEST INFRASTRUCTURE ONLY (CPU oracle). Never linked into the product library.
//
// Small dense helpers in double precision.  The reference uses Eigen 3.2.6 (not vendored, absent
// here): JacobiSVD<Matrix3d> (Solver/SRTSolver.cpp:109,165), SelfAdjointEigenSolver
// (SetUtils/PointSetUtils.cpp:27), Matrix3d::inverse (Alignment/Alignment.cpp:154,299).  These are
// restated with textbook algorithms: one-sided (Hestenes) Jacobi for the SVD (U completed by cross
// products when rank-deficient), cyclic Jacobi for the symmetric eigenproblem, cofactor inverse.  tests/test_oracle_linalg.py pins them against numpy.
#pragma once
#include <cmath>
#include <cstring>
#include <utility>

namespace orc {

struct M3 {
  double a[3][3];
};
struct V3 {
  double x[3];
};

inline M3 m3_zero() {
  M3 m;
  std::memset(&m, 0, sizeof(m));
  return m;
}
inline M3 m3_identity() {
  M3 m = m3_zero();
  m.a[0][0] = m.a[1][1] = m.a[2][2] = 1.0;
  return m;
}
inline M3 m3_mul(const M3& A, const M3& B) {
  M3 C;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0.0;
      for (int k = 0; k < 3; ++k) s += A.a[i][k] * B.a[k][j];
      C.a[i][j] = s;
    }
  return C;
}
inline M3 m3_transpose(const M3& A) {
  M3 T;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) T.a[i][j] = A.a[j][i];
  return T;
}
inline double m3_det(const M3& A) {
  return A.a[0][0] * (A.a[1][1] * A.a[2][2] - A.a[1][2] * A.a[2][1]) -
         A.a[0][1] * (A.a[1][0] * A.a[2][2] - A.a[1][2] * A.a[2][0]) +
         A.a[0][2] * (A.a[1][0] * A.a[2][1] - A.a[1][1] * A.a[2][0]);
}
inline M3 m3_inverse(const M3& A) {
  double d = m3_det(A);
  M3 I;
  I.a[0][0] = (A.a[1][1] * A.a[2][2] - A.a[1][2] * A.a[2][1]) / d;
  I.a[0][1] = (A.a[0][2] * A.a[2][1] - A.a[0][1] * A.a[2][2]) / d;
  I.a[0][2] = (A.a[0][1] * A.a[1][2] - A.a[0][2] * A.a[1][1]) / d;
  I.a[1][0] = (A.a[1][2] * A.a[2][0] - A.a[1][0] * A.a[2][2]) / d;
  I.a[1][1] = (A.a[0][0] * A.a[2][2] - A.a[0][2] * A.a[2][0]) / d;
  I.a[1][2] = (A.a[0][2] * A.a[1][0] - A.a[0][0] * A.a[1][2]) / d;
  I.a[2][0] = (A.a[1][0] * A.a[2][1] - A.a[1][1] * A.a[2][0]) / d;
  I.a[2][1] = (A.a[0][1] * A.a[2][0] - A.a[0][0] * A.a[2][1]) / d;
  I.a[2][2] = (A.a[0][0] * A.a[1][1] - A.a[0][1] * A.a[1][0]) / d;
  return I;
}
inline V3 m3_mulv(const M3& A, const V3& v) {
  V3 r;
  for (int i = 0; i < 3; ++i) r.x[i] = A.a[i][0] * v.x[0] + A.a[i][1] * v.x[1] + A.a[i][2] * v.x[2];
  return r;
}
inline double v3_dot(const V3& a, const V3& b) {
  return a.x[0] * b.x[0] + a.x[1] * b.x[1] + a.x[2] * b.x[2];
}
inline V3 v3_cross(const V3& a, const V3& b) {
  V3 c;
  c.x[0] = a.x[1] * b.x[2] - a.x[2] * b.x[1];
  c.x[1] = a.x[2] * b.x[0] - a.x[0] * b.x[2];
  c.x[2] = a.x[0] * b.x[1] - a.x[1] * b.x[0];
  return c;
}
inline double v3_norm(const V3& a) { return std::sqrt(v3_dot(a, a)); }

// Symmetric 3x3 eigen-decomposition by cyclic Jacobi: A = Q diag(w) Q^T, w ascending
// (SelfAdjointEigenSolver order), eigenvectors in the columns of Q.
inline void sym_eig3(const M3& A_in, double w[3], M3& Q) {
  M3 A = A_in;
  Q = m3_identity();
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = A.a[0][1] * A.a[0][1] + A.a[0][2] * A.a[0][2] + A.a[1][2] * A.a[1][2];
    double diag = A.a[0][0] * A.a[0][0] + A.a[1][1] * A.a[1][1] + A.a[2][2] * A.a[2][2];
    if (off <= 1e-300 || off <= 1e-32 * diag) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double apq = A.a[p][q];
        if (apq == 0.0) continue;
        double theta = (A.a[q][q] - A.a[p][p]) / (2.0 * apq);
        double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) {  // A <- A J
          double akp = A.a[k][p], akq = A.a[k][q];
          A.a[k][p] = c * akp - s * akq;
          A.a[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {  // A <- J^T A
          double apk = A.a[p][k], aqk = A.a[q][k];
          A.a[p][k] = c * apk - s * aqk;
          A.a[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          double qkp = Q.a[k][p], qkq = Q.a[k][q];
          Q.a[k][p] = c * qkp - s * qkq;
          Q.a[k][q] = s * qkp + c * qkq;
        }
      }
  }
  for (int i = 0; i < 3; ++i) w[i] = A.a[i][i];
  // sort ascending (selection sort on 3 columns)
  for (int i = 0; i < 2; ++i) {
    int m = i;
    for (int j = i + 1; j < 3; ++j)
      if (w[j] < w[m]) m = j;
    if (m != i) {
      std::swap(w[i], w[m]);
      for (int k = 0; k < 3; ++k) std::swap(Q.a[k][i], Q.a[k][m]);
    }
  }
}

// SVD S = U diag(sig) V^T, sig descending, U and V orthogonal (full): one-sided (Hestenes) Jacobi
// on the columns of S, so small singular values keep their relative accuracy.
inline void svd3(const M3& S, M3& U, double sig[3], M3& V) {
  M3 A = S;
  V = m3_identity();
  for (int sweep = 0; sweep < 60; ++sweep) {
    bool rotated = false;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double alpha = 0, beta = 0, gamma = 0;
        for (int k = 0; k < 3; ++k) {
          alpha += A.a[k][p] * A.a[k][p];
          beta += A.a[k][q] * A.a[k][q];
          gamma += A.a[k][p] * A.a[k][q];
        }
        if (gamma == 0.0 || std::fabs(gamma) <= 1e-16 * std::sqrt(alpha * beta)) continue;
        rotated = true;
        double zeta = (beta - alpha) / (2.0 * gamma);
        double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
        double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
        for (int k = 0; k < 3; ++k) {
          double ap = A.a[k][p], aq = A.a[k][q];
          A.a[k][p] = c * ap - s * aq;
          A.a[k][q] = s * ap + c * aq;
          double vp = V.a[k][p], vq = V.a[k][q];
          V.a[k][p] = c * vp - s * vq;
          V.a[k][q] = s * vp + c * vq;
        }
      }
    if (!rotated) break;
  }
  double nrm[3];
  int order[3] = {0, 1, 2};
  for (int j = 0; j < 3; ++j)
    nrm[j] = std::sqrt(A.a[0][j] * A.a[0][j] + A.a[1][j] * A.a[1][j] + A.a[2][j] * A.a[2][j]);
  for (int i = 0; i < 2; ++i)
    for (int j = i + 1; j < 3; ++j)
      if (nrm[order[j]] > nrm[order[i]]) std::swap(order[i], order[j]);
  M3 Vs;
  V3 u[3];
  int have = 0;
  for (int j = 0; j < 3; ++j) {
    int src = order[j];
    sig[j] = nrm[src];
    for (int k = 0; k < 3; ++k) Vs.a[k][j] = V.a[k][src];
    if (have == j && nrm[src] > 1e-300 && nrm[src] > 1e-14 * nrm[order[0]]) {
      for (int k = 0; k < 3; ++k) u[j].x[k] = A.a[k][src] / nrm[src];
      have = j + 1;
    }
  }
  V = Vs;
  if (have == 0) {
    u[0] = V3{{1, 0, 0}};
    have = 1;
  }
  if (have == 1) {
    V3 e = std::fabs(u[0].x[0]) < 0.9 ? V3{{1, 0, 0}} : V3{{0, 1, 0}};
    V3 c = v3_cross(u[0], e);
    double n = v3_norm(c);
    for (int k = 0; k < 3; ++k) u[1].x[k] = c.x[k] / n;
    have = 2;
  }
  if (have == 2) {
    u[2] = v3_cross(u[0], u[1]);
    double n = v3_norm(u[2]);
    for (int k = 0; k < 3; ++k) u[2].x[k] /= n;
  }
  for (int j = 0; j < 3; ++j)
    for (int k = 0; k < 3; ++k) U.a[k][j] = u[j].x[k];
}

}  // namespace orc
