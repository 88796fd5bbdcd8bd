// linalg3.h — 3x3 double-precision decompositions usable from host and device code.
//
// Replaces the Eigen calls on the path: JacobiSVD<Matrix3d> (Solver/SRTSolver.cpp:109,165),
// SelfAdjointEigenSolver (SetUtils/PointSetUtils.cpp:27), Matrix3d::inverse
// (Alignment/Alignment.cpp:154,299).  One thread does one decomposition; the kernels that call
// these are latency-trivial (a few hundred hypotheses at most).
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define MVS_HD __host__ __device__ __forceinline__
#else
#define MVS_HD inline
#endif

namespace mvs {

// All matrices row-major double[9].
MVS_HD void mat3_mul(const double* A, const double* B, double* C) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
MVS_HD void mat3_mul_bt(const double* A, const double* B, double* C) {  // C = A * B^T
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      C[3 * i + j] = A[3 * i] * B[3 * j] + A[3 * i + 1] * B[3 * j + 1] + A[3 * i + 2] * B[3 * j + 2];
}
MVS_HD double mat3_det(const double* A) {
  return A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) +
         A[2] * (A[3] * A[7] - A[4] * A[6]);
}
MVS_HD void mat3_inverse(const double* A, double* I) {
  double d = 1.0 / mat3_det(A);
  I[0] = (A[4] * A[8] - A[5] * A[7]) * d;
  I[1] = (A[2] * A[7] - A[1] * A[8]) * d;
  I[2] = (A[1] * A[5] - A[2] * A[4]) * d;
  I[3] = (A[5] * A[6] - A[3] * A[8]) * d;
  I[4] = (A[0] * A[8] - A[2] * A[6]) * d;
  I[5] = (A[2] * A[3] - A[0] * A[5]) * d;
  I[6] = (A[3] * A[7] - A[4] * A[6]) * d;
  I[7] = (A[1] * A[6] - A[0] * A[7]) * d;
  I[8] = (A[0] * A[4] - A[1] * A[3]) * d;
}

// One-sided Jacobi SVD: S = U diag(sig) V^T, sig descending.  Columns of the working copy are
// orthogonalised by plane rotations accumulated in V; U = normalised columns, completed by cross
// products when S is rank-deficient.
MVS_HD void svd3(const double* S, double* U, double* sig, double* V) {
  double A[9];
  for (int i = 0; i < 9; ++i) {
    A[i] = S[i];
    V[i] = (i % 4 == 0) ? 1.0 : 0.0;
  }
  for (int sweep = 0; sweep < 60; ++sweep) {
    bool rotated = false;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double alpha = 0, beta = 0, gamma = 0;
        for (int k = 0; k < 3; ++k) {
          alpha += A[3 * k + p] * A[3 * k + p];
          beta += A[3 * k + q] * A[3 * k + q];
          gamma += A[3 * k + p] * A[3 * k + q];
        }
        if (gamma == 0.0 || fabs(gamma) <= 1e-16 * sqrt(alpha * beta)) continue;
        rotated = true;
        double zeta = (beta - alpha) / (2.0 * gamma);
        double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
        for (int k = 0; k < 3; ++k) {
          double ap = A[3 * k + p], aq = A[3 * k + q];
          A[3 * k + p] = c * ap - s * aq;
          A[3 * k + q] = s * ap + c * aq;
          double vp = V[3 * k + p], vq = V[3 * k + q];
          V[3 * k + p] = c * vp - s * vq;
          V[3 * k + q] = s * vp + c * vq;
        }
      }
    if (!rotated) break;
  }
  double nrm[3];
  for (int j = 0; j < 3; ++j)
    nrm[j] = sqrt(A[j] * A[j] + A[3 + j] * A[3 + j] + A[6 + j] * A[6 + j]);
  // sort columns by descending norm (3-element network)
  int o0 = 0, o1 = 1, o2 = 2, tmp;
  if (nrm[o1] > nrm[o0]) { tmp = o0; o0 = o1; o1 = tmp; }
  if (nrm[o2] > nrm[o0]) { tmp = o0; o0 = o2; o2 = tmp; }
  if (nrm[o2] > nrm[o1]) { tmp = o1; o1 = o2; o2 = tmp; }
  int ord[3] = {o0, o1, o2};
  double Vs[9], Us[9];
  int have = 0;
  for (int j = 0; j < 3; ++j) {
    int src = ord[j];
    sig[j] = nrm[src];
    for (int k = 0; k < 3; ++k) Vs[3 * k + j] = V[3 * k + src];
    if (have == j && nrm[src] > 1e-300 && nrm[src] > 1e-14 * nrm[ord[0]]) {
      for (int k = 0; k < 3; ++k) Us[3 * k + j] = A[3 * k + src] / nrm[src];
      have = j + 1;
    }
  }
  for (int i = 0; i < 9; ++i) V[i] = Vs[i];
  if (have == 0) {
    Us[0] = 1.0; Us[3] = 0.0; Us[6] = 0.0;
    have = 1;
  }
  if (have == 1) {
    double e[3] = {0, 0, 0};
    if (fabs(Us[0]) < 0.9) e[0] = 1.0; else e[1] = 1.0;
    double c0 = Us[3] * e[2] - Us[6] * e[1];
    double c1 = Us[6] * e[0] - Us[0] * e[2];
    double c2 = Us[0] * e[1] - Us[3] * e[0];
    double n = sqrt(c0 * c0 + c1 * c1 + c2 * c2);
    Us[1] = c0 / n; Us[4] = c1 / n; Us[7] = c2 / n;
    have = 2;
  }
  if (have == 2) {
    double c0 = Us[3] * Us[7] - Us[6] * Us[4];
    double c1 = Us[6] * Us[1] - Us[0] * Us[7];
    double c2 = Us[0] * Us[4] - Us[3] * Us[1];
    double n = sqrt(c0 * c0 + c1 * c1 + c2 * c2);
    Us[2] = c0 / n; Us[5] = c1 / n; Us[8] = c2 / n;
  }
  for (int i = 0; i < 9; ++i) U[i] = Us[i];
}

// Kabsch rotation from the cross-covariance S (Solver/SRTSolver.cpp:109-120): R = V U^T, last
// singular direction flipped when |det R + 1| <= 1e-9; t = c2 - s R c1.
MVS_HD void kabsch_from_cov(const double* S, double scale, const double* c1, const double* c2,
                            double* R, double* t) {
  double U[9], V[9], sig[3];
  svd3(S, U, sig, V);
  mat3_mul_bt(V, U, R);
  if (fabs(mat3_det(R) + 1.0) <= 1e-9) {
    double Vf[9];
    for (int i = 0; i < 9; ++i) Vf[i] = V[i];
    Vf[2] = -Vf[2]; Vf[5] = -Vf[5]; Vf[8] = -Vf[8];
    mat3_mul_bt(Vf, U, R);
  }
  for (int i = 0; i < 3; ++i)
    t[i] = c2[i] - scale * (R[3 * i] * c1[0] + R[3 * i + 1] * c1[1] + R[3 * i + 2] * c1[2]);
}

// Symmetric eigen-decomposition (cyclic Jacobi): A = Q diag(w) Q^T, w ASCENDING, eigenvectors in
// the columns of Q.
MVS_HD void sym_eig3(const double* Ain, double* w, double* Q) {
  double A[9];
  for (int i = 0; i < 9; ++i) {
    A[i] = Ain[i];
    Q[i] = (i % 4 == 0) ? 1.0 : 0.0;
  }
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = A[1] * A[1] + A[2] * A[2] + A[5] * A[5];
    double dg = A[0] * A[0] + A[4] * A[4] + A[8] * A[8];
    if (off <= 1e-300 || off <= 1e-32 * dg) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double apq = A[3 * p + q];
        if (apq == 0.0) continue;
        double theta = (A[3 * q + q] - A[3 * p + p]) / (2.0 * apq);
        double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) {
          double akp = A[3 * k + p], akq = A[3 * k + q];
          A[3 * k + p] = c * akp - s * akq;
          A[3 * k + q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {
          double apk = A[3 * p + k], aqk = A[3 * q + k];
          A[3 * p + k] = c * apk - s * aqk;
          A[3 * q + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          double qkp = Q[3 * k + p], qkq = Q[3 * k + q];
          Q[3 * k + p] = c * qkp - s * qkq;
          Q[3 * k + q] = s * qkp + c * qkq;
        }
      }
  }
  w[0] = A[0]; w[1] = A[4]; w[2] = A[8];
  for (int i = 0; i < 2; ++i) {
    int m = i;
    for (int j = i + 1; j < 3; ++j)
      if (w[j] < w[m]) m = j;
    if (m != i) {
      double tw = w[i]; w[i] = w[m]; w[m] = tw;
      for (int k = 0; k < 3; ++k) {
        double tq = Q[3 * k + i]; Q[3 * k + i] = Q[3 * k + m]; Q[3 * k + m] = tq;
      }
    }
  }
}

}  // namespace mvs
