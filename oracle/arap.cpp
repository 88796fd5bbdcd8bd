// oracle/arap.cpp — TEST INFRASTRUCTURE ONLY (see oracle.h).
//
// As-rigid-as-possible surface deformation with hard positional constraints: the solver the
// reference runs through CGAL 4.6 `Surface_mesh_deformation` (Deformation/Deformation.cpp:256-260
// `insert_roi_vertices(all)`, :383-392 every sampled node a control vertex with a target position,
// :393 `preprocess()`, :398 `deform(5, 1e-4)`, :400 `overwrite_initial_geometry()`).  CGAL is not
// under /root/reference, so this restates the PUBLISHED algorithm (Sorkine & Alexa, "As-Rigid-As-
// Possible Surface Modeling", SGP 2007 — the method CGAL's ORIGINAL_ARAP mode implements) at the
// reference's call site; parity with CGAL's floating-point results is UNPINNED (no reference test
// pins them).  Specification shared with the CUDA path (csrc/arap.cu):
//   * weights  w_ij = max(0, 1/2 (cot a_ij + cot b_ij)) from the rest mesh (clamped so the system
//     stays positive semi-definite on obtuse triangulations);
//   * iteration k = 1..iters:  GLOBAL step: solve  sum_j w_ij (p'_i - p'_j) =
//     sum_j w_ij/2 (R_i + R_j)(p_i - p_j)  for the free vertices with the control vertices fixed at
//     their targets (R = I in the first iteration), by Jacobi-preconditioned conjugate gradients on
//     the three coordinate columns in lockstep, warm-started from the current positions;
//     LOCAL step:  S_i = sum_j w_ij (p_i - p_j)(p'_i - p'_j)^T = U S V^T,  R_i = V U^T (last
//     singular direction flipped when det < 0);  energy E = sum_i sum_j w_ij |e'_ij - R_i e_ij|^2;
//     stop when |E_prev - E| <= tol * E (tol > 0), as `deform(iterations, tolerance)` does.
#include <cmath>
#include <cstring>
#include <map>
#include <vector>

#include "linalg.h"
#include "oracle.h"

namespace {

struct Csr {
  std::vector<int> row_ptr, col;
  std::vector<double> w, diag;
};

Csr cot_laplacian(const float* pts, int64_t n, const int32_t* facets, int64_t nf) {
  // the same accumulation order as the device: half-edge contributions sorted by (i, j), equal
  // keys in emission order (facet-major), summed left to right, stored as float32
  std::vector<std::map<int, std::vector<float>>> acc((size_t)n);
  auto P = [&](int v, int a) { return (double)pts[3 * (size_t)v + a]; };
  for (int64_t f = 0; f < nf; ++f) {
    int v[3] = {facets[3 * f], facets[3 * f + 1], facets[3 * f + 2]};
    for (int c = 0; c < 3; ++c) {
      int o = v[c], a = v[(c + 1) % 3], b = v[(c + 2) % 3];
      double u[3], w[3];
      for (int k = 0; k < 3; ++k) {
        u[k] = P(a, k) - P(o, k);
        w[k] = P(b, k) - P(o, k);
      }
      double dot = u[0] * w[0] + u[1] * w[1] + u[2] * w[2];
      double cx = u[1] * w[2] - u[2] * w[1], cy = u[2] * w[0] - u[0] * w[2],
             cz = u[0] * w[1] - u[1] * w[0];
      double cr = std::sqrt(cx * cx + cy * cy + cz * cz);
      float half_cot = cr > 0.0 ? (float)(0.5 * dot / cr) : 0.0f;
      acc[(size_t)a][b].push_back(half_cot);
      acc[(size_t)b][a].push_back(half_cot);
    }
  }
  Csr L;
  L.row_ptr.assign((size_t)n + 1, 0);
  L.diag.assign((size_t)n, 0.0);
  for (int64_t i = 0; i < n; ++i) {
    for (auto& kv : acc[(size_t)i]) {
      float s = 0.f;
      for (float x : kv.second) s += x;
      float wv = s > 0.f ? s : 0.f;
      L.col.push_back(kv.first);
      L.w.push_back((double)wv);
    }
    L.row_ptr[(size_t)i + 1] = (int)L.col.size();
  }
  for (int64_t i = 0; i < n; ++i) {
    float d = 0.f;  // float32 running sum in column order, as the device does
    for (int t = L.row_ptr[i]; t < L.row_ptr[i + 1]; ++t) d += (float)L.w[t];
    L.diag[(size_t)i] = (double)d;
  }
  return L;
}

}  // namespace

extern "C" {

void orc_arap(const float* pts, int64_t n, const int32_t* facets, int64_t nf,
              const int32_t* ctrl_ids, const double* ctrl_xyz, int64_t m, int iters, double tol,
              int cg_iters, double cg_tol, double* out_xyz, double* energy_out, int32_t* iters_out) {
  Csr L = cot_laplacian(pts, n, facets, nf);
  std::vector<double> P((size_t)n * 3), X((size_t)n * 3), R((size_t)n * 9, 0.0);
  std::vector<char> fixed((size_t)n, 0);
  for (int64_t i = 0; i < n; ++i) {
    for (int a = 0; a < 3; ++a) P[3 * i + a] = X[3 * i + a] = (double)pts[3 * i + a];
    R[9 * i] = R[9 * i + 4] = R[9 * i + 8] = 1.0;
  }
  for (int64_t c = 0; c < m; ++c) {
    int v = ctrl_ids[c];
    if (v < 0 || v >= n) continue;
    fixed[(size_t)v] = 1;
    for (int a = 0; a < 3; ++a) X[3 * (size_t)v + a] = (double)(float)ctrl_xyz[3 * c + a];
  }
  std::vector<double> B((size_t)n * 3), r((size_t)n * 3), z((size_t)n * 3), p((size_t)n * 3),
      q((size_t)n * 3);
  double E_prev = 0.0, E = 0.0;
  int it = 0;
  for (; it < iters; ++it) {
    // ---- right-hand side
    for (int64_t i = 0; i < n; ++i) {
      double b[3] = {0, 0, 0};
      if (!fixed[(size_t)i])
        for (int t = L.row_ptr[i]; t < L.row_ptr[i + 1]; ++t) {
          int j = L.col[t];
          double e[3] = {P[3 * i] - P[3 * j], P[3 * i + 1] - P[3 * j + 1], P[3 * i + 2] - P[3 * j + 2]};
          for (int a = 0; a < 3; ++a) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += (R[9 * i + 3 * a + k] + R[9 * j + 3 * a + k]) * e[k];
            b[a] += 0.5 * L.w[t] * s;
          }
        }
      for (int a = 0; a < 3; ++a) B[3 * i + a] = b[a];
    }
    // ---- global step: Jacobi PCG, three columns in lockstep
    auto apply = [&](const std::vector<double>& x, std::vector<double>& y) {
      for (int64_t i = 0; i < n; ++i) {
        double s[3] = {0, 0, 0};
        if (!fixed[(size_t)i])
          for (int t = L.row_ptr[i]; t < L.row_ptr[i + 1]; ++t) {
            int j = L.col[t];
            for (int a = 0; a < 3; ++a) s[a] += L.w[t] * (x[3 * i + a] - x[3 * j + a]);
          }
        for (int a = 0; a < 3; ++a) y[3 * i + a] = s[a];
      }
    };
    apply(X, q);
    double rz[3] = {0, 0, 0}, bb[3] = {0, 0, 0}, rr[3] = {0, 0, 0};
    for (int64_t i = 0; i < n; ++i)
      for (int a = 0; a < 3; ++a) {
        double ri = fixed[(size_t)i] ? 0.0 : B[3 * i + a] - q[3 * i + a];
        double zi = L.diag[(size_t)i] > 0 ? ri / L.diag[(size_t)i] : 0.0;
        r[3 * i + a] = ri;
        z[3 * i + a] = zi;
        p[3 * i + a] = zi;
        rz[a] += ri * zi;
        bb[a] += ri * ri;
      }
    for (int a = 0; a < 3; ++a) rr[a] = bb[a];
    for (int k = 0; k < cg_iters; ++k) {
      if (cg_tol > 0.0 && rr[0] + rr[1] + rr[2] <= cg_tol * cg_tol * (bb[0] + bb[1] + bb[2])) break;
      apply(p, q);
      double pq[3] = {0, 0, 0};
      for (int64_t i = 0; i < n; ++i)
        for (int a = 0; a < 3; ++a) pq[a] += p[3 * i + a] * q[3 * i + a];
      double alpha[3], rz_new[3] = {0, 0, 0};
      for (int a = 0; a < 3; ++a) {
        alpha[a] = (pq[a] > 0.0 && rz[a] > 0.0) ? rz[a] / pq[a] : 0.0;
        rr[a] = 0.0;
      }
      for (int64_t i = 0; i < n; ++i)
        for (int a = 0; a < 3; ++a) {
          X[3 * i + a] += alpha[a] * p[3 * i + a];
          double ri = r[3 * i + a] - alpha[a] * q[3 * i + a];
          double zi = L.diag[(size_t)i] > 0 ? ri / L.diag[(size_t)i] : 0.0;
          r[3 * i + a] = ri;
          z[3 * i + a] = zi;
          rz_new[a] += ri * zi;
          rr[a] += ri * ri;
        }
      for (int a = 0; a < 3; ++a) {
        double beta = rz[a] > 0.0 ? rz_new[a] / rz[a] : 0.0;
        for (int64_t i = 0; i < n; ++i) p[3 * i + a] = z[3 * i + a] + beta * p[3 * i + a];
        rz[a] = rz_new[a];
      }
    }
    // ---- local step + energy
    E = 0.0;
    for (int64_t i = 0; i < n; ++i) {
      orc::M3 S = orc::m3_zero();
      for (int t = L.row_ptr[i]; t < L.row_ptr[i + 1]; ++t) {
        int j = L.col[t];
        for (int a = 0; a < 3; ++a)
          for (int b = 0; b < 3; ++b)
            S.a[a][b] += L.w[t] * (P[3 * i + a] - P[3 * j + a]) * (X[3 * i + b] - X[3 * j + b]);
      }
      orc::M3 U, V;
      double sig[3];
      orc::svd3(S, U, sig, V);
      orc::M3 Ri = orc::m3_mul(V, orc::m3_transpose(U));
      if (orc::m3_det(Ri) < 0.0) {
        for (int k = 0; k < 3; ++k) V.a[k][2] = -V.a[k][2];
        Ri = orc::m3_mul(V, orc::m3_transpose(U));
      }
      for (int a = 0; a < 3; ++a)
        for (int k = 0; k < 3; ++k) R[9 * i + 3 * a + k] = Ri.a[a][k];
      for (int t = L.row_ptr[i]; t < L.row_ptr[i + 1]; ++t) {
        int j = L.col[t];
        double d2 = 0;
        for (int a = 0; a < 3; ++a) {
          double re = 0;
          for (int k = 0; k < 3; ++k) re += Ri.a[a][k] * (P[3 * i + k] - P[3 * j + k]);
          double d = (X[3 * i + a] - X[3 * j + a]) - re;
          d2 += d * d;
        }
        E += L.w[t] * d2;
      }
    }
    if (tol > 0.0 && it > 0 && std::fabs(E_prev - E) <= tol * E) {
      ++it;
      break;
    }
    E_prev = E;
  }
  if (out_xyz) std::memcpy(out_xyz, X.data(), sizeof(double) * 3 * (size_t)n);
  if (energy_out) *energy_out = E;
  if (iters_out) *iters_out = it;
}

/* cotangent weights for tests: row_ptr (n+1), col / w (capacity entries); returns nnz */
int64_t orc_cot_weights(const float* pts, int64_t n, const int32_t* facets, int64_t nf,
                        int32_t* row_ptr, int32_t* col, double* w, int64_t capacity) {
  Csr L = cot_laplacian(pts, n, facets, nf);
  for (int64_t i = 0; i <= n; ++i) row_ptr[i] = L.row_ptr[(size_t)i];
  int64_t nnz = (int64_t)L.col.size();
  for (int64_t t = 0; t < nnz && t < capacity; ++t) {
    col[t] = L.col[(size_t)t];
    w[t] = L.w[(size_t)t];
  }
  return nnz;
}

}  // extern "C"
