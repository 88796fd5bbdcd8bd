// oracle/deform.cpp — TEST INFRASTRUCTURE ONLY (see oracle.h).
//
// Deformation graph: node sampling, node graph, node correspondences and smoothing restate
// Deformation/Deformation.cpp:63-153 and :266-381.  The solver stage (CGAL ARAP,
// Deformation.cpp:383-400) is replaced — as north_star prescribes — by an embedded-deformation
// Gauss-Newton solve whose formulation is fixed in SURVEY.md §8(a) D7-D9 and DESIGN.md; there are no
// reference lines for it ("parity unpinned" against CGAL), so it is pinned by the finite-difference
// and scipy checks in tests/test_oracle_deform.py.
#ifdef _OPENMP
#include <omp.h>
#endif
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <map>
#include <queue>
#include <vector>

#include "oracle.h"

namespace {

inline float dist2_f32(const float* a, const float* b) {
  float dx = a[0] - b[0];
  float dy = a[1] - b[1];
  float dz = a[2] - b[2];
  float xx = dx * dx;
  float yy = dy * dy;
  float zz = dz * dz;
  float s = xx + yy;
  return s + zz;
}

// dense symmetric positive definite inverse (Cholesky), n <= 12
bool spd_inverse(const double* A, int n, double* inv) {
  double L[144];
  std::memset(L, 0, sizeof(L));
  for (int i = 0; i < n; ++i)
    for (int j = 0; j <= i; ++j) {
      double s = A[i * n + j];
      for (int k = 0; k < j; ++k) s -= L[i * n + k] * L[j * n + k];
      if (i == j) {
        if (!(s > 0)) return false;
        L[i * n + i] = std::sqrt(s);
      } else {
        L[i * n + j] = s / L[j * n + j];
      }
    }
  // inverse of L (lower triangular)
  double Li[144];
  std::memset(Li, 0, sizeof(Li));
  for (int c = 0; c < n; ++c) {
    Li[c * n + c] = 1.0 / L[c * n + c];
    for (int i = c + 1; i < n; ++i) {
      double s = 0;
      for (int k = c; k < i; ++k) s -= L[i * n + k] * Li[k * n + c];
      Li[i * n + c] = s / L[i * n + i];
    }
  }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      double s = 0;
      for (int k = std::max(i, j); k < n; ++k) s += Li[k * n + i] * Li[k * n + j];
      inv[i * n + j] = s;
    }
  return true;
}

// Normal equations in the Kronecker-compressed block form used by the build:
//   H = sum over pattern (j,l) of kron(B_jl (4x4), I3)  +  blockdiag(Rot_j (9x9) padded to 12)  + lambda I
struct System {
  int64_t m = 0;
  std::vector<int> row_ptr, col;
  std::vector<double> B;    // nnz x 16
  std::vector<double> Rot;  // m x 81
  std::vector<double> g;    // m x 12 (J^T r)
  double lambda = 0;
  int find(int j, int l) const {
    auto b = col.begin() + row_ptr[j], e = col.begin() + row_ptr[j + 1];
    auto it = std::lower_bound(b, e, l);
    return (it != e && *it == l) ? (int)(it - col.begin()) : -1;
  }
};

void build_pattern(int64_t n, int64_t m, const int32_t* widx, const int32_t* graph, int gk1,
                   System& S) {
  std::vector<std::vector<int>> rows((size_t)m);
  for (int64_t j = 0; j < m; ++j) rows[j].push_back((int)j);
  for (int64_t i = 0; i < n; ++i)
    for (int a = 0; a < 4; ++a)
      for (int b = 0; b < 4; ++b) {
        int j = widx[4 * i + a], l = widx[4 * i + b];
        if (j >= 0 && l >= 0) rows[j].push_back(l);
      }
  for (int64_t j = 0; j < m; ++j)
    for (int e = 0; e < gk1; ++e) {
      int l = graph[j * gk1 + e];
      if (l >= 0 && l != j) {
        rows[j].push_back(l);
        rows[l].push_back((int)j);
      }
    }
  S.m = m;
  S.row_ptr.assign((size_t)m + 1, 0);
  S.col.clear();
  for (int64_t j = 0; j < m; ++j) {
    auto& r = rows[j];
    std::sort(r.begin(), r.end());
    r.erase(std::unique(r.begin(), r.end()), r.end());
    S.row_ptr[j + 1] = S.row_ptr[j] + (int)r.size();
    S.col.insert(S.col.end(), r.begin(), r.end());
  }
  S.B.assign(S.col.size() * 16, 0.0);
  S.Rot.assign((size_t)m * 81, 0.0);
  S.g.assign((size_t)m * 12, 0.0);
}

struct Problem {
  const float* pts;
  int64_t n;
  std::vector<double> nodes;  // m x 3
  int64_t m;
  const int32_t* widx;
  const float* w;
  const int32_t* graph;
  int gk1;
};

void deform_vertex(const Problem& P, const double* x, int64_t i, double out[3]) {
  out[0] = out[1] = out[2] = 0.0;
  for (int a = 0; a < 4; ++a) {
    int j = P.widx[4 * i + a];
    if (j < 0) continue;
    double wj = P.w[4 * i + a];
    const double* g = &P.nodes[3 * (size_t)j];
    const double* X = x + 12 * (size_t)j;
    double d[3] = {P.pts[3 * i] - g[0], P.pts[3 * i + 1] - g[1], P.pts[3 * i + 2] - g[2]};
    for (int r = 0; r < 3; ++r) {
      double v = X[r] * d[0] + X[3 + r] * d[1] + X[6 + r] * d[2] + g[r] + X[9 + r];
      out[r] += wj * v;
    }
  }
}

// The correspondence QUERY point of vertex i: the same sum evaluated in float32 with the node state
// rounded to float32 (the device's state IS float32), nodes in ASCENDING node-id order, and one fixed
// sequence of roundings per component — product, two fused multiply-adds, two additions, one fused
// multiply-add with the weight (std::fmaf is correctly rounded, like the device's fma.rn.f32).  The
// device evaluates deform_point() (csrc/gn_assemble.cu) in exactly this order, so that "same node
// state in -> bit-identical query point -> bit-identical exact-NN index out".
void deform_vertex_f32(const Problem& P, const double* x, int64_t i, float out[3]) {
  int order[4] = {0, 1, 2, 3};
  auto key = [&](int a) {
    int j = P.widx[4 * i + a];
    return j < 0 ? 0x7fffffff : j;
  };
  std::sort(order, order + 4, [&](int a, int b) { return key(a) < key(b); });
  out[0] = out[1] = out[2] = 0.0f;
  for (int o = 0; o < 4; ++o) {
    int a = order[o];
    int j = P.widx[4 * i + a];
    if (j < 0) continue;
    float wj = P.w[4 * i + a];
    float g[3], X[12];
    for (int r = 0; r < 3; ++r) g[r] = (float)P.nodes[3 * (size_t)j + r];
    for (int q = 0; q < 12; ++q) X[q] = (float)x[12 * (size_t)j + q];
    float d0 = P.pts[3 * i] - g[0], d1 = P.pts[3 * i + 1] - g[1], d2 = P.pts[3 * i + 2] - g[2];
    for (int r = 0; r < 3; ++r) {
      float s = std::fmaf(X[6 + r], d2, std::fmaf(X[3 + r], d1, X[r] * d0));
      s = (s + g[r]) + X[9 + r];
      out[r] = std::fmaf(wj, s, out[r]);
    }
  }
}

// assemble H (compressed) and g = J^T r at state x; returns energies (weighted)
void assemble(const Problem& P, const double* targets, const float* tw, const orc_gn_params& prm,
              const double* x, System& S, double energy[3]) {
  std::fill(S.B.begin(), S.B.end(), 0.0);
  std::fill(S.Rot.begin(), S.Rot.end(), 0.0);
  std::fill(S.g.begin(), S.g.end(), 0.0);
  S.lambda = prm.lambda;
  double e_fit = 0, e_rot = 0, e_reg = 0;
  // ---- fit term
  auto fit_vertex = [&](int64_t i, double* SB, double* Sg, double& ef) {
    double om = tw ? (double)tw[i] : 1.0;
    if (!(om > 0)) return;
    double vt[3];
    deform_vertex(P, x, i, vt);
    double r[3] = {vt[0] - targets[3 * i], vt[1] - targets[3 * i + 1], vt[2] - targets[3 * i + 2]};
    double wf = prm.w_fit * om;
    ef += wf * (r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    double b[4][4];
    int js[4];
    for (int a = 0; a < 4; ++a) {
      int j = P.widx[4 * i + a];
      js[a] = j;
      if (j < 0) continue;
      double wj = P.w[4 * i + a];
      const double* g = &P.nodes[3 * (size_t)j];
      b[a][0] = wj * (P.pts[3 * i] - g[0]);
      b[a][1] = wj * (P.pts[3 * i + 1] - g[1]);
      b[a][2] = wj * (P.pts[3 * i + 2] - g[2]);
      b[a][3] = wj;
    }
    for (int a = 0; a < 4; ++a) {
      if (js[a] < 0) continue;
      for (int c = 0; c < 4; ++c)
        for (int rr = 0; rr < 3; ++rr) Sg[12 * (size_t)js[a] + 3 * c + rr] += wf * b[a][c] * r[rr];
      for (int bb = 0; bb < 4; ++bb) {
        if (js[bb] < 0) continue;
        int slot = S.find(js[a], js[bb]);
        double* B = &SB[16 * (size_t)slot];
        for (int c = 0; c < 4; ++c)
          for (int d = 0; d < 4; ++d) B[4 * c + d] += wf * b[a][c] * b[bb][d];
      }
    }
  };
#ifdef _OPENMP
  {  // courtesy multi-thread build: per-thread partial systems, summed afterwards
    const int T = omp_get_max_threads();
    std::vector<std::vector<double>> Bt((size_t)T), gt((size_t)T);
    std::vector<double> et((size_t)T, 0.0);
#pragma omp parallel
    {
      const int t = omp_get_thread_num();
      Bt[t].assign(S.B.size(), 0.0);
      gt[t].assign(S.g.size(), 0.0);
#pragma omp for schedule(static)
      for (int64_t i = 0; i < P.n; ++i) fit_vertex(i, Bt[t].data(), gt[t].data(), et[t]);
    }
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < (int64_t)S.B.size(); ++k) {
      double acc = 0;
      for (int t = 0; t < T; ++t) acc += Bt[t][k];
      S.B[k] += acc;
    }
    for (size_t k = 0; k < S.g.size(); ++k)
      for (int t = 0; t < T; ++t) S.g[k] += gt[t][k];
    for (int t = 0; t < T; ++t) e_fit += et[t];
  }
#else
  for (int64_t i = 0; i < P.n; ++i) fit_vertex(i, S.B.data(), S.g.data(), e_fit);
#endif
  // ---- rotation term (6 residuals on the columns of A_j)
  for (int64_t j = 0; j < P.m; ++j) {
    const double* X = x + 12 * (size_t)j;
    const double* a1 = X;
    const double* a2 = X + 3;
    const double* a3 = X + 6;
    double res[6], J[6][9];
    std::memset(J, 0, sizeof(J));
    auto dot = [](const double* u, const double* v) { return u[0] * v[0] + u[1] * v[1] + u[2] * v[2]; };
    res[0] = dot(a1, a2);
    res[1] = dot(a1, a3);
    res[2] = dot(a2, a3);
    res[3] = dot(a1, a1) - 1.0;
    res[4] = dot(a2, a2) - 1.0;
    res[5] = dot(a3, a3) - 1.0;
    for (int r = 0; r < 3; ++r) {
      J[0][r] = a2[r]; J[0][3 + r] = a1[r];
      J[1][r] = a3[r]; J[1][6 + r] = a1[r];
      J[2][3 + r] = a3[r]; J[2][6 + r] = a2[r];
      J[3][r] = 2 * a1[r];
      J[4][3 + r] = 2 * a2[r];
      J[5][6 + r] = 2 * a3[r];
    }
    for (int k = 0; k < 6; ++k) {
      e_rot += prm.w_rot * res[k] * res[k];
      for (int p = 0; p < 9; ++p) {
        S.g[12 * (size_t)j + p] += prm.w_rot * J[k][p] * res[k];
        for (int q = 0; q < 9; ++q) S.Rot[81 * (size_t)j + 9 * p + q] += prm.w_rot * J[k][p] * J[k][q];
      }
    }
  }
  // ---- regularisation ("smoothness") term over graph edges j -> l
  for (int64_t j = 0; j < P.m; ++j) {
    const double* X = x + 12 * (size_t)j;
    const double* gj = &P.nodes[3 * (size_t)j];
    for (int e = 0; e < P.gk1; ++e) {
      int l = P.graph[j * P.gk1 + e];
      if (l < 0 || l == j) continue;
      const double* gl = &P.nodes[3 * (size_t)l];
      const double* Xl = x + 12 * (size_t)l;
      double ev[3] = {gl[0] - gj[0], gl[1] - gj[1], gl[2] - gj[2]};
      double r[3];
      for (int rr = 0; rr < 3; ++rr)
        r[rr] = X[rr] * ev[0] + X[3 + rr] * ev[1] + X[6 + rr] * ev[2] + gj[rr] + X[9 + rr] -
                (gl[rr] + Xl[9 + rr]);
      e_reg += prm.w_reg * (r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
      double a[4] = {ev[0], ev[1], ev[2], 1.0};
      double b[4] = {0, 0, 0, -1.0};
      int sjj = S.find((int)j, (int)j), sjl = S.find((int)j, l), slj = S.find(l, (int)j),
          sll = S.find(l, l);
      for (int c = 0; c < 4; ++c)
        for (int d = 0; d < 4; ++d) {
          S.B[16 * (size_t)sjj + 4 * c + d] += prm.w_reg * a[c] * a[d];
          S.B[16 * (size_t)sjl + 4 * c + d] += prm.w_reg * a[c] * b[d];
          S.B[16 * (size_t)slj + 4 * c + d] += prm.w_reg * b[c] * a[d];
          S.B[16 * (size_t)sll + 4 * c + d] += prm.w_reg * b[c] * b[d];
        }
      for (int c = 0; c < 4; ++c)
        for (int rr = 0; rr < 3; ++rr) {
          S.g[12 * (size_t)j + 3 * c + rr] += prm.w_reg * a[c] * r[rr];
          S.g[12 * (size_t)l + 3 * c + rr] += prm.w_reg * b[c] * r[rr];
        }
    }
  }
  if (energy) {
    energy[0] = e_fit;
    energy[1] = e_rot;
    energy[2] = e_reg;
  }
}

// y = H v
void spmv(const System& S, const double* v, double* y) {
#pragma omp parallel for schedule(static)
  for (int64_t j = 0; j < S.m; ++j) {
    double acc[12];
    for (int k = 0; k < 12; ++k) acc[k] = S.lambda * v[12 * j + k];
    const double* R = &S.Rot[81 * (size_t)j];
    for (int p = 0; p < 9; ++p) {
      double s = 0;
      for (int q = 0; q < 9; ++q) s += R[9 * p + q] * v[12 * j + q];
      acc[p] += s;
    }
    for (int e = S.row_ptr[j]; e < S.row_ptr[j + 1]; ++e) {
      int l = S.col[e];
      const double* B = &S.B[16 * (size_t)e];
      const double* vl = v + 12 * (size_t)l;
      for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b)
          for (int r = 0; r < 3; ++r) acc[3 * a + r] += B[4 * a + b] * vl[3 * b + r];
    }
    for (int k = 0; k < 12; ++k) y[12 * j + k] = acc[k];
  }
}

void diag_block(const System& S, int64_t j, double D[144]) {
  std::memset(D, 0, sizeof(double) * 144);
  int slot = S.find((int)j, (int)j);
  const double* B = &S.B[16 * (size_t)slot];
  for (int a = 0; a < 4; ++a)
    for (int b = 0; b < 4; ++b)
      for (int r = 0; r < 3; ++r) D[(3 * a + r) * 12 + 3 * b + r] += B[4 * a + b];
  for (int p = 0; p < 9; ++p)
    for (int q = 0; q < 9; ++q) D[p * 12 + q] += S.Rot[81 * (size_t)j + 9 * p + q];
  for (int k = 0; k < 12; ++k) D[k * 12 + k] += S.lambda;
}

// block-Jacobi PCG on H delta = -g, fixed iteration count (+ optional relative-residual stop)
void pcg(const System& S, int iters, double tol, std::vector<double>& delta) {
  size_t N = (size_t)S.m * 12;
  std::vector<double> Minv((size_t)S.m * 144);
#pragma omp parallel for schedule(static)
  for (int64_t j = 0; j < S.m; ++j) {
    double D[144];
    diag_block(S, j, D);
    if (!spd_inverse(D, 12, &Minv[144 * (size_t)j])) {
      // fall back to the plain diagonal
      std::memset(&Minv[144 * (size_t)j], 0, sizeof(double) * 144);
      for (int k = 0; k < 12; ++k) Minv[144 * (size_t)j + 13 * k] = 1.0 / D[13 * k];
    }
  }
  auto precond = [&](const std::vector<double>& r, std::vector<double>& z) {
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < S.m; ++j)
      for (int a = 0; a < 12; ++a) {
        double s = 0;
        for (int b = 0; b < 12; ++b) s += Minv[144 * (size_t)j + 12 * a + b] * r[12 * j + b];
        z[12 * j + a] = s;
      }
  };
  delta.assign(N, 0.0);
  std::vector<double> r(N), z(N), p(N), q(N);
  for (size_t i = 0; i < N; ++i) r[i] = -S.g[i];
  double bnorm2 = 0;
  for (size_t i = 0; i < N; ++i) bnorm2 += r[i] * r[i];
  precond(r, z);
  p = z;
  double rz = 0;
  for (size_t i = 0; i < N; ++i) rz += r[i] * z[i];
  for (int it = 0; it < iters; ++it) {
    spmv(S, p.data(), q.data());
    double pq = 0;
    for (size_t i = 0; i < N; ++i) pq += p[i] * q[i];
    double alpha = (pq > 0 && rz > 0) ? rz / pq : 0.0;
    double rn2 = 0;
    for (size_t i = 0; i < N; ++i) {
      delta[i] += alpha * p[i];
      r[i] -= alpha * q[i];
      rn2 += r[i] * r[i];
    }
    if (tol > 0 && rn2 <= tol * tol * bnorm2) break;
    precond(r, z);
    double rz_new = 0;
    for (size_t i = 0; i < N; ++i) rz_new += r[i] * z[i];
    double beta = rz > 0 ? rz_new / rz : 0.0;
    for (size_t i = 0; i < N; ++i) p[i] = z[i] + beta * p[i];
    rz = rz_new;
  }
}

// 0: bucketed strict-weak ordering of the node-correspondence candidates (the contract of the build);
// 1: the reference's literal, non-transitive priority-queue comparator (measurement only)
int g_literal_comparator = 0;

Problem make_problem(const float* pts, int64_t n, const int32_t* node_ids, int64_t m,
                     const int32_t* widx, const float* w, const int32_t* graph, int gk1) {
  Problem P;
  P.pts = pts;
  P.n = n;
  P.m = m;
  P.widx = widx;
  P.w = w;
  P.graph = graph;
  P.gk1 = gk1;
  P.nodes.resize((size_t)m * 3);
  for (int64_t j = 0; j < m; ++j)
    for (int a = 0; a < 3; ++a) P.nodes[3 * j + a] = pts[3 * (size_t)node_ids[j] + a];
  return P;
}

}  // namespace

extern "C" {

// Deformation/Deformation.cpp:63-106: scan vertices in index order; an unremoved vertex becomes a
// node and removes its `suppress_k` nearest neighbours (other than itself) from later selection.
int64_t orc_uniform_sampling(const float* pts, int64_t n, int suppress_k, int32_t* out_ids) {
  std::vector<int32_t> nn((size_t)n * suppress_k);
  orc_knn_kdtree(pts, n, pts, n, suppress_k, nn.data(), nullptr);
  std::vector<char> removed((size_t)n, 0);
  int64_t cnt = 0;
  for (int64_t i = 0; i < n; ++i) {
    if (removed[i]) continue;
    out_ids[cnt++] = (int32_t)i;
    for (int j = 0; j < suppress_k; ++j) {
      int32_t id = nn[i * suppress_k + j];
      if (id >= 0 && id != i) removed[id] = 1;
    }
  }
  return cnt;
}

// Farthest-point sampling to a prescribed node count (C3/C5 prescribe m):
// float32 squared distances, argmax ties -> lowest index.  Result returned in selection order.
int64_t orc_sample_fps(const float* pts, int64_t n, int64_t n_nodes, int64_t first,
                       int32_t* out_ids) {
  if (n_nodes > n) n_nodes = n;
  std::vector<float> mind((size_t)n, std::numeric_limits<float>::infinity());
  int64_t cur = first;
  for (int64_t s = 0; s < n_nodes; ++s) {
    out_ids[s] = (int32_t)cur;
    float best = -1.0f;
    int64_t bi = 0;
    for (int64_t i = 0; i < n; ++i) {
      float d = dist2_f32(pts + 3 * i, pts + 3 * cur);
      if (d < mind[i]) mind[i] = d;
      if (mind[i] > best) {
        best = mind[i];
        bi = i;
      }
    }
    cur = bi;
  }
  return n_nodes;
}

// Deformation/Deformation.cpp:108-153 KNearestNeighbor(K): (K+1) nearest nodes including self.
void orc_node_graph(const float* nodes, int64_t m, int K, int32_t* out_idx) {
  orc_knn_kdtree(nodes, m, nodes, m, K + 1, out_idx, nullptr);
}

// SURVEY §8(a) D7 weights: k nearest nodes, d_max = distance to the (k+1)-th,
// w = (1 - d/d_max)^2 normalised to 1; all float32 in a fixed order.
void orc_node_weights(const float* pts, int64_t n, const float* nodes, int64_t m, int k,
                      int32_t* out_idx, float* out_w) {
  std::vector<int32_t> idx((size_t)n * (k + 1));
  std::vector<float> d2((size_t)n * (k + 1));
  orc_knn_kdtree(pts, n, nodes, m, k + 1, idx.data(), d2.data());
  for (int64_t i = 0; i < n; ++i) {
    const int32_t* id = &idx[i * (k + 1)];
    const float* dd = &d2[i * (k + 1)];
    float w[8];
    int kk = 0;
    for (int a = 0; a < k; ++a)
      if (id[a] >= 0) kk = a + 1;
    float dmax = (id[k] >= 0) ? std::sqrt(dd[k]) : 0.0f;
    float sum = 0.0f;
    for (int a = 0; a < k; ++a) {
      if (a < kk && dmax > 0.0f) {
        float r = std::sqrt(dd[a]) / dmax;
        float u = 1.0f - r;
        w[a] = u * u;
      } else {
        w[a] = 0.0f;
      }
      sum = sum + w[a];
    }
    for (int a = 0; a < k; ++a) {
      out_idx[i * k + a] = id[a];
      if (sum > 0.0f)
        out_w[i * k + a] = w[a] / sum;
      else
        out_w[i * k + a] = (a < kk) ? 1.0f / (float)kk : 0.0f;
    }
  }
}

void orc_set_literal_comparator(int on) { g_literal_comparator = on; }

// Deformation/Deformation.cpp:155-230, the Deform(tpts, tnormals, threshold) variant: the 6-D tree
// is built with scale = 0 (:162-170), i.e. it is a plain position search; per node (:186-203): 1-NN
// in the target (float32 squared distance, exact here where FLANN approximates), skipped when
// dists[0] > threshold (float against double, :198) or when cos(dir, normal) < 0.5 with
// dir = p1 - p in double (:202-203; a NaN cosine compares false and the node is kept, as there).
void orc_node_targets_threshold(const float* nodes, const float* node_normals, int64_t m,
                                const float* tpts, int64_t nt, double threshold, double* controls,
                                int32_t* valid) {
  std::vector<int32_t> nn1((size_t)m);
  std::vector<float> d1((size_t)m);
  orc_knn_kdtree(nodes, m, tpts, nt, 1, nn1.data(), d1.data());
  for (int64_t i = 0; i < m; ++i) {
    valid[i] = 0;
    for (int a = 0; a < 3; ++a) controls[3 * i + a] = (double)nodes[3 * i + a];
    if ((double)d1[i] > threshold) continue;
    const float* p1 = tpts + 3 * (size_t)nn1[i];
    double dir[3], nr[3];
    for (int a = 0; a < 3; ++a) {
      dir[a] = (double)p1[a] - (double)nodes[3 * i + a];
      nr[a] = (double)node_normals[3 * i + a];
    }
    double dot = dir[0] * nr[0] + dir[1] * nr[1] + dir[2] * nr[2];
    double dl = std::sqrt(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
    double nl = std::sqrt(nr[0] * nr[0] + nr[1] * nr[1] + nr[2] * nr[2]);
    if (dot / (dl * nl) < 0.5) continue;
    valid[i] = 1;
    for (int a = 0; a < 3; ++a) controls[3 * i + a] = (double)p1[a];
  }
}

// Deformation/Deformation.cpp:266-357 (correspondence per node) and :359-381 (2 Jacobi sweeps).
void orc_node_targets(const float* nodes, const float* node_normals, int64_t m, const float* tpts,
                      const float* tnormals, int64_t nt, double proj_len_err, double proj_dist_err,
                      const int32_t* graph, int gk1, int smooth_sweeps, double* controls,
                      int32_t* valid) {
  const int maxResult = 10000;  // :244
  std::vector<double> orig((size_t)m * 3);
  for (int64_t i = 0; i < m * 3; ++i) orig[i] = controls[i] = (double)nodes[i];
  // brute-force radius search is O(m * nt); the kd-tree gives the 1-NN, candidates by scan.
  std::vector<int32_t> nn1((size_t)m);
  std::vector<float> d1((size_t)m);
  orc_knn_kdtree(nodes, m, tpts, nt, 1, nn1.data(), d1.data());
  struct Cand {
    float d2;
    int32_t idx;
  };
  std::vector<Cand> cands;
  for (int64_t i = 0; i < m; ++i) {
    valid[i] = 0;
    const float* q = nodes + 3 * i;
    float minDist = d1[i];
    float radius = minDist * 2.0f;  // :288 (squared units)
    cands.clear();
    for (int64_t j = 0; j < nt; ++j) {
      float d = dist2_f32(tpts + 3 * j, q);
      if (d <= radius) cands.push_back({d, (int32_t)j});
    }
    std::sort(cands.begin(), cands.end(), [](const Cand& a, const Cand& b) {
      return a.d2 < b.d2 || (a.d2 == b.d2 && a.idx < b.idx);
    });
    // :286-297,315: when all maxResult slots fill, no -1 sentinel is found, newSize stays 0 and
    // the node is skipped.  Replicated.
    if ((int)cands.size() >= maxResult) continue;
    double nrm[3] = {node_normals[3 * i], node_normals[3 * i + 1], node_normals[3 * i + 2]};
    double nlen = std::sqrt(nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2]);
    struct Item {
      double key;  // floor(projDist / 1e-6): strict-weak version of the 1e-6 tolerance (:321)
      double aplen;
      int32_t idx;
      double plen, pdist;
    };
    std::vector<Item> items;
    for (const Cand& c : cands) {
      const float* tn = tnormals + 3 * (size_t)c.idx;
      double dn = nrm[0] * (double)tn[0] + nrm[1] * (double)tn[1] + nrm[2] * (double)tn[2];
      if (!(dn > 0)) continue;  // :306-312
      const float* tp = tpts + 3 * (size_t)c.idx;
      double dir[3] = {(double)tp[0] - orig[3 * i], (double)tp[1] - orig[3 * i + 1],
                       (double)tp[2] - orig[3 * i + 2]};
      double plen = (dir[0] * nrm[0] + dir[1] * nrm[1] + dir[2] * nrm[2]) / nlen;
      double sq = (dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]) - plen * plen;
      double pdist = sq > 0 ? std::sqrt(sq) : 0.0;
      items.push_back({std::floor(pdist / 1e-6), std::fabs(plen), c.idx, plen, pdist});
    }
    if (items.empty()) continue;  // :315
    if (g_literal_comparator) {
      // the reference's literal ordering (:318-327): a std::priority_queue over `operator<` with the
      // pairwise |projDist - other| <= 1e-6 tolerance.  That relation is not transitive, so the pop
      // order depends on the heap algorithm and the push order (here: libstdc++'s heap, candidates
      // pushed in ascending (d2, index) order as a radius search returns them) — it cannot be pinned
      // across standard libraries, which is why the build and the default oracle use the bucketed
      // strict-weak version above.  Kept to MEASURE how often the two pick different top-8 sets
      // (tests/test_oracle_deform.py).
      struct QNode {
        double projLen, projDist;
        int idx;
        bool operator<(const QNode& o) const {
          if (std::fabs(projDist - o.projDist) <= 1e-6) return std::fabs(projLen) > std::fabs(o.projLen);
          return projDist > o.projDist;
        }
      };
      std::priority_queue<QNode> Q;
      for (const Item& it : items) Q.push(QNode{it.plen, it.pdist, it.idx});
      std::vector<Item> ordered;
      int take = std::min<int>(8, (int)Q.size());
      for (int j = 0; j < take; ++j) {
        QNode c = Q.top();
        Q.pop();
        ordered.push_back({0.0, std::fabs(c.projLen), c.idx, c.projLen, c.projDist});
      }
      items.swap(ordered);
    } else {
      std::sort(items.begin(), items.end(), [](const Item& a, const Item& b) {
        if (a.key != b.key) return a.key < b.key;
        if (a.aplen != b.aplen) return a.aplen < b.aplen;
        return a.idx < b.idx;
      });
    }
    int nb = std::min<int>(8, (int)items.size());  // :338
    double mlen = 0, mdist = 0, mp[3] = {0, 0, 0};
    for (int j = 0; j < nb; ++j) {
      mlen += items[j].plen;
      mdist += items[j].pdist;
      const float* tp = tpts + 3 * (size_t)items[j].idx;
      for (int a = 0; a < 3; ++a) mp[a] = mp[a] + (double)tp[a];
    }
    mlen /= nb;
    mdist /= nb;
    for (int a = 0; a < 3; ++a) mp[a] /= nb;
    if (mlen >= proj_len_err || mdist >= proj_dist_err) continue;  // :350
    double dir[3] = {mp[0] - orig[3 * i], mp[1] - orig[3 * i + 1], mp[2] - orig[3 * i + 2]};
    double dl = std::sqrt(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
    double cosv = std::fabs((dir[0] * nrm[0] + dir[1] * nrm[1] + dir[2] * nrm[2]) / (dl * nlen));
    if (cosv < 0.1) continue;  // :353 (NaN compares false, as in the reference)
    valid[i] = 1;
    for (int a = 0; a < 3; ++a) controls[3 * i + a] = mp[a];
  }
  // :359-381: displacement <- mean of the (K+1)-neighbourhood displacements, `smooth_sweeps` times
  std::vector<double> tmp((size_t)m * 3);
  double wgt = 1.0 / (double)gk1;
  for (int it = 0; it < smooth_sweeps; ++it) {
    for (int64_t i = 0; i < m; ++i) {
      double acc[3] = {0, 0, 0};
      for (int e = 0; e < gk1; ++e) {
        int l = graph[i * gk1 + e];
        if (l < 0) continue;
        for (int a = 0; a < 3; ++a) acc[a] = acc[a] + wgt * (controls[3 * l + a] - orig[3 * l + a]);
      }
      for (int a = 0; a < 3; ++a) tmp[3 * i + a] = orig[3 * i + a] + acc[a];
    }
    std::memcpy(controls, tmp.data(), sizeof(double) * 3 * m);
  }
}

void orc_deform_vertices(const float* pts, int64_t n, const float* nodes, int64_t m,
                         const int32_t* widx, const float* w, const double* x, double* out_pts) {
  Problem P;
  P.pts = pts;
  P.n = n;
  P.m = m;
  P.widx = widx;
  P.w = w;
  P.graph = nullptr;
  P.gk1 = 0;
  P.nodes.resize((size_t)m * 3);
  for (int64_t i = 0; i < m * 3; ++i) P.nodes[i] = nodes[i];
  for (int64_t i = 0; i < n; ++i) deform_vertex(P, x, i, out_pts + 3 * i);
}

void orc_gn_assemble_dense(const float* pts, int64_t n, const int32_t* node_ids, int64_t m,
                           const int32_t* widx, const float* w, const int32_t* graph, int gk1,
                           const double* targets, const float* tweights, const orc_gn_params* prm,
                           const double* x, double* H, double* g) {
  Problem P = make_problem(pts, n, node_ids, m, widx, w, graph, gk1);
  System S;
  build_pattern(n, m, widx, graph, gk1, S);
  assemble(P, targets, tweights, *prm, x, S, nullptr);
  size_t N = (size_t)m * 12;
  std::memset(H, 0, sizeof(double) * N * N);
  for (int64_t j = 0; j < m; ++j) {
    for (int e = S.row_ptr[j]; e < S.row_ptr[j + 1]; ++e) {
      int l = S.col[e];
      for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b)
          for (int r = 0; r < 3; ++r)
            H[(12 * j + 3 * a + r) * N + 12 * (size_t)l + 3 * b + r] += S.B[16 * (size_t)e + 4 * a + b];
    }
    for (int p = 0; p < 9; ++p)
      for (int q = 0; q < 9; ++q) H[(12 * j + p) * N + 12 * j + q] += S.Rot[81 * (size_t)j + 9 * p + q];
    for (int k = 0; k < 12; ++k) H[(12 * j + k) * N + 12 * j + k] += S.lambda;
  }
  std::memcpy(g, S.g.data(), sizeof(double) * N);
}

void orc_gn_solve(const float* pts, const float* normals, int64_t n, const int32_t* node_ids,
                  int64_t m, const int32_t* widx, const float* w, const int32_t* graph, int gk1,
                  const double* targets, const float* tweights, const float* tpts,
                  const float* tnormals, int64_t nt, const orc_gn_params* prm, double* x,
                  double* out_pts, double energy[3]) {
  Problem P = make_problem(pts, n, node_ids, m, widx, w, graph, gk1);
  System S;
  build_pattern(n, m, widx, graph, gk1, S);
  std::vector<double> tg;
  std::vector<float> tw;
  std::vector<float> qf;
  std::vector<int32_t> nn;
  std::vector<float> nd2;
  std::vector<double> delta;
  void* tree = (prm->corr_mode == 1) ? orc_kdtree_build(tpts, nt) : nullptr;  // once per solve
  for (int it = 0; it < prm->gn_iters; ++it) {
    const double* T = targets;
    const float* W = tweights;
    if (prm->corr_mode == 1) {
      // re-derive correspondences: exact 1-NN of the current (float32) vertex positions
      tg.resize((size_t)n * 3);
      tw.resize((size_t)n);
      qf.resize((size_t)n * 3);
      nn.resize((size_t)n);
      nd2.resize((size_t)n);
#pragma omp parallel for schedule(static)
      for (int64_t i = 0; i < n; ++i) {
        deform_vertex_f32(P, x, i, &qf[3 * i]);
      }
      orc_kdtree_query(tree, qf.data(), n, 1, nn.data(), nd2.data());
#pragma omp parallel for schedule(static)
      for (int64_t i = 0; i < n; ++i) {
        int32_t c = nn[i];
        float om = 1.0f;
        if (prm->max_corr_dist2 > 0 && (double)nd2[i] > prm->max_corr_dist2) om = 0.0f;
        if (prm->min_corr_cos > -1.0 && normals && tnormals) {
          double dn = (double)normals[3 * i] * tnormals[3 * (size_t)c] +
                      (double)normals[3 * i + 1] * tnormals[3 * (size_t)c + 1] +
                      (double)normals[3 * i + 2] * tnormals[3 * (size_t)c + 2];
          if (dn < prm->min_corr_cos) om = 0.0f;
        }
        tw[i] = om;
        for (int a = 0; a < 3; ++a) tg[3 * i + a] = (double)tpts[3 * (size_t)c + a];
      }
      T = tg.data();
      W = tw.data();
    }
    assemble(P, T, W, *prm, x, S, energy);
    pcg(S, prm->pcg_iters, prm->pcg_tol, delta);
    for (size_t i = 0; i < (size_t)m * 12; ++i) x[i] += delta[i];
  }
  if (tree) orc_kdtree_free(tree);
  if (out_pts)
    for (int64_t i = 0; i < n; ++i) deform_vertex(P, x, i, out_pts + 3 * i);
}

// Correspondences of one Gauss-Newton iteration at node state x (m x 12): the float32 query point
// of every vertex (deform_vertex_f32) and its exact 1-NN in the target (lowest index on ties).
void orc_gn_correspondences(const float* pts, int64_t n, const int32_t* node_ids, int64_t m,
                            const int32_t* widx, const float* w, const double* x, const float* tpts,
                            int64_t nt, float* query_xyz, int32_t* out_idx, float* out_d2) {
  Problem P = make_problem(pts, n, node_ids, m, widx, w, nullptr, 0);
  std::vector<float> qf((size_t)n * 3);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) deform_vertex_f32(P, x, i, &qf[3 * i]);
  if (query_xyz) std::memcpy(query_xyz, qf.data(), sizeof(float) * 3 * n);
  if (out_idx && out_d2) {
    void* tree = orc_kdtree_build(tpts, nt);
    orc_kdtree_query(tree, qf.data(), n, 1, out_idx, out_d2);
    orc_kdtree_free(tree);
  }
}

// Deformation/Deformation.h:111-128 FacetNormal (normalised cross((p2-p1),(p3-p1))) and :86-109
// VertexNormal (sum of incident unit facet normals, normalised).
void orc_vertex_normals(const double* pts, int64_t n, const int32_t* facets, int64_t nf,
                        double* normals) {
  std::memset(normals, 0, sizeof(double) * 3 * n);
  for (int64_t f = 0; f < nf; ++f) {
    const double* p1 = pts + 3 * (size_t)facets[3 * f];
    const double* p2 = pts + 3 * (size_t)facets[3 * f + 1];
    const double* p3 = pts + 3 * (size_t)facets[3 * f + 2];
    double u[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
    double v[3] = {p3[0] - p1[0], p3[1] - p1[1], p3[2] - p1[2]};
    double c[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
    double l = std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
    if (!(l > 0)) continue;
    for (int k = 0; k < 3; ++k)
      for (int a = 0; a < 3; ++a) normals[3 * (size_t)facets[3 * f + k] + a] += c[a] / l;
  }
  for (int64_t i = 0; i < n; ++i) {
    double* q = normals + 3 * i;
    double l = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
    if (l > 0)
      for (int a = 0; a < 3; ++a) q[a] /= l;
  }
}

}  // extern "C"
