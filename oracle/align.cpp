// oracle/align.cpp — TEST INFRASTRUCTURE ONLY (see oracle.h).
//
// CPU restatement, double precision, single thread, of the reference's rigid+scale pre-alignment:
//   SetUtils/PointSetUtils.cpp:3-62      barycenter, bounding box, PCA pivots
//   SetUtils/UnionSetUtils.cpp:4-45      union-find (rank == component size)
//   Alignment/Alignment.cpp:11-76        Align
//   Alignment/Alignment.cpp:79-233       RemoveGround
//   Alignment/Alignment.cpp:235-314      InitAlignment
//   Alignment/Alignment.cpp:316-421      LocalAlignment
//   Alignment/Alignment.cpp:423-546      LocalAlignmentCore (+ Common/Utils.h:124-149 CalcRotation)
//   Alignment/Alignment.cpp:618-654      RetainConnectRegion
//   Alignment/Alignment.cpp:31-34, Processor/Processor.cpp:1025-1026,1177-1189   apply / invert SRT
//
// Decisions where the reference is unpinned or broken, identical in the
// CUDA path:
//  * PCA eigenvector signs are solver-dependent in the reference (Eigen SelfAdjointEigenSolver);
//    here every pivot is normalised so that its largest-magnitude component is positive, before the
//    reference's own dot-product flips (Alignment.cpp:255-256, :444) are applied.
//  * `numeric_limits<double>::min()` as the initial value of max-reductions (:100-101,181,281-282,
//    :501-502) is replaced by lowest() (quirk 10) — except tMax1/tMax2/maxDist in RemoveGround whose
//    operands are non-negative, where the value is irrelevant.
//  * LocalAlignmentCore erases "the first label found by unordered_map iteration that the other
//    side lacks" (:481-497): iteration order is implementation-defined; the smallest such label is
//    erased here.
//  * CalcRotation's acos argument is clamped to [-1,1] and parallel axes give R = I (the reference
//    yields NaN there); an empty limb subset leaves that limb untouched (the reference reads
//    uninitialised indices); an empty NECK set skips the neck translation (the reference divides by 0).
#include <cmath>
#include <cstring>
#include <limits>
#include <numeric>
#include <unordered_map>
#include <vector>

#include "linalg.h"
#include "oracle.h"

namespace {
using namespace orc;

struct Stats {
  V3 bary{{0, 0, 0}};
  V3 bmin{{0, 0, 0}}, bmax{{0, 0, 0}};
  M3 pivots;  // columns = axes by descending eigenvalue
  double evals[3] = {0, 0, 0};
};

inline double* M(M3& m) { return &m.a[0][0]; }
inline const double* M(const M3& m) { return &m.a[0][0]; }

V3 pt(const double* p, int64_t i) { return V3{{p[3 * i], p[3 * i + 1], p[3 * i + 2]}}; }

// PointSetUtils::SetInput + CalcPivots over the points with sel[i] != 0 (sel == nullptr: all)
Stats stats_of(const double* pts, int64_t n, const unsigned char* sel) {
  Stats s;
  int64_t cnt = 0;
  bool first = true;
  for (int64_t i = 0; i < n; ++i) {
    if (sel && !sel[i]) continue;
    V3 p = pt(pts, i);
    for (int a = 0; a < 3; ++a) s.bary.x[a] += p.x[a];
    if (first) {
      s.bmin = s.bmax = p;
      first = false;
    } else
      for (int a = 0; a < 3; ++a) {
        s.bmin.x[a] = std::min(s.bmin.x[a], p.x[a]);
        s.bmax.x[a] = std::max(s.bmax.x[a], p.x[a]);
      }
    ++cnt;
  }
  if (cnt == 0) {
    s.pivots = m3_identity();
    return s;
  }
  for (int a = 0; a < 3; ++a) s.bary.x[a] /= (double)cnt;
  M3 C = m3_zero();
  for (int64_t i = 0; i < n; ++i) {
    if (sel && !sel[i]) continue;
    V3 p = pt(pts, i);
    double d[3] = {p.x[0] - s.bary.x[0], p.x[1] - s.bary.x[1], p.x[2] - s.bary.x[2]};
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) M(C)[3 * a + b] += d[a] * d[b];
  }
  double denom = cnt > 1 ? (double)(cnt - 1) : 1.0;
  for (int k = 0; k < 9; ++k) M(C)[k] /= denom;
  double w[3];
  M3 Q;
  sym_eig3(C, w, Q);  // ascending
  for (int c = 0; c < 3; ++c) {
    int src = 2 - c;  // PointSetUtils.cpp:35: pivots.col(i) = vectors.col(2 - i)
    double v[3] = {M(Q)[0 * 3 + src], M(Q)[1 * 3 + src], M(Q)[2 * 3 + src]};
    double l = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    int big = 0;
    for (int a = 1; a < 3; ++a)
      if (std::fabs(v[a]) > std::fabs(v[big])) big = a;
    double sg = v[big] < 0 ? -1.0 : 1.0;
    for (int a = 0; a < 3; ++a) M(s.pivots)[3 * a + c] = sg * v[a] / l;
    s.evals[c] = w[src];
  }
  return s;
}

V3 col(const M3& A, int c) { return V3{{A.a[0][c], A.a[1][c], A.a[2][c]}}; }
void set_col(M3& A, int c, const V3& v) {
  A.a[0][c] = v.x[0];
  A.a[1][c] = v.x[1];
  A.a[2][c] = v.x[2];
}
V3 neg(const V3& a) { return V3{{-a.x[0], -a.x[1], -a.x[2]}}; }

struct Range {
  double tmin = std::numeric_limits<double>::max();
  double tmax = std::numeric_limits<double>::lowest();
  int64_t imin = -1, imax = -1;
};

// t = axis.(p - origin)/|axis|^2, strict comparisons: the first extremum wins (Alignment.cpp:506-511)
Range range_of(const double* pts, int64_t n, const unsigned char* sel, const V3& origin,
               const V3& axis) {
  Range r;
  double nn = v3_norm(axis) * v3_norm(axis);
  for (int64_t i = 0; i < n; ++i) {
    if (sel && !sel[i]) continue;
    V3 p = pt(pts, i);
    V3 d{{p.x[0] - origin.x[0], p.x[1] - origin.x[1], p.x[2] - origin.x[2]}};
    double t = v3_dot(axis, d) / nn;
    if (r.tmin > t) {
      r.tmin = t;
      r.imin = i;
    }
    if (r.tmax < t) {
      r.tmax = t;
      r.imax = i;
    }
  }
  return r;
}

// Common/Utils.h:124-149
M3 calc_rotation(const V3& before, const V3& after) {
  double lb = v3_norm(before), la = v3_norm(after);
  V3 b{{before.x[0] / lb, before.x[1] / lb, before.x[2] / lb}};
  V3 a{{after.x[0] / la, after.x[1] / la, after.x[2] / la}};
  double value = v3_dot(b, a);
  value = std::max(-1.0, std::min(1.0, value));
  double angle = std::acos(value);
  V3 u = v3_cross(b, a);
  double lu = v3_norm(u);
  if (!(lu > 0.0)) return m3_identity();
  for (int k = 0; k < 3; ++k) u.x[k] /= lu;
  double c = std::cos(angle), s = std::sin(angle);
  M3 R;
  M(R)[0] = c + u.x[0] * u.x[0] * (1 - c);
  M(R)[1] = u.x[0] * u.x[1] * (1 - c) - u.x[2] * s;
  M(R)[2] = u.x[1] * s + u.x[0] * u.x[2] * (1 - c);
  M(R)[3] = u.x[2] * s + u.x[0] * u.x[1] * (1 - c);
  M(R)[4] = c + u.x[1] * u.x[1] * (1 - c);
  M(R)[5] = -u.x[0] * s + u.x[1] * u.x[2] * (1 - c);
  M(R)[6] = -u.x[1] * s + u.x[0] * u.x[2] * (1 - c);
  M(R)[7] = u.x[0] * s + u.x[1] * u.x[2] * (1 - c);
  M(R)[8] = c + u.x[2] * u.x[2] * (1 - c);
  return R;
}

void apply_masked(double* pts, double* normals, int64_t n, const int32_t* labels, unsigned mask,
                  double scale, const M3& R, const V3& t) {
  for (int64_t i = 0; i < n; ++i) {
    if (labels) {
      int32_t l = labels[i];
      if (l < 0 || l > 31 || !((mask >> l) & 1u)) continue;
    }
    V3 p = m3_mulv(R, pt(pts, i));
    for (int a = 0; a < 3; ++a) pts[3 * i + a] = scale * p.x[a] + t.x[a];
    if (normals) {
      V3 q = m3_mulv(R, pt(normals, i));
      for (int a = 0; a < 3; ++a) normals[3 * i + a] = q.x[a];
    }
  }
}

std::vector<unsigned char> select(const int32_t* labels, int64_t n, unsigned mask) {
  std::vector<unsigned char> s((size_t)n, 0);
  for (int64_t i = 0; i < n; ++i) {
    int32_t l = labels[i];
    s[i] = (l >= 0 && l <= 31 && ((mask >> l) & 1u)) ? 1 : 0;
  }
  return s;
}

unsigned present(const int32_t* labels, int64_t n, unsigned mask) {
  unsigned out = 0;
  for (int64_t i = 0; i < n; ++i) {
    int32_t l = labels[i];
    if (l >= 0 && l <= 31 && ((mask >> l) & 1u)) out |= 1u << l;
  }
  return out;
}

// Alignment.cpp:423-546.  Returns false (identity) when a subset is empty.
bool local_core(const double* src, const int32_t* sl, int64_t ns, const double* tgt,
                const int32_t* tl, int64_t nt, unsigned sel_mask, int label, double* scale, M3* R,
                V3* translate) {
  std::vector<unsigned char> ssel = select(sl, ns, sel_mask), tsel = select(tl, nt, sel_mask);
  unsigned sp = present(sl, ns, sel_mask), tp = present(tl, nt, sel_mask);
  if (!sp || !tp) return false;
  Stats s1 = stats_of(src, ns, ssel.data()), s2 = stats_of(tgt, nt, tsel.data());
  M3 tp_ = s2.pivots;
  if (v3_dot(col(s1.pivots, 0), col(tp_, 0)) < 0) set_col(tp_, 0, neg(col(tp_, 0)));
  // label-set bookkeeping (:477-497)
  int cs = __builtin_popcount(sp), ct = __builtin_popcount(tp);
  unsigned scnt = sp, tcnt = tp;
  if (cs < ct) {
    unsigned miss = tcnt & ~scnt;
    if (miss) tcnt &= ~(1u << __builtin_ctz(miss));
  } else if (cs > ct) {
    unsigned miss = scnt & ~tcnt;
    if (miss) scnt &= ~(1u << __builtin_ctz(miss));
  }
  std::vector<unsigned char> s2sel = select(sl, ns, scnt), t2sel = select(tl, nt, tcnt);
  Range r1 = range_of(src, ns, s2sel.data(), s1.bary, col(s1.pivots, 0));
  Range r2 = range_of(tgt, nt, t2sel.data(), s2.bary, col(tp_, 0));
  if (r1.imin < 0 || r2.imin < 0) return false;
  double a0 = r1.tmin, a1 = r1.tmax, b0 = r2.tmin, b1 = r2.tmax;
  int64_t f1 = r1.imin, n1 = r1.imax, f2 = r2.imin, n2 = r2.imax;
  if (sl[n1] != label) {
    std::swap(a0, a1);
    std::swap(f1, n1);
  }
  if (tl[n2] != label) {
    std::swap(b0, b1);
    std::swap(f2, n2);
  }
  (void)f2;
  *scale = (b1 - b0) / (a1 - a0);
  *R = calc_rotation(col(s1.pivots, 0), col(tp_, 0));
  V3 X = pt(src, f1);  // src_[fidx1] + baryCenter1 == the original point
  V3 RX = m3_mulv(*R, X);
  for (int a = 0; a < 3; ++a) translate->x[a] = X.x[a] - (*scale) * RX.x[a];
  return true;
}

}  // namespace

extern "C" {

void orc_apply_srt(double* pts, double* normals, int64_t n, double scale, const double R[9],
                   const double t[3]) {
  M3 Rm;
  std::memcpy(M(Rm), R, sizeof(double) * 9);
  apply_masked(pts, normals, n, nullptr, 0, scale, Rm, V3{{t[0], t[1], t[2]}});
}

// Processor/Processor.cpp:1177-1189: p_k = (1/s) R^T (p - t); n_k = R^T n
void orc_apply_srt_inverse(double* pts, double* normals, int64_t n, double scale,
                           const double R[9], const double t[3]) {
  M3 Rm;
  std::memcpy(M(Rm), R, sizeof(double) * 9);
  M3 Rt = m3_transpose(Rm);
  for (int64_t i = 0; i < n; ++i) {
    V3 d{{pts[3 * i] - t[0], pts[3 * i + 1] - t[1], pts[3 * i + 2] - t[2]}};
    V3 p = m3_mulv(Rt, d);
    for (int a = 0; a < 3; ++a) pts[3 * i + a] = p.x[a] / scale;
    if (normals) {
      V3 q = m3_mulv(Rt, pt(normals, i));
      for (int a = 0; a < 3; ++a) normals[3 * i + a] = q.x[a];
    }
  }
}

void orc_pointset_stats(const double* pts, int64_t n, double bary[3], double bmin[3],
                        double bmax[3], double pivots[9], double evals[3]) {
  Stats s = stats_of(pts, n, nullptr);
  for (int a = 0; a < 3; ++a) {
    if (bary) bary[a] = s.bary.x[a];
    if (bmin) bmin[a] = s.bmin.x[a];
    if (bmax) bmax[a] = s.bmax.x[a];
    if (evals) evals[a] = s.evals[a];
  }
  if (pivots) std::memcpy(pivots, M(s.pivots), sizeof(double) * 9);
}

void orc_project_range(const double* pts, int64_t n, const int32_t* labels, int32_t label,
                       const double origin[3], const double axis[3], double* tmin, double* tmax,
                       int64_t* argmin, int64_t* argmax) {
  std::vector<unsigned char> sel;
  if (labels) {
    sel.resize((size_t)n);
    for (int64_t i = 0; i < n; ++i) sel[i] = labels[i] == label;
  }
  Range r = range_of(pts, n, labels ? sel.data() : nullptr, V3{{origin[0], origin[1], origin[2]}},
                     V3{{axis[0], axis[1], axis[2]}});
  if (tmin) *tmin = r.tmin;
  if (tmax) *tmax = r.tmax;
  if (argmin) *argmin = r.imin;
  if (argmax) *argmax = r.imax;
}

// Alignment.cpp:618-654 with SetUtils/UnionSetUtils.cpp
void orc_retain_connected(double* pts, double* normals, int64_t* n_io, int32_t* facets,
                          int64_t* nf_io) {
  int64_t n = *n_io, nf = *nf_io;
  if (n <= 0) return;
  std::vector<int32_t> par((size_t)n), rank((size_t)n, 1);
  std::iota(par.begin(), par.end(), 0);
  auto find = [&](int32_t x) {
    int32_t r = x;
    while (par[r] != r) r = par[r];
    int32_t i = x;
    while (i != r) {
      int32_t j = par[i];
      par[i] = r;
      i = j;
    }
    return r;
  };
  auto merge = [&](int32_t x, int32_t y) {
    int32_t a = find(x), b = find(y);
    if (a == b) return;
    if (rank[a] > rank[b]) {
      par[b] = a;
      rank[a] += rank[b];
    } else {
      par[a] = b;
      rank[b] += rank[a];
    }
  };
  for (int64_t f = 0; f < nf; ++f) {
    merge(facets[3 * f], facets[3 * f + 1]);
    merge(facets[3 * f], facets[3 * f + 2]);
  }
  int32_t maxRank = rank[0], r = 0;
  for (int64_t i = 1; i < n; ++i)
    if (maxRank < rank[i]) {
      maxRank = rank[i];
      r = (int32_t)i;
    }
  std::vector<int32_t> mp((size_t)n, -1);
  int64_t m = 0;
  for (int64_t i = 0; i < n; ++i) {
    if (find((int32_t)i) != r) continue;
    for (int a = 0; a < 3; ++a) {
      pts[3 * m + a] = pts[3 * i + a];
      if (normals) normals[3 * m + a] = normals[3 * i + a];
    }
    mp[i] = (int32_t)m++;
  }
  int64_t mf = 0;
  for (int64_t f = 0; f < nf; ++f) {
    if (find(facets[3 * f]) != r) continue;
    int32_t a = mp[facets[3 * f]], b = mp[facets[3 * f + 1]], c = mp[facets[3 * f + 2]];
    facets[3 * mf] = a;
    facets[3 * mf + 1] = b;
    facets[3 * mf + 2] = c;
    ++mf;
  }
  *n_io = m;
  *nf_io = mf;
}

// Alignment.cpp:79-233
void orc_remove_ground(double* pts, double* normals, int64_t* n_io, int32_t* facets,
                       int64_t* nf_io, double dist_thres, double ground_ratio,
                       double ground_ray[3]) {
  int64_t n = *n_io, nf = *nf_io;
  Stats s = stats_of(pts, n, nullptr);
  V3 pivot = col(s.pivots, 0);
  double nn = v3_norm(pivot) * v3_norm(pivot);
  std::vector<double> t((size_t)n);
  double tMax1 = std::numeric_limits<double>::min(), tMax2 = std::numeric_limits<double>::min();
  for (int64_t i = 0; i < n; ++i) {
    V3 p = pt(pts, i);
    V3 d{{p.x[0] - s.bary.x[0], p.x[1] - s.bary.x[1], p.x[2] - s.bary.x[2]}};
    t[i] = v3_dot(pivot, d) / nn;
    if (t[i] < 0)
      tMax1 = std::max(-t[i], tMax1);
    else
      tMax2 = std::max(t[i], tMax2);
  }
  std::vector<int64_t> idx1, idx2;
  for (int64_t i = 0; i < n; ++i) {
    if (t[i] < 0) {
      if (-t[i] > tMax1 * dist_thres) idx1.push_back(i);
    } else {
      if (t[i] > tMax2 * dist_thres) idx2.push_back(i);
    }
  }
  const std::vector<int64_t>& idx = idx1.size() > idx2.size() ? idx1 : idx2;
  V3 gr = idx1.size() > idx2.size() ? neg(pivot) : pivot;
  for (int a = 0; a < 3; ++a) ground_ray[a] = gr.x[a];
  // plane fit (:148-160)
  M3 A = m3_zero();
  V3 b{{0, 0, 0}};
  for (int64_t i : idx) {
    V3 c = pt(pts, i);
    for (int a = 0; a < 3; ++a) {
      for (int k = 0; k < 3; ++k) M(A)[3 * a + k] += c.x[a] * c.x[k];
      b.x[a] += c.x[a];
    }
  }
  V3 ans = m3_mulv(m3_inverse(A), b);
  ans = neg(ans);
  double d = 1.0 / v3_norm(ans);
  double la = v3_norm(ans);
  for (int a = 0; a < 3; ++a) ans.x[a] /= la;
  if (v3_dot(ans, pivot) < 0) {
    ans = neg(ans);
    d = -d;
  }
  double maxDist = std::numeric_limits<double>::min();
  std::vector<double> dist(idx.size());
  for (size_t k = 0; k < idx.size(); ++k) {
    dist[k] = std::fabs(v3_dot(ans, pt(pts, idx[k])) + d);
    maxDist = std::max(maxDist, dist[k]);
  }
  double threshold = maxDist * ground_ratio;
  std::vector<unsigned char> remove((size_t)n, 0);
  for (size_t k = 0; k < idx.size(); ++k)
    if (dist[k] < threshold) remove[idx[k]] = 1;
  std::vector<int32_t> mp((size_t)n, -1);
  int64_t m = 0;
  for (int64_t i = 0; i < n; ++i) {
    if (remove[i]) continue;
    for (int a = 0; a < 3; ++a) {
      pts[3 * m + a] = pts[3 * i + a];
      if (normals) normals[3 * m + a] = normals[3 * i + a];
    }
    mp[i] = (int32_t)m++;
  }
  int64_t mf = 0;
  for (int64_t f = 0; f < nf; ++f) {
    int32_t a = mp[facets[3 * f]], bb = mp[facets[3 * f + 1]], c = mp[facets[3 * f + 2]];
    if (a < 0 || bb < 0 || c < 0) continue;
    facets[3 * mf] = a;
    facets[3 * mf + 1] = bb;
    facets[3 * mf + 2] = c;
    ++mf;
  }
  *n_io = m;
  *nf_io = mf;
  orc_retain_connected(pts, normals, n_io, facets, nf_io);
}

// Alignment.cpp:235-314
void orc_init_alignment(const double* src, int64_t ns, const double* tgt, int64_t nt,
                        const double ground_ray[3], const double view_ray[3], double transform[16],
                        double* scale) {
  Stats s1 = stats_of(src, ns, nullptr), s2 = stats_of(tgt, nt, nullptr);
  M3 tp = s2.pivots;
  V3 gr{{ground_ray[0], ground_ray[1], ground_ray[2]}}, vr{{view_ray[0], view_ray[1], view_ray[2]}};
  if (v3_dot(gr, col(tp, 0)) < 0) set_col(tp, 0, neg(col(tp, 0)));
  if (v3_dot(vr, col(tp, 2)) < 0) set_col(tp, 2, neg(col(tp, 2)));
  Range r1 = range_of(src, ns, nullptr, s1.bary, col(s1.pivots, 0));
  Range r2 = range_of(tgt, nt, nullptr, s2.bary, col(tp, 0));
  double sc = (r2.tmax - r2.tmin) / (r1.tmax - r1.tmin);
  M3 R = m3_mul(tp, m3_inverse(s1.pivots));
  V3 Rb = m3_mulv(R, s1.bary);
  V3 c0 = col(tp, 0);
  double k = r2.tmax - r1.tmax * sc;
  for (int i = 0; i < 16; ++i) transform[i] = (i % 5 == 0) ? 1.0 : 0.0;
  for (int a = 0; a < 3; ++a) {
    for (int b = 0; b < 3; ++b) transform[4 * a + b] = M(R)[3 * a + b];
    transform[4 * a + 3] = c0.x[a] * k + s2.bary.x[a] - sc * Rb.x[a];
  }
  *scale = sc;
}

// Alignment.cpp:11-76
void orc_align(double* src, double* src_normals, int64_t ns, const int32_t* src_labels,
               double* tgt, double* tgt_normals, int64_t* nt, int32_t* tgt_facets, int64_t* ntf,
               const double view_ray[3], double dist_thres, double ground_ratio,
               int remove_ground, int local_align, int neck_label, int32_t* out_tgt_labels,
               double ground_ray[3]) {
  if (remove_ground)
    orc_remove_ground(tgt, tgt_normals, nt, tgt_facets, ntf, dist_thres, ground_ratio, ground_ray);
  double T[16], scale = 1.0;
  orc_init_alignment(src, ns, tgt, *nt, ground_ray, view_ray, T, &scale);
  M3 R;
  V3 tr;
  for (int a = 0; a < 3; ++a) {
    for (int b = 0; b < 3; ++b) M(R)[3 * a + b] = T[4 * a + b];
    tr.x[a] = T[4 * a + 3];
  }
  apply_masked(src, src_normals, ns, nullptr, 0, scale, R, tr);
  std::vector<int32_t> tl((size_t)*nt);
  orc_part_recog(src, ns, src_labels, tgt, *nt, tl.data());
  // neck-centroid translation (:54-64)
  V3 bc1{{0, 0, 0}}, bc2{{0, 0, 0}};
  int64_t c1 = 0, c2 = 0;
  for (int64_t i = 0; i < ns; ++i)
    if (src_labels[i] == neck_label) {
      ++c1;
      for (int a = 0; a < 3; ++a) bc1.x[a] += src[3 * i + a];
    }
  for (int64_t i = 0; i < *nt; ++i)
    if (tl[i] == neck_label) {
      ++c2;
      for (int a = 0; a < 3; ++a) bc2.x[a] += tgt[3 * i + a];
    }
  if (c1 > 0 && c2 > 0) {
    V3 off;
    for (int a = 0; a < 3; ++a) off.x[a] = bc2.x[a] / (double)c2 - bc1.x[a] / (double)c1;
    for (int64_t i = 0; i < ns; ++i)
      for (int a = 0; a < 3; ++a) src[3 * i + a] = src[3 * i + a] + off.x[a];
  }
  if (local_align) {
    // (selection mask, end label, apply mask) per limb — Alignment.cpp:326-420; PART enum values
    // PartRecognition/PartRecognition.h:13-30
    struct Limb { unsigned sel; int label; unsigned apply; };
    const Limb limbs[4] = {
        {(1u << 2) | (1u << 3) | (1u << 4), 4, (1u << 2) | (1u << 3) | (1u << 4)},
        {(1u << 5) | (1u << 6) | (1u << 7), 7, (1u << 5) | (1u << 6) | (1u << 7)},
        {(1u << 8) | (1u << 9), 9, (1u << 8) | (1u << 9) | (1u << 10)},
        {(1u << 11) | (1u << 12), 12, (1u << 11) | (1u << 12) | (1u << 13)}};
    // all four subsets are gathered from the same pre-limb geometry (:326-371); the label sets
    // are disjoint, so evaluating limb by limb on the progressively transformed src is identical
    for (const Limb& L : limbs) {
      double sc = 1.0;
      M3 Rl = m3_identity();
      V3 tl3{{0, 0, 0}};
      if (!local_core(src, src_labels, ns, tgt, tl.data(), *nt, L.sel, L.label, &sc, &Rl, &tl3))
        continue;
      apply_masked(src, src_normals, ns, src_labels, L.apply, sc, Rl, tl3);
    }
  }
  if (out_tgt_labels) std::memcpy(out_tgt_labels, tl.data(), sizeof(int32_t) * (size_t)*nt);
}

// Alignment.cpp:423-546 alone: the limb = the points whose label is in sel_mask; transform row-major
// 4x4 = [R | translate] as the reference returns it.  Returns 0 when a subset is empty.
int orc_local_alignment_core(const double* src, const int32_t* s_labels, int64_t ns,
                             const double* tgt, const int32_t* t_labels, int64_t nt,
                             unsigned sel_mask, int label, double transform[16], double* scale) {
  double sc = 1.0;
  M3 R = m3_identity();
  V3 t{{0, 0, 0}};
  bool ok = local_core(src, s_labels, ns, tgt, t_labels, nt, sel_mask, label, &sc, &R, &t);
  for (int i = 0; i < 16; ++i) transform[i] = (i % 5 == 0) ? 1.0 : 0.0;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) transform[4 * i + j] = M(R)[3 * i + j];
    transform[4 * i + 3] = t.x[i];
  }
  *scale = sc;
  return ok ? 1 : 0;
}

}  // extern "C"
