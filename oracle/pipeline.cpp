// oracle/pipeline.cpp — TEST INFRASTRUCTURE ONLY (see oracle.h).
// CPU restatement of the callers / data formats either side of the solver kernels:
//   Processor::RemoveOutliers            Processor/Processor.cpp:177-269
//   Processor::CheckConsistencyCore      Processor/Processor.cpp:66-115
//   Depth2Model::SaveModel               Depth2Model/Depth2Model.cpp:7-82
#include <cmath>
#include <cstring>
#include <vector>

#include "oracle.h"

extern "C" {

double orc_srt_inliers(const double* p, const double* q, int64_t n, const orc_camera* c1,
                       const orc_camera* c2, double scale, const double R[9], const double t[3],
                       double pixel_thr, uint8_t* flags) {
  double sR[9], iR[9];
  double inv_s = 1.0 / scale;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      sR[3 * i + j] = scale * R[3 * i + j];
      iR[3 * i + j] = inv_s * R[3 * j + i];
    }
  double err_all = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    const double* p1 = p + 3 * i;
    const double* p2 = q + 3 * i;
    double tp[3], tq[3], d[3] = {p2[0] - t[0], p2[1] - t[1], p2[2] - t[2]};
    for (int a = 0; a < 3; ++a) {
      tp[a] = sR[3 * a] * p1[0] + sR[3 * a + 1] * p1[1] + sR[3 * a + 2] * p1[2] + t[a];
      tq[a] = iR[3 * a] * d[0] + iR[3 * a + 1] * d[1] + iR[3 * a + 2] * d[2];
    }
    int32_t u1, v1, u2, v2, u1_, v1_, u2_, v2_;
    orc_project(c2, tp, &u1, &v1);
    orc_project(c2, p2, &u2, &v2);
    orc_project(c1, tq, &u2_, &v2_);
    orc_project(c1, p1, &u1_, &v1_);
    double e1 = std::sqrt((double)(u1 - u2) * (u1 - u2) + (double)(v1 - v2) * (v1 - v2));
    double e2 = std::sqrt((double)(u1_ - u2_) * (u1_ - u2_) + (double)(v1_ - v2_) * (v1_ - v2_));
    err_all += (e1 + e2) * 0.5;
    flags[i] = (e1 <= pixel_thr && e2 <= pixel_thr) ? 1 : 0;
  }
  return err_all;
}

namespace {
struct SeededDraw {
  uint64_t seed;
  int round;
};
void seeded_triples(int64_t size, int iter_num, int32_t* out, void* user) {
  SeededDraw* d = static_cast<SeededDraw*>(user);
  orc_draw_triples(size, iter_num, d->seed + (uint64_t)d->round++, out);
}
}  // namespace

// the round loop with the hypothesis samples of every round supplied by `draw` (called once per round
// with the number of matches still in play) — the reference draws them from the C library's rand()
// stream, which only the reference itself can replay (tests/test_ref_parity.py hands it over)
void orc_remove_outliers_cb(const double* p_in, const double* q_in, int64_t n, const orc_camera* c1,
                            const orc_camera* c2, int rounds, int iter_num, double pixel_err,
                            double adapt_ratio, orc_triples_fn draw, void* user, int32_t* keep_idx,
                            int64_t* n_keep, double* err, double* inlier_ratio) {
  std::vector<double> p(p_in, p_in + 3 * n), q(q_in, q_in + 3 * n);
  std::vector<int32_t> idx((size_t)n);
  for (int64_t i = 0; i < n; ++i) idx[i] = (int32_t)i;
  int64_t size = n;
  double e = HUGE_VAL, ratio = 1.0;
  for (int k = 0; k < rounds; ++k) {
    if (size < 3) break;
    std::vector<int32_t> triples((size_t)iter_num * 3);
    draw(size, iter_num, triples.data(), user);
    double sc, R[9], t[3];
    int32_t best;
    orc_srt_ransac(p.data(), q.data(), size, c1, c2, triples.data(), iter_num, &sc, R, t, &best,
                   nullptr);
    std::vector<uint8_t> flags((size_t)size);
    double err_all = orc_srt_inliers(p.data(), q.data(), size, c1, c2, sc, R, t, pixel_err * ratio,
                                     flags.data());
    int64_t m = 0;
    for (int64_t i = 0; i < size; ++i) {
      if (!flags[i]) continue;
      for (int a = 0; a < 3; ++a) {
        p[3 * m + a] = p[3 * i + a];
        q[3 * m + a] = q[3 * i + a];
      }
      idx[m++] = idx[i];
    }
    ratio *= adapt_ratio;
    e = err_all / (double)size;
    size = m;
    if (m < 3) break;
  }
  std::memcpy(keep_idx, idx.data(), sizeof(int32_t) * (size_t)size);
  *n_keep = size;
  if (err) *err = e;
  if (inlier_ratio) *inlier_ratio = 0.0;
}

void orc_remove_outliers(const double* p_in, const double* q_in, int64_t n, const orc_camera* c1,
                         const orc_camera* c2, int rounds, int iter_num, double pixel_err,
                         double adapt_ratio, uint64_t seed, int32_t* keep_idx, int64_t* n_keep,
                         double* err, double* inlier_ratio) {
  SeededDraw d{seed, 0};
  orc_remove_outliers_cb(p_in, q_in, n, c1, c2, rounds, iter_num, pixel_err, adapt_ratio,
                         seeded_triples, &d, keep_idx, n_keep, err, inlier_ratio);
}

void orc_depth2model(const float* dsp, const orc_camera* cam, double min_dsp, double max_dsp,
                     double smooth_threshold, double* out_xyz, int32_t* tex_index, int32_t* facets,
                     int64_t* n_points, int64_t* n_facets) {
  int w = cam->w, h = cam->h;
  std::vector<int32_t> tab((size_t)w * h, 0);
  int64_t np = 0;
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      double d = (double)dsp[(size_t)y * w + x];
      if (d > 0) {
        if (d > max_dsp || d < min_dsp) continue;
        tab[(size_t)y * w + x] = (int32_t)++np;
        if (out_xyz) orc_unproject(cam, x, y, 1.0 / d, out_xyz + 3 * (np - 1));
        if (tex_index) tex_index[np - 1] = y * w + x;
      }
    }
  float threshold = (float)(smooth_threshold * (max_dsp - min_dsp) / 100);
  int64_t nf = 0;
  auto D = [&](int x, int y) { return (double)dsp[(size_t)y * w + x]; };
  auto T = [&](int x, int y) { return tab[(size_t)y * w + x]; };
  for (int y = 0; y < h - 1; ++y)
    for (int x = 0; x < w - 1; ++x) {
      if (T(x, y) == 0 || T(x + 1, y + 1) == 0) continue;
      if (T(x, y + 1) != 0 && std::fabs(D(x, y) - D(x, y + 1)) <= threshold &&
          std::fabs(D(x + 1, y + 1) - D(x, y + 1)) <= threshold &&
          std::fabs(D(x, y) - D(x + 1, y + 1)) <= threshold) {
        if (facets) {
          facets[3 * nf] = T(x, y) - 1;
          facets[3 * nf + 1] = T(x, y + 1) - 1;
          facets[3 * nf + 2] = T(x + 1, y + 1) - 1;
        }
        ++nf;
      }
      if (T(x + 1, y) != 0 && std::fabs(D(x, y) - D(x + 1, y)) <= threshold &&
          std::fabs(D(x + 1, y + 1) - D(x + 1, y)) <= threshold &&
          std::fabs(D(x + 1, y + 1) - D(x, y)) <= threshold) {
        if (facets) {
          facets[3 * nf] = T(x, y) - 1;
          facets[3 * nf + 1] = T(x + 1, y + 1) - 1;
          facets[3 * nf + 2] = T(x + 1, y) - 1;
        }
        ++nf;
      }
    }
  if (n_points) *n_points = np;
  if (n_facets) *n_facets = nf;
}

void orc_check_consistency(float* depth, const orc_camera* cam, const float* ref_depths,
                           const orc_camera* ref_cams, int nref, double min_dsp, double max_dsp,
                           double reproj_err) {
  int w = cam->w, h = cam->h;
  for (int j = 0; j < h; ++j)
    for (int i = 0; i < w; ++i) {
      float& out = depth[(size_t)j * w + i];
      double dp = (double)out;
      if (!(dp >= min_dsp && dp <= max_dsp)) {
        out = 0.f;
        continue;
      }
      double p3d[3];
      orc_unproject(cam, i, j, 1.0 / dp, p3d);
      for (int k = 0; k < nref; ++k) {
        int32_t u, v;
        orc_project(&ref_cams[k], p3d, &u, &v);
        if (!(u >= 0 && u < ref_cams[k].w && v >= 0 && v < ref_cams[k].h)) {
          out = 0.f;
          break;
        }
        double rd = (double)ref_depths[(size_t)k * w * h + (size_t)v * w + u];
        if (!(rd >= min_dsp && rd <= max_dsp)) {
          out = 0.f;
          break;
        }
        double q3d[3];
        orc_unproject(&ref_cams[k], u, v, 1.0 / rd, q3d);
        orc_project(cam, q3d, &u, &v);
        if (!(u >= 0 && u < w && v >= 0 && v < h)) {
          out = 0.f;
          break;
        }
        double e = std::sqrt((double)((i - u) * (i - u) + (j - v) * (j - v)));
        if (e > reproj_err) {
          out = 0.f;
          break;
        }
      }
    }
}

}  // extern "C"
