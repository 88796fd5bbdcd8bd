// mvs_shim.h — header-only C++ mirror of the reference's class interface over the C ABI
// (include/mvs_b200.h).  A `Processor`-style caller keeps its call sequences:
//
//   SRTSolver solver(200); solver.SetPrintFlag(false); solver.SetInput(matches, camI, camJ);
//   solver.EstimateTransformRansac(scale, R, t);                        [Processor.cpp:202-205]
//   Alignment align; align.Align(src, sn, sf, tgt, tn, tf, viewRay);    [Processor.cpp:1130-1131]
//   Deformation deform(src, sn, sf); deform.Deform(tgt, tn, 100.0, 100.0);
//   exportOBJ(path, &deform);                                           [Processor.cpp:1135-1137]
//   Model2Depth::SetInput(points_, normals_, facets_, cameras, imgdirs); Model2Depth::Run(argc, argv);
//                                                                       [Processor.cpp:1190-1191]
//
// Types: the reference uses Eigen::Vector3d / Matrix3d (Eigen is not available in this image).
// With MVS_SHIM_USE_EIGEN defined the shim uses them; otherwise layout-compatible PODs
// (3 contiguous doubles; row-major 3x3 with operator()(r,c)).
// Error convention: the reference prints to cerr and calls exit(-1); the shim throws mvs::Error
// (status + mvs_last_error()) instead.  There is no CPU fallback.
#ifndef MVS_SHIM_H
#define MVS_SHIM_H

#include <stdint.h>

#include <array>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/mvs_b200.h"

#ifdef MVS_SHIM_USE_EIGEN
#include <Eigen/Dense>
#endif

namespace mvs {

struct Error : std::runtime_error {
  int status;
  Error(int s, const char* msg) : std::runtime_error(std::string("mvs_b200: ") + msg), status(s) {}
};
inline void check(int status) {
  if (status != MVS_OK) throw Error(status, mvs_last_error());
}

#ifdef MVS_SHIM_USE_EIGEN
using Vector3d = Eigen::Vector3d;
struct Matrix3d : Eigen::Matrix<double, 3, 3, Eigen::RowMajor> {
  using Eigen::Matrix<double, 3, 3, Eigen::RowMajor>::Matrix;
};
#else
struct Vector3d {
  double v[3];
  Vector3d() : v{0, 0, 0} {}
  Vector3d(double x, double y, double z) : v{x, y, z} {}
  double& operator[](int i) { return v[i]; }
  const double& operator[](int i) const { return v[i]; }
  double* data() { return v; }
  const double* data() const { return v; }
};
struct Matrix3d {
  double m[9];
  Matrix3d() : m{1, 0, 0, 0, 1, 0, 0, 0, 1} {}
  double& operator()(int r, int c) { return m[3 * r + c]; }
  const double& operator()(int r, int c) const { return m[3 * r + c]; }
  double* data() { return m; }
  const double* data() const { return m; }
};
#endif
static_assert(sizeof(Vector3d) == 3 * sizeof(double), "Vector3d must be 3 packed doubles");

// one library context per process (the reference is single-threaded with global state)
inline mvs_ctx* context() {
  static mvs_ctx* ctx = [] {
    mvs_ctx* c = nullptr;
    check(mvs_ctx_create(&c, 0, nullptr));
    return c;
  }();
  return ctx;
}

// [ref: Camera/Camera.h:9-49]
class Camera {
 public:
  void SetK(const Matrix3d& K) { std::memcpy(c_.K, K.data(), sizeof(c_.K)); }
  void SetRT(const Matrix3d& R, const Vector3d& t) {
    std::memcpy(c_.R, R.data(), sizeof(c_.R));
    std::memcpy(c_.t, t.data(), sizeof(c_.t));
  }
  void SetW(int w) { c_.w = w; }
  void SetH(int h) { c_.h = h; }
  int W() const { return c_.w; }
  int H() const { return c_.h; }
  const mvs_camera& c() const { return c_; }

 private:
  mvs_camera c_{};
};

// [ref: Solver/SRTSolver.h:8-39]
class SRTSolver {
 public:
  explicit SRTSolver(int iter_num_ = 100) {
    check(mvs_srt_create(context(), &h_));
    check(mvs_srt_set_iteration_num(h_, iter_num_));
  }
  ~SRTSolver() { mvs_srt_destroy(h_); }
  SRTSolver(const SRTSolver&) = delete;
  SRTSolver& operator=(const SRTSolver&) = delete;

  void SetInput(const std::vector<std::pair<Vector3d, Vector3d>>& matches, const Camera& cam1,
                const Camera& cam2) {
    std::vector<double> p(3 * matches.size()), q(3 * matches.size());
    for (size_t i = 0; i < matches.size(); ++i)
      for (int a = 0; a < 3; ++a) {
        p[3 * i + a] = matches[i].first[a];
        q[3 * i + a] = matches[i].second[a];
      }
    check(mvs_srt_set_input(h_, p.data(), q.data(), (int64_t)matches.size(), &cam1.c(), &cam2.c()));
  }
  void SetIterationNum(int it) { check(mvs_srt_set_iteration_num(h_, it)); }
  void SetPrintFlag(bool on) { check(mvs_srt_set_print(h_, on ? 1 : 0)); }
  // [ref: SRTSolver.cpp:272-275] closed form
  void EstimateTransform(double& scale, Matrix3d& R, Vector3d& t) {
    check(mvs_srt_estimate(h_, &scale, R.data(), t.data()));
  }
  // [ref: SRTSolver.cpp:277-280]; the reference draws its 3-point samples from unseeded rand():
  // here they come from a counter-based RNG with `seed`, or are supplied (iter_num x 3 indices)
  void EstimateTransformRansac(double& scale, Matrix3d& R, Vector3d& t, uint64_t seed = 0,
                               const int32_t* triples = nullptr) {
    check(mvs_srt_estimate_ransac(h_, triples, seed, &scale, R.data(), t.data(), nullptr, nullptr));
  }
  // [ref: SRTSolver.cpp:256-270]
  void EstimateTransform(double& scale, double R[3][3], double t[3], uint64_t seed = 0) {
    check(mvs_srt_estimate_array(h_, nullptr, seed, &scale, R, t));
  }
  double ResidualError(double scale, const Matrix3d& R, const Vector3d& t) {
    double e = 0;
    check(mvs_srt_residual(h_, scale, R.data(), t.data(), &e));
    return e;
  }

 private:
  mvs_srt* h_ = nullptr;
};

// [ref: PartRecognition/PartRecognition.h:32-49]; LoadParts (file parsing) stays with the caller
class PartRecognition {
 public:
  void SetInput(const std::vector<Vector3d>& points_) { points = points_; }
  void SetLabels(const std::vector<int>& parts_) { parts = parts_; }
  std::vector<Vector3d>& GetPoints() { return points; }
  std::vector<int>& GetLabels() { return parts; }
  void PartRecog(const std::vector<Vector3d>& points_, std::vector<int>& outParts) {
    outParts.resize(points_.size());
    check(mvs_part_recog(context(), points.empty() ? nullptr : points[0].data(),
                         (int64_t)points.size(), parts.data(),
                         points_.empty() ? nullptr : points_[0].data(), (int64_t)points_.size(),
                         outParts.data()));
  }

 protected:
  std::vector<Vector3d> points;
  std::vector<int> parts;
};

// [ref: Alignment/Alignment.h:21-36]  (arguments are mutated in place, like the reference)
class Alignment {
 public:
  mvs_align_params params;
  std::vector<int> s_labels;  // per-template-vertex PART labels (Template/part/parts)
  std::vector<int> t_labels;  // transferred labels of the last Align
  Alignment() { check(mvs_align_default_params(&params)); }

  void Align(std::vector<Vector3d>& src, std::vector<Vector3d>& s_normals, std::vector<int>& s_facets,
             std::vector<Vector3d>& tgt, std::vector<Vector3d>& t_normals, std::vector<int>& t_facets,
             const Vector3d& viewRay) {
    (void)s_facets;
    int64_t nt = (int64_t)tgt.size(), ntf = (int64_t)t_facets.size() / 3;
    t_labels.assign(tgt.size(), 0);
    double ground[3] = {0, 0, 0};
    check(mvs_align(context(), src[0].data(), s_normals.empty() ? nullptr : s_normals[0].data(),
                    (int64_t)src.size(), s_labels.data(), tgt[0].data(),
                    t_normals.empty() ? nullptr : t_normals[0].data(), &nt, t_facets.data(), &ntf,
                    viewRay.data(), &params, t_labels.data(), ground));
    tgt.resize((size_t)nt);
    if (!t_normals.empty()) t_normals.resize((size_t)nt);
    t_facets.resize((size_t)ntf * 3);
    t_labels.resize((size_t)nt);
  }
  Vector3d RemoveGround(std::vector<Vector3d>& points, std::vector<Vector3d>& normals,
                        std::vector<int>& facets) {
    int64_t n = (int64_t)points.size(), nf = (int64_t)facets.size() / 3;
    Vector3d ground;
    check(mvs_remove_ground(context(), points[0].data(), normals.empty() ? nullptr : normals[0].data(),
                            &n, facets.data(), &nf, &params, ground.data()));
    points.resize((size_t)n);
    if (!normals.empty()) normals.resize((size_t)n);
    facets.resize((size_t)nf * 3);
    return ground;
  }
  void RetainConnectRegion(std::vector<Vector3d>& points, std::vector<Vector3d>& normals,
                           std::vector<int>& facets) {
    int64_t n = (int64_t)points.size(), nf = (int64_t)facets.size() / 3;
    bool hn = normals.size() == points.size() && !normals.empty();
    check(mvs_retain_connected(points[0].data(), hn ? normals[0].data() : nullptr, &n, facets.data(),
                               &nf));
    points.resize((size_t)n);
    if (hn) normals.resize((size_t)n);
    facets.resize((size_t)nf * 3);
  }
};

// [ref: Deformation/Deformation.h:224-252]
class Deformation {
 public:
  // Solver behind Deform(tpts, tnormals, projLenErr, projDistErr): the embedded-deformation
  // Gauss-Newton loop (north_star) or the as-rigid-as-possible solver the reference runs through
  // CGAL (Deformation.cpp:383-400; deform(5, 1e-4)).
  enum Solver { GAUSS_NEWTON, ARAP };
  Solver solver = GAUSS_NEWTON;
  mvs_gn_params gn;
  mvs_arap_params arap;
  Deformation(const std::vector<Vector3d>& points_, const std::vector<Vector3d>& normals_,
              const std::vector<int>& facets_)
      : n_(points_.size()) {
    check(mvs_gn_default_params(&gn));
    check(mvs_arap_default_params(&arap));
    check(mvs_deform_create(context(), points_[0].data(),
                            normals_.empty() ? nullptr : normals_[0].data(), (int64_t)points_.size(),
                            facets_.data(), (int64_t)facets_.size() / 3, &h_));
  }
  ~Deformation() { mvs_deform_destroy(h_); }
  Deformation(const Deformation&) = delete;
  Deformation& operator=(const Deformation&) = delete;

  void UniformSampling() { check(mvs_deform_uniform_sampling(h_, 16)); }  // Deformation.cpp:97
  void SetSamplingIndex(const std::vector<int>& ids) {
    check(mvs_deform_set_nodes(h_, ids.data(), (int64_t)ids.size()));
  }
  // [ref: Deformation.h:237-240]
  void Deform(const std::vector<Vector3d>& tpts, const std::vector<Vector3d>& tnormals,
              double threshold) {
    set_target(tpts, tnormals);
    check(mvs_deform_run_threshold(h_, threshold, &gn));
  }
  // [ref: Deformation.h:241-245]
  void Deform(const std::vector<Vector3d>& tpts, const std::vector<Vector3d>& tnormals,
              double projLenErr, double projDistErr) {
    set_target(tpts, tnormals);
    if (solver == ARAP)
      check(mvs_deform_run_arap(h_, projLenErr, projDistErr, &arap));
    else
      check(mvs_deform_run(h_, projLenErr, projDistErr, &gn));
  }
  // Device-side shortcut (not in the reference): the target is the set of valid pixels of an
  // inverse-depth image — Image3D::SolveUnProjectionD (Image3D.cpp:92-106) + "keep the valid points"
  // + Deform's target upload, without the host vectors.  Then run with DeformCurrentTarget().
  int64_t SetTargetDepth(const std::vector<float>& inv_depth, const Camera& cam, double minDsp,
                         double maxDsp) {
    int64_t nt = 0;
    check(mvs_deform_set_target_depth(h_, inv_depth.data(), &cam.c(), minDsp, maxDsp, &nt));
    return nt;
  }
  void DeformCurrentTarget(double projLenErr, double projDistErr) {
    if (solver == ARAP)
      check(mvs_deform_run_arap(h_, projLenErr, projDistErr, &arap));
    else
      check(mvs_deform_run(h_, projLenErr, projDistErr, &gn));
  }
  // replace CGAL's Polyhedron() accessor (Deformation.h:233)
  std::vector<Vector3d> Vertices() const {
    std::vector<Vector3d> out(count());
    if (!out.empty()) check(mvs_deform_get_vertices(h_, out[0].data()));
    return out;
  }
  std::vector<Vector3d> Normals() const {
    std::vector<Vector3d> out(count());
    if (!out.empty()) check(mvs_deform_get_normals(h_, out[0].data()));
    return out;
  }
  std::vector<int> Facets() const {
    int64_t nf = 0;
    check(mvs_deform_get_facets(h_, nullptr, &nf));
    std::vector<int> f((size_t)nf * 3);
    if (nf) check(mvs_deform_get_facets(h_, f.data(), &nf));
    return f;
  }
  mvs_deform* handle() const { return h_; }

 private:
  size_t count() const { return n_; }
  void set_target(const std::vector<Vector3d>& tpts, const std::vector<Vector3d>& tnormals) {
    check(mvs_deform_set_target(h_, tpts[0].data(), tnormals.empty() ? nullptr : tnormals[0].data(),
                                (int64_t)tpts.size()));
  }
  mvs_deform* h_ = nullptr;
  size_t n_ = 0;
};

// [ref: Deformation/Deformation.h:174-221]
inline void exportOBJ(const std::string& filename, Deformation* deform) {
  check(mvs_deform_export_obj(deform->handle(), filename.c_str()));
}

// [ref: Model2Depth/Model2Depth.h:17-54] process-global static state, like the reference
class Model2Depth {
 public:
  static void SetInput(const std::vector<std::vector<Vector3d>>& points_,
                       const std::vector<std::vector<Vector3d>>& normals_,
                       const std::vector<std::vector<int>>& facets_,
                       const std::vector<std::vector<Camera>>& cameras_,
                       const std::vector<std::string>& path_) {
    (void)normals_;
    S().points = points_;
    S().facets = facets_;
    S().cameras = cameras_;
    S().path = path_;
  }
  // renders every (sequence, frame) and writes <path>DATA/Render/_depth<frame>.raw
  // (Model2Depth.cpp:149-151); the directory must exist (the reference creates it with Win32 calls)
  static void Run(int /*argc*/ = 0, char** /*argv*/ = nullptr) {
    State& s = S();
    s.images.clear();
    for (size_t k = 0; k < s.points.size(); ++k) {
      std::vector<mvs_camera> cams;
      for (const Camera& c : s.cameras[k]) cams.push_back(c.c());
      if (cams.empty()) continue;
      size_t px = (size_t)cams[0].w * cams[0].h;
      std::vector<float> img(px * cams.size());
      check(mvs_model2depth_f64(context(), s.points[k].empty() ? nullptr : s.points[k][0].data(),
                                (int64_t)s.points[k].size(), s.facets[k].data(),
                                (int64_t)s.facets[k].size() / 3, cams.data(), (int)cams.size(),
                                znear, zfar, img.data()));
      if (k < s.path.size() && !s.path[k].empty())
        for (size_t f = 0; f < cams.size(); ++f) {
          std::string fn = s.path[k] + "DATA/Render/_depth" + std::to_string(f) + ".raw";
          mvs_save_depth_raw(fn.c_str(), img.data() + f * px, cams[0].w, cams[0].h);
        }
      s.images.push_back(std::move(img));
    }
  }
  static const std::vector<std::vector<float>>& Images() { return S().images; }
  static constexpr float znear = 0.01f;   // Model2Depth.cpp:19
  static constexpr float zfar = 2000.0f;  // Model2Depth.cpp:20

 private:
  struct State {
    std::vector<std::vector<Vector3d>> points;
    std::vector<std::vector<int>> facets;
    std::vector<std::vector<Camera>> cameras;
    std::vector<std::string> path;
    std::vector<std::vector<float>> images;
  };
  static State& S() {
    static State s;
    return s;
  }
};

// [ref: Processor/Processor.cpp:629-826] the independent iterations of the sequence-pair loop in one
// call: per pair the matched pixels are unprojected, RemoveOutliers + SRTSolver give (s, R, t), the
// Depth2Model mesh of view A is moved into the frame of view B, deformed onto the valid points of
// view B and re-projected with Model2Depth.  Inputs stay where Processor keeps them (host vectors).
struct ViewPair {
  const std::vector<float>* depth_a = nullptr;  // w*h inverse depths (the `.raw` layout)
  const std::vector<float>* depth_b = nullptr;
  Camera cam_a, cam_b;
  std::vector<int> matches;  // 4 ints per match: u_a, v_a, u_b, v_b
  unsigned long long seed = 0;
};
struct ViewPairResult {
  double scale = 1.0, err = 0.0, residual = 0.0;
  Matrix3d R;
  Vector3d t;
  std::vector<int> kept;             // indices of the matches that survive RemoveOutliers
  std::vector<Vector3d> points;      // view-A mesh in the frame of view B
  std::vector<Vector3d> deformed;    // after the deformation onto view B
  std::vector<int> facets;
  std::vector<float> render;         // Model2Depth of the deformed mesh in camera B
};
inline std::vector<ViewPairResult> StitchPairs(const std::vector<ViewPair>& pairs,
                                               const mvs_pair_params* params = nullptr) {
  const size_t n = pairs.size();
  std::vector<mvs_pair_job> jobs(n);
  std::vector<mvs_pair_result> res(n);
  std::vector<ViewPairResult> out(n);
  std::vector<mvs_camera> cams(2 * n);
  for (size_t k = 0; k < n; ++k) {
    const ViewPair& p = pairs[k];
    cams[2 * k] = p.cam_a.c();
    cams[2 * k + 1] = p.cam_b.c();
    const size_t px = (size_t)p.cam_a.W() * p.cam_a.H();
    ViewPairResult& o = out[k];
    o.kept.resize(p.matches.size() / 4);
    o.points.resize(px);
    o.deformed.resize(px);
    o.facets.resize(6 * (size_t)(p.cam_a.W() - 1) * (p.cam_a.H() - 1));
    o.render.resize(px);
    mvs_pair_job& j = jobs[k];
    std::memset(&j, 0, sizeof(j));
    j.depth_a = p.depth_a ? p.depth_a->data() : nullptr;
    j.depth_b = p.depth_b ? p.depth_b->data() : nullptr;
    j.cam_a = &cams[2 * k];
    j.cam_b = &cams[2 * k + 1];
    j.matches = p.matches.data();
    j.n_matches = (int64_t)(p.matches.size() / 4);
    j.seed = p.seed;
    j.kept_idx = o.kept.data();
    j.template_xyz = o.points[0].data();
    j.facets = o.facets.data();
    j.deformed_xyz = o.deformed[0].data();
    j.render = o.render.data();
  }
  check(mvs_stitch_pairs(context(), jobs.data(), (int64_t)n, params, res.data()));
  for (size_t k = 0; k < n; ++k) {
    ViewPairResult& o = out[k];
    o.scale = res[k].scale;
    o.err = res[k].err;
    o.residual = res[k].residual;
    std::memcpy(o.R.data(), res[k].R, sizeof(res[k].R));
    std::memcpy(o.t.data(), res[k].t, sizeof(res[k].t));
    o.kept.resize((size_t)res[k].n_kept);
    o.points.resize((size_t)res[k].n_template);
    o.deformed.resize((size_t)res[k].n_template);
    o.facets.resize(3 * (size_t)res[k].n_facets);
  }
  return out;
}

}  // namespace mvs
#endif  // MVS_SHIM_H
