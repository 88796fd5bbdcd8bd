// oracle/ref_shim_processor.cpp — TEST INFRASTRUCTURE ONLY.
//
// C-ABI wrappers around two stages of the reference's OWN Processor class, compiled from the source
// text of Processor/Processor.cpp (oracle/Makefile target `ref`):
//   Processor::RemoveOutliers        Processor/Processor.cpp:177-269  
//   Processor::CheckConsistencyCore  Processor/Processor.cpp:66-115
// Processor.cpp as a whole also drives feature detection (SiftGPU), reconstruction (the proprietary
// ZJU::GeoRec), OpenGL rendering and OpenCV image windows; none of that is in this image and none of
// it is on the compared path.  Their HEADERS parse against the stand-ins in oracle/compat/, and the
// few entry points Processor.cpp references are defined below as traps: reaching one is a test bug.
// Marshalling only; every number comes out of the reference's code.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <string>
#include <vector>

#include <Eigen/Dense>

// RemoveOutliers / CheckConsistencyCore are private, and so are Image3D's point table and camera
// (Processor/Processor.h:19-58, Image3D/Image3D.h:31-41): the tests fill an Image3D directly
// instead of loading a .jpg / .raw pair from disk
#define private public
#include "Processor/Processor.h"
#undef private
#include "Common/Utils.h"
#include "FeatureProc/FeatureProc.h"
#include "Model2Depth/Model2Depth.h"
#include "Parameter/ParamParser.h"
#include "Reconstruction/GeometryRec.h"

namespace {

[[noreturn]] void out_of_scope(const char* what) {
  fprintf(stderr, "oracle/_ref: %s is outside the compared path and has no implementation here\n", what);
  abort();
}

struct RefCam {  // same layout as orc_camera / mvs_camera
  double K[9];
  double R[9];
  double t[3];
  int32_t w, h;
};

Camera make_camera(const RefCam* c, int w, int h) {
  Camera cam;
  Eigen::Matrix3d K, R;
  Eigen::Vector3d t;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) {
      K(i, j) = c->K[3 * i + j];
      R(i, j) = c->R[3 * i + j];
    }
    t[i] = c->t[i];
  }
  cam.SetK(K);
  cam.SetRT(R, t);
  cam.SetW(w);
  cam.SetH(h);
  return cam;
}

}  // namespace

// ---- traps for the out-of-scope subsystems Processor.cpp links against
void FeatureProc::DetectFeature(const std::vector<std::string>,
                                std::vector<std::vector<SiftGPU::SiftKeypoint> >&,
                                std::vector<std::vector<float> >&) {
  out_of_scope("FeatureProc::DetectFeature (SiftGPU)");
}
void FeatureProc::MatchFeature(
    const std::vector<std::vector<SiftGPU::SiftKeypoint> >&,
    const std::vector<std::vector<SiftGPU::SiftKeypoint> >&, const std::vector<std::vector<float> >&,
    const std::vector<std::vector<float> >&,
    std::vector<std::vector<std::vector<std::pair<Eigen::Vector2i, Eigen::Vector2i> > > >&) {
  out_of_scope("FeatureProc::MatchFeature (SiftGPU)");
}
void GeometryRec::Init(const std::string, const double, const double, const int, const double,
                       const double, const int, const int, const int, const int, const double,
                       const int, const int) {
  out_of_scope("GeometryRec::Init (ZJU::GeoRec)");
}
bool GeometryRec::RunPoisson(const std::string) { out_of_scope("GeometryRec::RunPoisson (ZJU::GeoRec)"); }
void Image3D::LoadModel(const std::string, const std::string, const Camera&) {
  out_of_scope("Image3D::LoadModel (OpenCV image I/O, GrabCut)");
}
std::vector<std::vector<Eigen::Vector3f> > Model2Depth::points;
std::vector<std::vector<Eigen::Vector3f> > Model2Depth::normals;
std::vector<std::vector<int> > Model2Depth::facets;
std::vector<std::vector<Camera> > Model2Depth::cameras;
std::vector<std::string> Model2Depth::path;
int Model2Depth::w = 0;
int Model2Depth::h = 0;
void Model2Depth::Run(int, char*[]) { out_of_scope("Model2Depth::Run (OpenGL / GLUT)"); }

extern "C" {

// seed the C library's generator the reference's Shuffle draws from (Common/Utils.h:25-34)
void ref_srand(unsigned seed) { srand(seed); }

// the next iter_num hypothesis samples SRTSolver::EstimateRTRansac would draw for `size` matches
// (Solver/SRTSolver.cpp:152) from the CURRENT state of that generator; signature of orc_triples_fn,
// so the oracle's round loop can be handed the reference's own stream
void ref_srt_triples_next(int64_t size, int iter_num, int32_t* triples, void*) {
  for (int k = 0; k < iter_num; ++k) {
    int idx[3];
    Shuffle(idx, (int)size, 3);
    for (int j = 0; j < 3; ++j) triples[3 * k + j] = idx[j];
  }
}

// Processor::RemoveOutliers after srand(seed): the 3-D matches (p_i, q_i) are presented as the
// pixels (i, 0) of two one-row Image3D objects whose point tables hold p and q.  Returns the number
// of surviving matches; keep_idx = their original indices.
int64_t ref_remove_outliers(const double* p, const double* q, int64_t n, const RefCam* c1,
                            const RefCam* c2, double pixel_err, double adapt_ratio, unsigned seed,
                            int32_t* keep_idx, double* err, double* inlier_ratio) {
  ParamParser::pixel_err = pixel_err;
  ParamParser::adapt_pixel_err_ratio = adapt_ratio;
  Image3D im, jm;
  im.cam = make_camera(c1, (int)n, 1);  // W() only indexes the point table (Image3D.h:15)
  jm.cam = make_camera(c2, (int)n, 1);
  im.point3d.resize((size_t)n);
  jm.point3d.resize((size_t)n);
  std::vector<std::pair<Eigen::Vector2i, Eigen::Vector2i> > matches((size_t)n);
  for (int64_t i = 0; i < n; ++i) {
    im.point3d[(size_t)i] = Eigen::Vector3d(p[3 * i], p[3 * i + 1], p[3 * i + 2]);
    jm.point3d[(size_t)i] = Eigen::Vector3d(q[3 * i], q[3 * i + 1], q[3 * i + 2]);
    Eigen::Vector2i px;
    px[0] = (int)i;
    px[1] = 0;
    matches[(size_t)i] = std::make_pair(px, px);
  }
  srand(seed);
  Processor proc;
  double ir = 0.0, e = 0.0;
  proc.RemoveOutliers(im, jm, ir, e, matches);
  for (size_t k = 0; k < matches.size(); ++k) keep_idx[k] = matches[k].first[0];
  if (err) *err = e;
  if (inlier_ratio) *inlier_ratio = ir;
  return (int64_t)matches.size();
}

// Processor::CheckConsistencyCore: depth (w*h inverse depths, in/out) of the current view against
// nref reference views
void ref_check_consistency(double* depth, const RefCam* cam, const double* ref_depths,
                           const RefCam* ref_cams, int nref, double min_dsp, double max_dsp,
                           int reproj_err) {
  ParamParser::m_fMinDsp = min_dsp;
  ParamParser::m_fMaxDsp = max_dsp;
  ParamParser::reproj_err = reproj_err;
  const int w = cam->w, h = cam->h;
  Camera cur = make_camera(cam, w, h);
  std::vector<Camera> refs;
  std::vector<std::vector<double> > rd((size_t)nref);
  for (int k = 0; k < nref; ++k) {
    refs.push_back(make_camera(ref_cams + k, ref_cams[k].w, ref_cams[k].h));
    rd[(size_t)k].assign(ref_depths + (size_t)k * w * h, ref_depths + (size_t)(k + 1) * w * h);
  }
  std::vector<double> d(depth, depth + (size_t)w * h);
  Processor proc;
  proc.CheckConsistencyCore(cur, refs, d, rd);
  for (size_t i = 0; i < d.size(); ++i) depth[i] = d[i];
}

}  // extern "C"
