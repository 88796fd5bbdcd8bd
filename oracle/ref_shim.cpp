// oracle/ref_shim.cpp — TEST INFRASTRUCTURE ONLY.
//
// C-ABI wrappers around the reference's OWN first-party classes, compiled unmodified from
// /root/reference/MultiViewStitch (Solver/SRTSolver.cpp, Camera/Camera.cpp,
// SetUtils/PointSetUtils.cpp, SetUtils/UnionSetUtils.cpp, Depth2Model/Depth2Model.cpp and the
// static helpers of Common/Utils.h) against the stand-in headers in oracle/compat/ (Eigen, Win32 and
// OpenCV are not in this image).  The result, oracle/_ref/libmvsref.so, is the reference itself for
// this arithmetic: tests/test_ref_parity.py checks the restatement in oracle/*.cpp against it, and
// tests/golden/ref_*.npz holds vectors generated from it (tests/golden/make_ref_golden.py) so the
// check also runs where /root/reference does not exist.  This file contains marshalling only; every
// number comes out of the reference's code.  What the stand-in Eigen contributes is the 3x3
// SVD / symmetric eigen-decomposition (any correct decomposition gives the same R = V U^T and the
// same pivots up to sign).
#include <stdint.h>
#include <stdlib.h>

#include <string>
#include <vector>

#include "Camera/Camera.h"
#include "Common/Utils.h"
#include "Depth2Model/Depth2Model.h"
#include "SetUtils/PointSetUtils.h"
#include "SetUtils/UnionSetUtils.h"
#include "Solver/SRTSolver.h"

namespace {

struct RefCam {  // same layout as orc_camera / mvs_camera
  double K[9];
  double R[9];
  double t[3];
  int32_t w, h;
};

Camera make_camera(const RefCam* c) {
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
  cam.SetW(c->w);
  cam.SetH(c->h);
  return cam;
}

std::vector<std::pair<Eigen::Vector3d, Eigen::Vector3d>> make_matches(const double* p,
                                                                      const double* q, int64_t n) {
  std::vector<std::pair<Eigen::Vector3d, Eigen::Vector3d>> m((size_t)n);
  for (int64_t i = 0; i < n; ++i) {
    m[i].first = Eigen::Vector3d(p[3 * i], p[3 * i + 1], p[3 * i + 2]);
    m[i].second = Eigen::Vector3d(q[3 * i], q[3 * i + 1], q[3 * i + 2]);
  }
  return m;
}

void put(const Eigen::Matrix3d& R, const Eigen::Vector3d& t, double* Ro, double* to) {
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) Ro[3 * i + j] = R(i, j);
    to[i] = t[i];
  }
}

RefCam identity_cam() {
  RefCam c{};
  c.K[0] = c.K[4] = c.K[8] = 1.0;
  c.R[0] = c.R[4] = c.R[8] = 1.0;
  c.w = c.h = 1;
  return c;
}

}  // namespace

extern "C" {

// SRTSolver::EstimateTransform(double&, Matrix3d&, Vector3d&)  Solver/SRTSolver.cpp:272-275
void ref_srt_estimate(const double* p, const double* q, int64_t n, const RefCam* c1,
                      const RefCam* c2, double* scale, double R[9], double t[3]) {
  RefCam id = identity_cam();
  SRTSolver s(1);
  s.SetPrintFlag(false);
  s.SetInput(make_matches(p, q, n), make_camera(c1 ? c1 : &id), make_camera(c2 ? c2 : &id));
  Eigen::Matrix3d Rm;
  Eigen::Vector3d tm;
  s.EstimateTransform(*scale, Rm, tm);
  put(Rm, tm, R, t);
}

// SRTSolver::ResidualError  Solver/SRTSolver.cpp:6-29
double ref_srt_residual(const double* p, const double* q, int64_t n, const RefCam* c1,
                        const RefCam* c2, double scale, const double R[9], const double t[3]) {
  SRTSolver s(1);
  s.SetPrintFlag(false);
  s.SetInput(make_matches(p, q, n), make_camera(c1), make_camera(c2));
  Eigen::Matrix3d Rm;
  Eigen::Vector3d tm;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) Rm(i, j) = R[3 * i + j];
    tm[i] = t[i];
  }
  return s.ResidualError(scale, Rm, tm);
}

// The hypothesis samples SRTSolver::EstimateRTRansac draws after srand(seed): iter_num calls of
// Shuffle(idx, n, 3)  (Common/Utils.h:25-34, Solver/SRTSolver.cpp:152-153), replayed here so the
// same triples can be handed to the oracle and to the CUDA path.
void ref_srt_triples(int64_t n, int iter_num, unsigned seed, int32_t* triples) {
  srand(seed);
  for (int k = 0; k < iter_num; ++k) {
    int idx[3];
    Shuffle(idx, (int)n, 3);
    for (int j = 0; j < 3; ++j) triples[3 * k + j] = idx[j];
  }
}

// SRTSolver::EstimateTransformRansac  Solver/SRTSolver.cpp:277-280 after srand(seed);
// array != 0: the EstimateTransform(double&, double[3][3], double[3]) overload (:256-270).
void ref_srt_ransac(const double* p, const double* q, int64_t n, const RefCam* c1, const RefCam* c2,
                    int iter_num, unsigned seed, int array, double* scale, double R[9],
                    double t[3]) {
  SRTSolver s(iter_num);
  s.SetPrintFlag(false);
  s.SetInput(make_matches(p, q, n), make_camera(c1), make_camera(c2));
  srand(seed);
  if (array) {
    double Ra[3][3], ta[3];
    s.EstimateTransform(*scale, Ra, ta);
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) R[3 * i + j] = Ra[i][j];
      t[i] = ta[i];
    }
  } else {
    Eigen::Matrix3d Rm;
    Eigen::Vector3d tm;
    s.EstimateTransformRansac(*scale, Rm, tm);
    put(Rm, tm, R, t);
  }
}

// Camera::GetImgCoordFromWorld / GetWorldCoordFromImg  Camera/Camera.cpp:40-72
void ref_project(const RefCam* c, const double* pts, int64_t n, int32_t* uv) {
  Camera cam = make_camera(c);
  for (int64_t i = 0; i < n; ++i) {
    int u, v;
    cam.GetImgCoordFromWorld(Eigen::Vector3d(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]), u, v);
    uv[2 * i] = u;
    uv[2 * i + 1] = v;
  }
}
void ref_unproject(const RefCam* c, const int32_t* uv, const double* depth, int64_t n,
                   double* out) {
  Camera cam = make_camera(c);
  for (int64_t i = 0; i < n; ++i) {
    Eigen::Vector3d p;
    cam.GetWorldCoordFromImg(uv[2 * i], uv[2 * i + 1], depth[i], p);
    out[3 * i] = p[0];
    out[3 * i + 1] = p[1];
    out[3 * i + 2] = p[2];
  }
}
// Camera::GetObjAbsTransformGL / GetFrustumGL / GetProjectGL  Camera/Camera.cpp:6-38; matrices
// row-major (element (i,j) at 4 i + j), frustum = left, right, bottom, top
void ref_gl_matrices(const RefCam* c, float znear, float zfar, float modelview[16],
                     float frustum[4], float proj[16]) {
  Camera cam = make_camera(c);
  Eigen::Matrix4f mv = cam.GetObjAbsTransformGL();
  float l, r, b, t;
  cam.GetFrustumGL(znear, l, r, b, t);
  Eigen::Matrix4f pr = cam.GetProjectGL(l, r, b, t, znear, zfar);
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      modelview[4 * i + j] = mv(i, j);
      proj[4 * i + j] = pr(i, j);
    }
  frustum[0] = l;
  frustum[1] = r;
  frustum[2] = b;
  frustum[3] = t;
}

// PointSetUtils::SetInput / GetBarycenter / GetBoundingBoxCenter / CalcPivots(pivots, 3)
// SetUtils/PointSetUtils.cpp:3-62; pivots row-major 3x3 (column i = i-th pivot)
void ref_pointset(const double* pts, int64_t n, double bary[3], double bbox_center[3],
                  double pivots[9]) {
  std::vector<Eigen::Vector3d> v((size_t)n);
  for (int64_t i = 0; i < n; ++i) v[i] = Eigen::Vector3d(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
  PointSetUtils ps;
  ps.SetInput(v);
  Eigen::Vector3d b = ps.GetBarycenter(), c = ps.GetBoundingBoxCenter();
  Eigen::MatrixXd P;
  ps.CalcPivots(P, 3);
  for (int i = 0; i < 3; ++i) {
    bary[i] = b[i];
    bbox_center[i] = c[i];
    for (int j = 0; j < 3; ++j) pivots[3 * i + j] = P(i, j);
  }
}

// UnionSet driven exactly as Alignment::RetainConnectRegion drives it (Alignment/Alignment.cpp:
// 621-626): Merge(f0,f1), Merge(f0,f2) per facet, ProminentRepresent(); keep[i] = (root == Find(i)).
// Returns the number of connected components (UnionSet::ConnectRegion).
int ref_connected_keep(int64_t n, const int32_t* facets, int64_t nf, uint8_t* keep) {
  UnionSet us((int)n);
  for (int64_t i = 0; i < 3 * nf; i += 3) {
    us.Merge(facets[i], facets[i + 1]);
    us.Merge(facets[i], facets[i + 2]);
  }
  int r = us.ProminentRepresent();
  for (int64_t i = 0; i < n; ++i) keep[i] = (r == us.Find((int)i)) ? 1 : 0;
  return us.ConnectRegion();
}

// Depth2Model(min, max, smooth).SaveModel(depth, cam)  Depth2Model/Depth2Model.cpp:7-82
void ref_depth2model(const float* inv_depth, const RefCam* c, double min_dsp, double max_dsp,
                     double smooth, double* out_xyz, int32_t* tex, int32_t* facets,
                     int64_t* n_points, int64_t* n_facets) {
  Camera cam = make_camera(c);
  std::vector<double> depth((size_t)c->w * c->h);
  for (size_t i = 0; i < depth.size(); ++i) depth[i] = inv_depth[i];  // as LoadDepth, Utils.h:166-175
  Depth2Model d2m(min_dsp, max_dsp, smooth);
  d2m.SaveModel(depth, cam);
  *n_points = (int64_t)d2m.point3d.size();
  *n_facets = (int64_t)d2m.facets.size() / 3;
  if (out_xyz)
    for (size_t i = 0; i < d2m.point3d.size(); ++i)
      for (int a = 0; a < 3; ++a) out_xyz[3 * i + a] = d2m.point3d[i][a];
  if (tex)
    for (size_t i = 0; i < d2m.texIndex.size(); ++i) tex[i] = d2m.texIndex[i];
  if (facets)
    for (size_t i = 0; i < d2m.facets.size(); ++i) facets[i] = d2m.facets[i];
}

// Common/Utils.h:140-149 CalcRotation, :151-162 AxisAngleTransform
void ref_calc_rotation(const double a[3], const double b[3], double R[9]) {
  Eigen::Matrix3d Rm;
  CalcRotation(Eigen::Vector3d(a[0], a[1], a[2]), Eigen::Vector3d(b[0], b[1], b[2]), Rm);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R[3 * i + j] = Rm(i, j);
}
void ref_axis_angle(const double axis[3], double angle_deg, const double pb[3], double pa[3]) {
  Eigen::Vector3d out;
  AxisAngleTransform(Eigen::Vector3d(axis[0], axis[1], axis[2]), angle_deg,
                     Eigen::Vector3d(pb[0], pb[1], pb[2]), out);
  for (int i = 0; i < 3; ++i) pa[i] = out[i];
}

// Common/Utils.h:166-185 LoadDepth / SaveDepth (the `.raw` wire format)
void ref_save_depth(const char* filename, const double* depth, int64_t n) {
  std::vector<double> d(depth, depth + n);
  SaveDepth(filename, d);
}
void ref_load_depth(const char* filename, double* depth, int w, int h) {
  std::vector<double> d;
  LoadDepth(filename, d, w, h);
  for (size_t i = 0; i < d.size(); ++i) depth[i] = d[i];
}

}  // extern "C"
