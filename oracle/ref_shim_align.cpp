// oracle/ref_shim_align.cpp — TEST INFRASTRUCTURE ONLY.
//
// C-ABI wrappers around the reference's OWN Alignment, PartRecognition and PlyObj (Mesh) classes.
// Their .cpp files are compiled from the reference's source text as it lies under /root/reference
// (Alignment/Alignment.cpp, PartRecognition/PartRecognition.cpp, PlyObj/PlyObj.cpp,
// Parameter/ParamParser.cpp); g++ has no switch for ONE MSVC extension their two headers use —
// a temporary bound to a non-const reference as a default argument (Alignment/Alignment.h:31-36,
// PlyObj/PlyObj.h:89-90,237-238,293-294,345-346) — so oracle/Makefile passes those two headers
// through a single sed rule (`T &x = T()` -> `T &x = msvc_default_arg<T >()`, compat/msvc_prelude.h)
// into a temporary directory that exists only while the recipe runs.  No arithmetic is touched and
// nothing of the reference is stored in the repository.
//
// cv::flann (the approximate nearest-neighbour search inside PartRecognition::PartRecog) is the
// exact stand-in of compat/opencv2\opencv.hpp; Eigen is compat/Eigen/Dense.  This file contains
// marshalling only.
#include <stdint.h>
#include <stdlib.h>
#include <unistd.h>

#include <fstream>
#include <iostream>
#include <iterator>
#include <map>
#include <queue>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

#include <Eigen/Dense>

// the reference keeps the stages this file calls one by one private (Alignment/Alignment.h:37-73,
// PlyObj/PlyObj.h Mesh, PartRecognition.h:50-52); access specifiers do not change layout or symbols
#define private public
#define protected public
#include "Alignment/Alignment.h"
#include "PartRecognition/PartRecognition.h"
#include "PlyObj/PlyObj.h"
#undef private
#undef protected
#include "Parameter/ParamParser.h"

namespace {

typedef std::vector<Eigen::Vector3d> Pts;

Pts to_pts(const double* p, int64_t n) {
  Pts v((size_t)(p ? n : 0));
  for (size_t i = 0; i < v.size(); ++i) v[i] = Eigen::Vector3d(p[3 * i], p[3 * i + 1], p[3 * i + 2]);
  return v;
}
void from_pts(const Pts& v, double* p) {
  if (!p) return;
  for (size_t i = 0; i < v.size(); ++i)
    for (int k = 0; k < 3; ++k) p[3 * i + k] = v[i][k];
}
std::vector<int> to_ints(const int32_t* a, int64_t n) {
  return std::vector<int>(a, a + (a ? n : 0));
}
void put4(const Eigen::Matrix4d& T, double* out) {
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) out[4 * i + j] = T(i, j);
}

// run f with the process's working directory set to dir (Alignment::Align opens
// "./Template/part/parts" and writes debug meshes to "./Result/..." — Alignment.cpp:35-40)
struct InDir {
  char old[4096];
  bool ok;
  explicit InDir(const char* dir) : ok(false) {
    if (dir && getcwd(old, sizeof(old)) && chdir(dir) == 0) ok = true;
  }
  ~InDir() {
    if (ok && chdir(old) != 0) abort();
  }
};

}  // namespace

extern "C" {

// Alignment::RetainConnectRegion, Alignment/Alignment.cpp:618-654
void ref_retain_connected(double* pts, double* normals, int64_t* n, int32_t* facets, int64_t* nf) {
  Pts p = to_pts(pts, *n), nr = to_pts(normals, *n);
  std::vector<int> f = to_ints(facets, *nf * 3);
  Alignment a;
  a.RetainConnectRegion(p, nr, f);
  from_pts(p, pts);
  if (nr.size() >= p.size()) {
    nr.resize(p.size());  // the reference leaves the tail of `normals` in place (:641-643)
    from_pts(nr, normals);
  }
  *n = (int64_t)p.size();
  *nf = (int64_t)f.size() / 3;
  for (size_t i = 0; i < f.size(); ++i) facets[i] = f[i];
}

// Alignment::RemoveGround, Alignment/Alignment.cpp:79-233 (ends with RetainConnectRegion)
void ref_remove_ground(double* pts, double* normals, int64_t* n, int32_t* facets, int64_t* nf,
                       double dist_thres, const char* workdir, double ground_ray[3]) {
  InDir cd(workdir);
  ParamParser::dist_thres = dist_thres;
  Pts p = to_pts(pts, *n), nr = to_pts(normals, *n);
  std::vector<int> f = to_ints(facets, *nf * 3);
  Alignment a;
  Eigen::Vector3d g = a.RemoveGround(p, nr, f);
  from_pts(p, pts);
  nr.resize(p.size());
  from_pts(nr, normals);
  *n = (int64_t)p.size();
  *nf = (int64_t)f.size() / 3;
  for (size_t i = 0; i < f.size(); ++i) facets[i] = f[i];
  for (int k = 0; k < 3; ++k) ground_ray[k] = g[k];
}

// Alignment::InitAlignment, Alignment/Alignment.cpp:235-314
void ref_init_alignment(const double* src, int64_t ns, const double* tgt, int64_t nt,
                        const double ground_ray[3], const double view_ray[3], double transform[16],
                        double* scale) {
  Pts s = to_pts(src, ns), t = to_pts(tgt, nt), none;
  Eigen::Matrix4d T;
  double sc = 1.0;
  Alignment a;
  a.InitAlignment(s, none, t, none, Eigen::Vector3d(ground_ray[0], ground_ray[1], ground_ray[2]),
                  Eigen::Vector3d(view_ray[0], view_ray[1], view_ray[2]), T, sc);
  put4(T, transform);
  *scale = sc;
}

// Alignment::LocalAlignmentCore, Alignment/Alignment.cpp:423-546
void ref_local_alignment_core(const double* src, const int32_t* s_labels, int64_t ns,
                              const double* tgt, const int32_t* t_labels, int64_t nt, int slabel,
                              int tlabel, double transform[16], double* scale) {
  Pts s = to_pts(src, ns), t = to_pts(tgt, nt);
  std::vector<int> sl = to_ints(s_labels, ns), tl = to_ints(t_labels, nt);
  Eigen::Matrix4d T;
  double sc = 1.0;
  Alignment a;
  a.LocalAlignmentCore(s, sl, t, tl, slabel, tlabel, T, sc);
  put4(T, transform);
  *scale = sc;
}

// PartRecognition::PartRecog, PartRecognition/PartRecognition.cpp:50-77
void ref_part_recog(const double* tmpl, const int32_t* labels, int64_t nt, const double* query,
                    int64_t nq, int32_t* out) {
  PartRecognition pr;
  pr.SetInput(to_pts(tmpl, nt));
  pr.parts = to_ints(labels, nt);
  std::vector<int> o;
  pr.PartRecog(to_pts(query, nq), o);
  for (int64_t i = 0; i < nq; ++i) out[i] = o[(size_t)i];
}

// Alignment::Align, Alignment/Alignment.cpp:11-76, as a whole.  workdir must hold
// Template/part/parts (the template's labels in the reference's own file format,
// PartRecognition.cpp:7-48); src/s_normals are transformed in place, tgt/t_normals/t_facets are
// trimmed in place (ground removal + largest connected region).
void ref_align(double* src, double* s_normals, int64_t ns, const int32_t* s_facets, int64_t nsf,
               double* tgt, double* t_normals, int64_t* nt, int32_t* t_facets, int64_t* ntf,
               const double view_ray[3], double dist_thres, const char* workdir) {
  InDir cd(workdir);
  ParamParser::dist_thres = dist_thres;
  Pts s = to_pts(src, ns), sn = to_pts(s_normals, ns), t = to_pts(tgt, *nt), tn = to_pts(t_normals, *nt);
  std::vector<int> sf = to_ints(s_facets, nsf * 3), tf = to_ints(t_facets, *ntf * 3);
  Alignment a;
  a.Align(s, sn, sf, t, tn, tf, Eigen::Vector3d(view_ray[0], view_ray[1], view_ray[2]));
  from_pts(s, src);
  from_pts(sn, s_normals);
  from_pts(t, tgt);
  tn.resize(t.size());
  from_pts(tn, t_normals);
  *nt = (int64_t)t.size();
  *ntf = (int64_t)tf.size() / 3;
  for (size_t i = 0; i < tf.size(); ++i) t_facets[i] = tf[i];
}

// Mesh::CalculateVertexNormals, PlyObj/PlyObj.cpp:139-185
void ref_mesh_vertex_normals(const double* pts, int64_t n, const int32_t* facets, int64_t nf,
                             double* out) {
  Mesh m;
  m.Vertices = to_pts(pts, n);
  m.Facets = to_ints(facets, nf * 3);
  m.CalculateVertexNormals();
  from_pts(m.VertexNormals, out);
}

// Mesh::WriteObj, PlyObj/PlyObj.cpp:18-27,77-137
void ref_write_obj(const char* filename, const double* pts, const double* normals, int64_t n,
                   const int32_t* facets, int64_t nf) {
  Mesh m;
  m.WriteObj(filename, to_pts(pts, n), to_pts(normals, n), to_ints(facets, nf * 3));
}

// Mesh::ReadObj, PlyObj/PlyObj.cpp:3-16,29-75; returns -1 when a capacity is too small
int ref_read_obj(const char* filename, double* pts, double* normals, int64_t cap_n, int32_t* facets,
                 int64_t cap_f, int64_t* n, int64_t* nf) {
  Mesh m;
  Pts p, nr;
  std::vector<int> f;
  m.ReadObj(filename, p, nr, f);
  *n = (int64_t)p.size();
  *nf = (int64_t)f.size() / 3;
  if (*n > cap_n || *nf > cap_f || nr.size() != p.size()) return -1;
  from_pts(p, pts);
  from_pts(nr, normals);
  for (size_t i = 0; i < f.size(); ++i) facets[i] = f[i];
  return 0;
}

}  // extern "C"
