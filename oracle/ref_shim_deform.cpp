// oracle/ref_shim_deform.cpp — TEST INFRASTRUCTURE ONLY.
//
// C-ABI wrappers around the reference's OWN Deformation class (Deformation/Deformation.h,
// Deformation/Deformation.cpp compiled from its source text; oracle/Makefile target `ref`):
// node sampling (D2), node graph (D3), node correspondences with the 1-NN threshold rule (D4') and
// with the radius / projection rule (D4), the two smoothing sweeps (D5), and the CGAL-path vertex
// normals + exportOBJ (D10).  CGAL, boost and OpenCV-FLANN are stand-ins
// (compat/CGAL/mvs_cgal_standin.h, compat/opencv2\opencv.hpp): nearest-neighbour searches are exact,
// and Surface_mesh_deformation RECORDS the control vertices and target positions it is handed —
// those constraints are what this file returns.  The as-rigid-as-possible solve behind them (row D6)
// is not part of the stand-in.  Marshalling only.
#include <stdint.h>
#include <stdlib.h>
#include <unistd.h>

#include <algorithm>
#include <fstream>
#include <iostream>
#include <iterator>
#include <map>
#include <queue>
#include <sstream>
#include <string>
#include <vector>

#include <Eigen/Dense>
#include <CGAL/mvs_cgal_standin.h>

// sampIdx, poly and KNearestNeighbor are private (Deformation/Deformation.h:246-251)
#define private public
#include "Deformation/Deformation.h"
#undef private

namespace {

typedef std::vector<Eigen::Vector3d> Pts;

Pts to_pts(const double* p, int64_t n) {
  Pts v((size_t)(p ? n : 0));
  for (size_t i = 0; i < v.size(); ++i) v[i] = Eigen::Vector3d(p[3 * i], p[3 * i + 1], p[3 * i + 2]);
  return v;
}
std::vector<int> to_ints(const int32_t* a, int64_t n) { return std::vector<int>(a, a + (a ? n : 0)); }

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

// Deformation::UniformSampling, Deformation/Deformation.cpp:63-106; returns the node count
int64_t ref_deform_uniform_sampling(const double* pts, const double* normals, int64_t n,
                                    const int32_t* facets, int64_t nf, const char* workdir,
                                    int32_t* out_ids) {
  InDir cd(workdir);
  Deformation d(to_pts(pts, n), to_pts(normals, n), to_ints(facets, nf * 3));
  d.UniformSampling();
  for (size_t i = 0; i < d.sampIdx.size(); ++i) out_ids[i] = d.sampIdx[i];
  return (int64_t)d.sampIdx.size();
}

// Deformation::KNearestNeighbor(K), Deformation/Deformation.cpp:108-153; out: m x (K + 1) node indices
void ref_deform_node_graph(const double* pts, const double* normals, int64_t n, const int32_t* facets,
                           int64_t nf, const int32_t* ids, int64_t m, int K, int32_t* out_idx,
                           double* out_w) {
  Deformation d(to_pts(pts, n), to_pts(normals, n), to_ints(facets, nf * 3));
  d.sampIdx = to_ints(ids, m);
  std::vector<std::vector<std::pair<int, double> > > g = d.KNearestNeighbor(K);
  for (int64_t i = 0; i < m; ++i)
    for (int j = 0; j <= K; ++j) {
      out_idx[i * (K + 1) + j] = g[(size_t)i][(size_t)j].first;
      if (out_w) out_w[i * (K + 1) + j] = g[(size_t)i][(size_t)j].second;
    }
}

static int64_t copy_log(int32_t* out_vertex, double* out_target, int* deform_calls) {
  const CGAL::standin::Deformation_log& l = CGAL::standin::deformation_log();
  for (size_t i = 0; i < l.constraints.size(); ++i) {
    out_vertex[i] = (int32_t)l.constraints[i].vertex;
    for (int k = 0; k < 3; ++k) out_target[3 * i + k] = l.constraints[i].target[k];
  }
  if (deform_calls) *deform_calls = l.deform_calls;
  return (int64_t)l.constraints.size();
}

// Deformation::Deform(tpts, tnormals, projLenErr, projDistErr), Deformation/Deformation.cpp:232-402,
// up to the constraints handed to CGAL: one (vertex id, target position) per node, in node order.
int64_t ref_deform_node_targets(const double* pts, const double* normals, int64_t n,
                                const int32_t* facets, int64_t nf, const int32_t* ids, int64_t m,
                                const double* tpts, const double* tnormals, int64_t nt,
                                double proj_len_err, double proj_dist_err, const char* workdir,
                                int32_t* out_vertex, double* out_target) {
  InDir cd(workdir);
  Deformation d(to_pts(pts, n), to_pts(normals, n), to_ints(facets, nf * 3));
  d.sampIdx = to_ints(ids, m);  // empty: Deform runs UniformSampling itself (:246-248)
  d.Deform(to_pts(tpts, nt), to_pts(tnormals, nt), proj_len_err, proj_dist_err);
  return copy_log(out_vertex, out_target, nullptr);
}

// Deformation::Deform(tpts, tnormals, threshold), Deformation/Deformation.cpp:155-230: constraints
// only for the nodes that pass the distance and the cosine test.
int64_t ref_deform_node_targets_threshold(const double* pts, const double* normals, int64_t n,
                                          const int32_t* facets, int64_t nf, const int32_t* ids,
                                          int64_t m, const double* tpts, const double* tnormals,
                                          int64_t nt, double threshold, int32_t* out_vertex,
                                          double* out_target) {
  Deformation d(to_pts(pts, n), to_pts(normals, n), to_ints(facets, nf * 3));
  d.sampIdx = to_ints(ids, m);
  d.Deform(to_pts(tpts, nt), to_pts(tnormals, nt), threshold);
  return copy_log(out_vertex, out_target, nullptr);
}

// exportOBJ(filename, &polyhedron), Deformation/Deformation.h:174-221, with the vertex normals of
// :86-150 (sum of unit facet normals around the vertex, normalised)
void ref_deform_export_obj(const char* filename, const double* pts, const double* normals, int64_t n,
                           const int32_t* facets, int64_t nf) {
  Deformation d(to_pts(pts, n), to_pts(normals, n), to_ints(facets, nf * 3));
  CgalPolyhedron poly = d.Polyhedron();
  exportOBJ(filename, &poly);
}

// 1 if the reference accepts the mesh (Deformation.cpp:42-45 aborts the process otherwise)
int ref_deform_mesh_is_valid(const double* pts, int64_t n, const int32_t* facets, int64_t nf) {
  typedef CgalPolyhedron::HalfedgeDS HDS;
  CgalPolyhedron poly;
  Build_triangle<HDS> bt(to_pts(pts, n), to_ints(facets, nf * 3));
  poly.delegate(bt);
  return poly.is_valid() ? 1 : 0;
}

}  // extern "C"
