// oracle/nn.cpp — TEST INFRASTRUCTURE ONLY (see oracle.h).
//
// Exact nearest neighbours.  The reference uses OpenCV-3.0 FLANN randomized kd-trees
// (approximate, non-deterministic): PartRecognition/PartRecognition.cpp:53-74,
// Deformation/Deformation.cpp:68-97,110-134,245-301.  The contract of the build is what FLANN
// approximates: exact kNN in float32 squared-L2 with lowest-index tie-break.
#include <algorithm>
#include <cmath>
#include <limits>
#include <vector>

#include "oracle.h"

namespace {

// float32, fixed order, no contraction (compiled with -ffp-contract=off)
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

struct Best {
  int k;
  float* d2;
  int32_t* idx;
  int filled;
  void init(int k_, float* d, int32_t* i) {
    k = k_;
    d2 = d;
    idx = i;
    filled = 0;
    for (int j = 0; j < k; ++j) {
      d2[j] = std::numeric_limits<float>::infinity();
      idx[j] = -1;
    }
  }
  inline float worst() const { return d2[k - 1]; }
  inline bool less(float da, int32_t ia, float db, int32_t ib) const {
    if (ib < 0) return true;
    return da < db || (da == db && ia < ib);
  }
  inline void push(float d, int32_t id) {
    if (!less(d, id, d2[k - 1], idx[k - 1])) return;
    int j = k - 1;
    while (j > 0 && less(d, id, d2[j - 1], idx[j - 1])) {
      d2[j] = d2[j - 1];
      idx[j] = idx[j - 1];
      --j;
    }
    d2[j] = d;
    idx[j] = id;
  }
};

struct KdTree {
  const float* pts;
  std::vector<int32_t> order;
  struct Node {
    int32_t lo, hi;     // range in order
    int32_t left, right;  // children or -1
    int axis;
    float split;
  };
  std::vector<Node> nodes;
  static constexpr int kLeaf = 12;

  int build(int32_t lo, int32_t hi) {
    Node nd;
    nd.lo = lo;
    nd.hi = hi;
    nd.left = nd.right = -1;
    nd.axis = 0;
    nd.split = 0;
    int id = (int)nodes.size();
    nodes.push_back(nd);
    if (hi - lo <= kLeaf) return id;
    float mn[3], mx[3];
    for (int a = 0; a < 3; ++a) {
      mn[a] = std::numeric_limits<float>::infinity();
      mx[a] = -mn[a];
    }
    for (int32_t i = lo; i < hi; ++i)
      for (int a = 0; a < 3; ++a) {
        float v = pts[3 * (int64_t)order[i] + a];
        mn[a] = std::min(mn[a], v);
        mx[a] = std::max(mx[a], v);
      }
    int axis = 0;
    for (int a = 1; a < 3; ++a)
      if (mx[a] - mn[a] > mx[axis] - mn[axis]) axis = a;
    if (!(mx[axis] > mn[axis])) return id;  // all coincide: keep as a (big) leaf
    int32_t mid = lo + (hi - lo) / 2;
    std::nth_element(order.begin() + lo, order.begin() + mid, order.begin() + hi,
                     [&](int32_t a, int32_t b) {
                       float va = pts[3 * (int64_t)a + axis], vb = pts[3 * (int64_t)b + axis];
                       return va < vb || (va == vb && a < b);
                     });
    float split = pts[3 * (int64_t)order[mid] + axis];
    nodes[id].axis = axis;
    nodes[id].split = split;
    int l = build(lo, mid);
    int r = build(mid, hi);
    nodes[id].left = l;
    nodes[id].right = r;
    return id;
  }

  void search(int nid, const float* q, Best& best) const {
    const Node& nd = nodes[nid];
    if (nd.left < 0) {
      for (int32_t i = nd.lo; i < nd.hi; ++i) {
        int32_t id = order[i];
        best.push(dist2_f32(pts + 3 * (int64_t)id, q), id);
      }
      return;
    }
    float diff = q[nd.axis] - nd.split;
    int first = diff < 0 ? nd.left : nd.right;
    int second = diff < 0 ? nd.right : nd.left;
    search(first, q, best);
    // left holds coords <= split, right holds coords >= split: every point on the far side is at
    // least |diff| away along the axis; fl(diff*diff) is a lower bound of its fl(d2)
    float plane = diff * diff;
    if (!(plane > best.worst())) search(second, q, best);
  }
};

}  // namespace

extern "C" {

void orc_knn_brute(const float* q, int64_t nq, const float* t, int64_t nt, int k,
                   int32_t* out_idx, float* out_d2) {
  std::vector<float> tmpd((size_t)k);
  std::vector<int32_t> tmpi((size_t)k);
  for (int64_t i = 0; i < nq; ++i) {
    Best b;
    b.init(k, tmpd.data(), tmpi.data());
    const float* qq = q + 3 * i;
    for (int64_t j = 0; j < nt; ++j) b.push(dist2_f32(t + 3 * j, qq), (int32_t)j);
    for (int j = 0; j < k; ++j) {
      if (out_idx) out_idx[i * k + j] = tmpi[j];
      if (out_d2) out_d2[i * k + j] = tmpd[j];
    }
  }
}

// A tree that outlives one query batch: the Gauss-Newton loop searches the same target set every
// iteration (the reference builds its FLANN index once per Deform call, Deformation.cpp:245-252).
void* orc_kdtree_build(const float* t, int64_t nt) {
  KdTree* tree = new KdTree;
  tree->pts = t;
  tree->order.resize((size_t)nt);
  for (int64_t i = 0; i < nt; ++i) tree->order[i] = (int32_t)i;
  tree->nodes.reserve((size_t)(nt / 4 + 16));
  if (nt > 0) tree->build(0, (int32_t)nt);
  return tree;
}

void orc_kdtree_free(void* handle) { delete static_cast<KdTree*>(handle); }

void orc_kdtree_query(const void* handle, const float* q, int64_t nq, int k, int32_t* out_idx,
                      float* out_d2) {
  const KdTree& tree = *static_cast<const KdTree*>(handle);
  const bool any = !tree.order.empty();
  // queries are independent; the OpenMP courtesy build (liboracle_mt.so) spreads them over threads,
  // the default single-thread build ignores the pragmas
#pragma omp parallel
  {
    std::vector<float> tmpd((size_t)k);
    std::vector<int32_t> tmpi((size_t)k);
#pragma omp for schedule(dynamic, 2048)
    for (int64_t i = 0; i < nq; ++i) {
      Best b;
      b.init(k, tmpd.data(), tmpi.data());
      if (any) tree.search(0, q + 3 * i, b);
      for (int j = 0; j < k; ++j) {
        if (out_idx) out_idx[i * k + j] = tmpi[j];
        if (out_d2) out_d2[i * k + j] = tmpd[j];
      }
    }
  }
}

void orc_knn_kdtree(const float* q, int64_t nq, const float* t, int64_t nt, int k,
                    int32_t* out_idx, float* out_d2) {
  void* tree = orc_kdtree_build(t, nt);
  orc_kdtree_query(tree, q, nq, k, out_idx, out_d2);
  orc_kdtree_free(tree);
}

// PartRecognition/PartRecognition.cpp:50-77: template -> float32 (:54-60), query cast to float
// (:66), 1-NN, copy the label.
void orc_part_recog(const double* tmpl, int64_t nt, const int32_t* labels, const double* query,
                    int64_t nq, int32_t* out_labels) {
  std::vector<float> tf((size_t)nt * 3), qf((size_t)nq * 3);
  for (int64_t i = 0; i < nt * 3; ++i) tf[i] = (float)tmpl[i];
  for (int64_t i = 0; i < nq * 3; ++i) qf[i] = (float)query[i];
  std::vector<int32_t> nn((size_t)nq);
  orc_knn_kdtree(qf.data(), nq, tf.data(), nt, 1, nn.data(), nullptr);
  for (int64_t i = 0; i < nq; ++i) out_labels[i] = labels[nn[i]];
}

}  // extern "C"
