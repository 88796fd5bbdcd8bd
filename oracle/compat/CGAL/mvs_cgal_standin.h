// oracle/compat/CGAL/mvs_cgal_standin.h — TEST INFRASTRUCTURE ONLY.
//
// A minimal stand-in for the part of CGAL 4.6 (and the two boost names) that the reference's
// Deformation/Deformation.h + Deformation.cpp mention, so that those files compile from
// /root/reference with g++ (oracle/Makefile target `ref`; CGAL and boost are not in this image).
// Written from the CGAL reference manual; not CGAL code.  What it provides:
//   * Simple_cartesian<double> points / vectors, cross_product, NULL_VECTOR;
//   * Polyhedron_3 as an indexed triangle soup with vertex ids in insertion order (what
//     Polyhedron_incremental_builder_3 + set_halfedgeds_items_id give), vertex / facet / point
//     iteration, facet and vertex circulators over the facets' own halfedges (border halfedges are
//     not materialised: the reference skips them, Deformation.h:99);
//   * Surface_mesh_deformation as a RECORDER: it keeps the control vertices and target positions the
//     reference hands to CGAL (Deformation.cpp:206-207, :389-390) — the output of rows D2-D5 — and
//     does not move the mesh.  CGAL's as-rigid-as-possible solve itself (row D6) is NOT restated
//     here; it stays "parity unpinned" (oracle/arap.cpp restates the published algorithm);
//   * Surface_mesh_simplification::edge_collapse as a no-op returning 0 (Deformation::Simplification
//     is not on the path: its only call, Processor.cpp, is commented out).
#ifndef MVS_COMPAT_CGAL_STANDIN_H
#define MVS_COMPAT_CGAL_STANDIN_H
#include <cmath>
#include <cstddef>
#include <map>
#include <set>
#include <stdexcept>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

namespace CGAL {

struct Null_vector {};
static const Null_vector NULL_VECTOR = Null_vector();

template <class FT>
struct Vector_3T {
  FT c[3];
  Vector_3T() { c[0] = c[1] = c[2] = FT(0); }
  Vector_3T(const Null_vector&) { c[0] = c[1] = c[2] = FT(0); }
  Vector_3T(FT x, FT y, FT z) { c[0] = x; c[1] = y; c[2] = z; }
  FT operator[](int i) const { return c[i]; }
  FT x() const { return c[0]; }
  FT y() const { return c[1]; }
  FT z() const { return c[2]; }
  Vector_3T operator+(const Vector_3T& o) const { return Vector_3T(c[0] + o.c[0], c[1] + o.c[1], c[2] + o.c[2]); }
  Vector_3T operator-(const Vector_3T& o) const { return Vector_3T(c[0] - o.c[0], c[1] - o.c[1], c[2] - o.c[2]); }
  Vector_3T operator/(FT s) const { return Vector_3T(c[0] / s, c[1] / s, c[2] / s); }
  Vector_3T operator*(FT s) const { return Vector_3T(c[0] * s, c[1] * s, c[2] * s); }
  FT operator*(const Vector_3T& o) const { return c[0] * o.c[0] + c[1] * o.c[1] + c[2] * o.c[2]; }  // scalar product
};

template <class FT>
struct Point_3T {
  FT c[3];
  Point_3T() { c[0] = c[1] = c[2] = FT(0); }
  Point_3T(FT x, FT y, FT z) { c[0] = x; c[1] = y; c[2] = z; }
  FT operator[](int i) const { return c[i]; }
  FT x() const { return c[0]; }
  FT y() const { return c[1]; }
  FT z() const { return c[2]; }
  Vector_3T<FT> operator-(const Point_3T& o) const {
    return Vector_3T<FT>(c[0] - o.c[0], c[1] - o.c[1], c[2] - o.c[2]);
  }
};

template <class FT>
Vector_3T<FT> cross_product(const Vector_3T<FT>& a, const Vector_3T<FT>& b) {
  return Vector_3T<FT>(a.c[1] * b.c[2] - a.c[2] * b.c[1], a.c[2] * b.c[0] - a.c[0] * b.c[2],
                       a.c[0] * b.c[1] - a.c[1] * b.c[0]);
}

template <class FT_>
struct Simple_cartesian {
  typedef FT_ FT;
  typedef Point_3T<FT_> Point_3;
  typedef Vector_3T<FT_> Vector_3;
};
template <class K>
struct Polyhedron_traits_with_normals_3 : K {};

class Assertion_exception : public std::logic_error {
 public:
  Assertion_exception(const std::string&, const std::string&, const std::string&, int,
                      const std::string& msg)
      : std::logic_error(msg) {}
};

struct Polyhedron_items_with_id_3 {};
struct Polyhedron_items_3 {};

template <class HDS>
class Modifier_base {
 public:
  virtual void operator()(HDS& hds) = 0;
  virtual ~Modifier_base() {}
};

// ---- circulators: a cyclic walk over a list of halfedges
template <class H>
class Ring_circulator {
 public:
  Ring_circulator() : ring_(nullptr), k_(0) {}
  Ring_circulator(const std::vector<H*>* ring, size_t k) : ring_(ring), k_(k) {}
  H& operator*() const { return *(*ring_)[k_]; }
  H* operator->() const { return (*ring_)[k_]; }
  Ring_circulator& operator++() {
    k_ = (k_ + 1) % ring_->size();
    return *this;
  }
  bool operator==(const Ring_circulator& o) const { return ring_ == o.ring_ && k_ == o.k_; }
  bool operator!=(const Ring_circulator& o) const { return !(*this == o); }
  bool empty() const { return !ring_ || ring_->empty(); }

 private:
  const std::vector<H*>* ring_;
  size_t k_;
};
template <class H>
bool is_empty_range(const Ring_circulator<H>& a, const Ring_circulator<H>&) {
  return a.empty();
}
template <class It>
bool is_empty_range(const It& a, const It& b) {
  return a == b;
}
#define CGAL_For_all(ic1, ic2)                                                         \
  for (bool _circ_loop_flag = !::CGAL::is_empty_range(ic1, ic2); _circ_loop_flag; \
       _circ_loop_flag = ((++ic1) != (ic2)))

template <class Kernel, class Items = Polyhedron_items_3>
class Polyhedron_3 {
 public:
  typedef Kernel Traits;
  typedef typename Kernel::Point_3 Point_3;
  struct Vertex;
  struct Halfedge;
  struct Facet;
  typedef Vertex* Vertex_handle;
  typedef const Vertex* Vertex_const_handle;
  typedef Halfedge* Halfedge_handle;
  typedef const Halfedge* Halfedge_const_handle;
  typedef Facet* Facet_handle;
  typedef Vertex* Vertex_iterator;
  typedef const Vertex* Vertex_const_iterator;
  typedef Facet* Facet_iterator;
  typedef const Facet* Facet_const_iterator;
  typedef Ring_circulator<const Halfedge> Halfedge_around_vertex_const_circulator;
  typedef Ring_circulator<const Halfedge> Halfedge_around_facet_const_circulator;

  struct Halfedge {
    Vertex* v;   // the vertex the halfedge points to
    Halfedge* nxt;
    Facet* f;
    const Vertex* vertex() const { return v; }
    Vertex* vertex() { return v; }
    const Halfedge* next() const { return nxt; }
    Halfedge* next() { return nxt; }
    const Facet* facet() const { return f; }
    Facet* facet() { return f; }
    bool is_border() const { return f == nullptr; }
  };
  struct Vertex {
    typedef typename Kernel::Point_3 Point;
    Point pt;
    std::size_t id_;
    std::vector<const Halfedge*> in;  // the facets' halfedges that point to this vertex
    Point& point() { return pt; }
    const Point& point() const { return pt; }
    std::size_t& id() { return id_; }
    std::size_t id() const { return id_; }
    Halfedge_around_vertex_const_circulator vertex_begin() const {
      return Halfedge_around_vertex_const_circulator(&in, 0);
    }
  };
  struct Facet {
    std::vector<const Halfedge*> ring;  // its three halfedges, in order
    const Halfedge* halfedge() const { return ring[0]; }
    Halfedge_around_facet_const_circulator facet_begin() const {
      return Halfedge_around_facet_const_circulator(&ring, 0);
    }
  };
  struct HalfedgeDS {
    typedef typename Polyhedron_3::Vertex Vertex;
    typedef typename Polyhedron_3::Halfedge Halfedge;
    typedef typename Polyhedron_3::Facet Face;
    std::vector<Vertex> vertices;
    std::vector<Halfedge> halfedges;
    std::vector<Facet> facets;
    bool valid = true;
  };
  class Point_iterator {
   public:
    explicit Point_iterator(Vertex* p = nullptr) : p_(p) {}
    Point_3& operator*() const { return p_->pt; }
    Point_3* operator->() const { return &p_->pt; }
    Point_iterator& operator++() {
      ++p_;
      return *this;
    }
    bool operator==(const Point_iterator& o) const { return p_ == o.p_; }
    bool operator!=(const Point_iterator& o) const { return p_ != o.p_; }

   private:
    Vertex* p_;
  };

  void delegate(Modifier_base<HalfedgeDS>& m) { m(hds_); }
  bool is_valid() const { return hds_.valid; }
  std::size_t size_of_vertices() const { return hds_.vertices.size(); }
  std::size_t size_of_facets() const { return hds_.facets.size(); }
  std::size_t size_of_halfedges() const {
    std::set<std::pair<std::size_t, std::size_t> > e;
    for (const Facet& f : hds_.facets)
      for (int k = 0; k < 3; ++k) {
        std::size_t a = f.ring[(k + 2) % 3]->v->id_, b = f.ring[k]->v->id_;
        e.insert(std::make_pair(a < b ? a : b, a < b ? b : a));
      }
    return 2 * e.size();
  }
  Vertex_iterator vertices_begin() { return hds_.vertices.data(); }
  Vertex_iterator vertices_end() { return hds_.vertices.data() + hds_.vertices.size(); }
  Vertex_const_iterator vertices_begin() const { return hds_.vertices.data(); }
  Vertex_const_iterator vertices_end() const { return hds_.vertices.data() + hds_.vertices.size(); }
  Facet_iterator facets_begin() { return hds_.facets.data(); }
  Facet_iterator facets_end() { return hds_.facets.data() + hds_.facets.size(); }
  Facet_const_iterator facets_begin() const { return hds_.facets.data(); }
  Facet_const_iterator facets_end() const { return hds_.facets.data() + hds_.facets.size(); }
  Point_iterator points_begin() { return Point_iterator(vertices_begin()); }
  Point_iterator points_end() { return Point_iterator(vertices_end()); }

  // a copy must re-seat the internal pointers (Deformation::Polyhedron() returns by value)
  Polyhedron_3() {}
  Polyhedron_3(const Polyhedron_3& o) { copy_from(o); }
  Polyhedron_3& operator=(const Polyhedron_3& o) {
    if (this != &o) copy_from(o);
    return *this;
  }

 private:
  void copy_from(const Polyhedron_3& o) {
    hds_ = o.hds_;
    const Vertex* v0 = o.hds_.vertices.data();
    const Halfedge* h0 = o.hds_.halfedges.data();
    const Facet* f0 = o.hds_.facets.data();
    for (Halfedge& h : hds_.halfedges) {
      h.v = hds_.vertices.data() + (h.v - v0);
      h.nxt = hds_.halfedges.data() + (h.nxt - h0);
      h.f = h.f ? hds_.facets.data() + (h.f - f0) : nullptr;
    }
    for (Vertex& v : hds_.vertices)
      for (const Halfedge*& h : v.in) h = hds_.halfedges.data() + (h - h0);
    for (Facet& f : hds_.facets)
      for (const Halfedge*& h : f.ring) h = hds_.halfedges.data() + (h - h0);
  }
  HalfedgeDS hds_;
};

template <class HDS>
class Polyhedron_incremental_builder_3 {
 public:
  Polyhedron_incremental_builder_3(HDS& hds, bool = false) : hds_(hds) {}
  void begin_surface(std::size_t nv, std::size_t nf, std::size_t = 0) {
    hds_.vertices.clear();
    hds_.halfedges.clear();
    hds_.facets.clear();
    hds_.vertices.reserve(nv);
    hds_.halfedges.reserve(3 * nf);
    hds_.facets.reserve(nf);
    nv_ = nv;
    nf_ = nf;
    hds_.valid = true;
  }
  template <class P>
  void add_vertex(const P& p) {
    typename HDS::Vertex v;
    v.pt = p;
    v.id_ = hds_.vertices.size();
    hds_.vertices.push_back(v);
  }
  void begin_facet() { cur_.clear(); }
  void add_vertex_to_facet(std::size_t i) { cur_.push_back(i); }
  void end_facet() {
    if (cur_.size() != 3 || hds_.facets.size() >= nf_) {
      hds_.valid = false;
      return;
    }
    for (std::size_t i : cur_)
      if (i >= hds_.vertices.size()) {
        hds_.valid = false;
        return;
      }
    // an oriented 2-manifold uses every directed edge at most once
    for (int k = 0; k < 3; ++k)
      if (!edges_.insert(std::make_pair(cur_[k], cur_[(k + 1) % 3])).second) hds_.valid = false;
    std::size_t h0 = hds_.halfedges.size();
    hds_.facets.push_back(typename HDS::Face());
    typename HDS::Face* f = &hds_.facets.back();
    for (int k = 0; k < 3; ++k) {
      typename HDS::Halfedge h;
      h.v = &hds_.vertices[cur_[(k + 1) % 3]];  // halfedge k runs from corner k to corner k + 1
      h.nxt = nullptr;
      h.f = f;
      hds_.halfedges.push_back(h);
    }
    for (int k = 0; k < 3; ++k) {
      typename HDS::Halfedge* h = &hds_.halfedges[h0 + k];
      h->nxt = &hds_.halfedges[h0 + (k + 1) % 3];
    }
    // CGAL: f->halfedge()->vertex(), ->next()->vertex(), ->next()->next()->vertex() walk the facet
    // in the order its vertices were added, starting at the first one
    for (int k = 0; k < 3; ++k) {
      const typename HDS::Halfedge* h = &hds_.halfedges[h0 + (k + 2) % 3];
      f->ring.push_back(h);
      h->v->in.push_back(h);
    }
  }
  void end_surface() {}
  bool error() const { return !hds_.valid; }

 private:
  HDS& hds_;
  std::size_t nv_ = 0, nf_ = 0;
  std::vector<std::size_t> cur_;
  std::set<std::pair<std::size_t, std::size_t> > edges_;
};

template <class Poly>
void set_halfedgeds_items_id(Poly& p) {
  std::size_t i = 0;
  for (typename Poly::Vertex_iterator it = p.vertices_begin(); it != p.vertices_end(); ++it) it->id() = i++;
}

// ---- what boost::graph_traits<Polyhedron>::vertex_iterator is: dereferences to a vertex_descriptor
template <class V>
class Descriptor_iterator {
 public:
  explicit Descriptor_iterator(V* p = nullptr) : p_(p) {}
  V* operator*() const { return p_; }
  Descriptor_iterator& operator++() {
    ++p_;
    return *this;
  }
  Descriptor_iterator operator+(std::ptrdiff_t n) const { return Descriptor_iterator(p_ + n); }
  bool operator==(const Descriptor_iterator& o) const { return p_ == o.p_; }
  bool operator!=(const Descriptor_iterator& o) const { return p_ != o.p_; }

 private:
  V* p_;
};
template <class K, class I>
std::pair<Descriptor_iterator<typename Polyhedron_3<K, I>::Vertex>,
          Descriptor_iterator<typename Polyhedron_3<K, I>::Vertex> >
vertices(Polyhedron_3<K, I>& p) {
  typedef Descriptor_iterator<typename Polyhedron_3<K, I>::Vertex> It;
  return std::make_pair(It(p.vertices_begin()), It(p.vertices_end()));
}
namespace cpp11 {
template <class V>
Descriptor_iterator<V> next(const Descriptor_iterator<V>& it, std::ptrdiff_t n = 1) {
  return it + n;
}
}  // namespace cpp11

// ---- Surface_mesh_deformation: records what the reference asks for
namespace standin {
struct Constraint {
  std::size_t vertex;
  double target[3];
};
struct Deformation_log {
  std::vector<Constraint> constraints;  // in the order of set_target_position
  std::size_t roi = 0;
  int deform_calls = 0;
  unsigned iterations = 0;
  double tolerance = 0.0;
};
inline Deformation_log& deformation_log() {
  static Deformation_log log;
  return log;
}
}  // namespace standin

template <class Poly>
class Surface_mesh_deformation {
 public:
  typedef typename Poly::Point_3 Point;
  typedef typename Poly::Vertex* vertex_descriptor;
  explicit Surface_mesh_deformation(Poly&) { standin::deformation_log() = standin::Deformation_log(); }
  template <class It>
  void insert_roi_vertices(It b, It e) {
    for (; b != e; ++b) ++standin::deformation_log().roi;
  }
  bool insert_roi_vertex(vertex_descriptor) {
    ++standin::deformation_log().roi;
    return true;
  }
  bool insert_control_vertex(vertex_descriptor v) {
    if (!control_.insert(v->id()).second) return false;
    return true;
  }
  void set_target_position(vertex_descriptor v, const Point& p) {
    if (!control_.count(v->id())) return;  // CGAL ignores vertices that are not control vertices
    standin::Constraint c;
    c.vertex = v->id();
    for (int k = 0; k < 3; ++k) c.target[k] = p[k];
    standin::deformation_log().constraints.push_back(c);
  }
  bool preprocess() { return true; }
  void deform() { ++standin::deformation_log().deform_calls; }
  void deform(unsigned iterations, double tolerance) {
    standin::Deformation_log& l = standin::deformation_log();
    ++l.deform_calls;
    l.iterations = iterations;
    l.tolerance = tolerance;
  }
  void overwrite_initial_geometry() {}

 private:
  std::set<std::size_t> control_;
};

// ---- Surface_mesh_simplification: names only (Deformation::Simplification is never called)
enum vertex_external_index_t { vertex_external_index };
enum halfedge_external_index_t { halfedge_external_index };
template <class Tag, class Poly>
int get(Tag, Poly&) {
  return 0;
}
struct Named_params {
  template <class T>
  Named_params halfedge_index_map(const T&) const { return *this; }
  template <class T>
  Named_params get_cost(const T&) const { return *this; }
  template <class T>
  Named_params get_placement(const T&) const { return *this; }
};
template <class T>
Named_params vertex_index_map(const T&) {
  return Named_params();
}
namespace Surface_mesh_simplification {
template <class P>
struct Count_stop_predicate {
  explicit Count_stop_predicate(int) {}
};
template <class P>
struct Edge_length_cost {};
template <class P>
struct Midpoint_placement {};
template <class P, class Stop, class Params>
int edge_collapse(P&, const Stop&, const Params&) {
  return 0;
}
}  // namespace Surface_mesh_simplification

}  // namespace CGAL

namespace boost {
template <class G>
struct graph_traits;
template <class K, class I>
struct graph_traits<CGAL::Polyhedron_3<K, I> > {
  typedef typename CGAL::Polyhedron_3<K, I>::Vertex* vertex_descriptor;
  typedef CGAL::Descriptor_iterator<typename CGAL::Polyhedron_3<K, I>::Vertex> vertex_iterator;
};
template <class A, class B>
std::tuple<A&, B&> tie(A& a, B& b) {
  return std::tuple<A&, B&>(a, b);
}
}  // namespace boost

#endif
