// A Processor-style caller written against the C++ shim (multiviewstitch_b200/shim/mvs_shim.h):
// the reference's call sequences for the hot path, on synthetic data with planted answers.
// Exit codes: 0 ok, 3 no CUDA device (the library has no CPU fallback), 1 wrong result.
#include <cmath>
#include <cstdio>
#include <random>

#include "../../multiviewstitch_b200/shim/mvs_shim.h"

using namespace mvs;

static Matrix3d axis_angle(double ax, double ay, double az, double ang) {
  double l = std::sqrt(ax * ax + ay * ay + az * az);
  ax /= l; ay /= l; az /= l;
  double c = std::cos(ang), s = std::sin(ang);
  Matrix3d R;
  R(0, 0) = c + ax * ax * (1 - c);      R(0, 1) = ax * ay * (1 - c) - az * s; R(0, 2) = ax * az * (1 - c) + ay * s;
  R(1, 0) = ay * ax * (1 - c) + az * s; R(1, 1) = c + ay * ay * (1 - c);      R(1, 2) = ay * az * (1 - c) - ax * s;
  R(2, 0) = az * ax * (1 - c) - ay * s; R(2, 1) = az * ay * (1 - c) + ax * s; R(2, 2) = c + az * az * (1 - c);
  return R;
}

int main() {
  try {
    int ndev = 0;
    if (mvs_device_count(&ndev) != MVS_OK || ndev == 0) {
      // prove the product path fails loudly instead of falling back
      try {
        SRTSolver s(10);
        std::printf("unexpected: SRTSolver constructed without a device\n");
        return 1;
      } catch (const Error& e) {
        std::printf("no device: %s (status %d)\n", e.what(), e.status);
        return 3;
      }
    }
    std::mt19937 rng(5);
    std::normal_distribution<double> N(0.0, 1.0);
    // ---- SRTSolver: Processor.cpp:814-818 call sequence, planted similarity
    const double s0 = 1.37;
    Matrix3d R0 = axis_angle(1, 2, 3, 0.4);
    Vector3d t0(0.3, -0.2, 0.5);
    std::vector<std::pair<Vector3d, Vector3d>> matches;
    for (int i = 0; i < 5000; ++i) {
      Vector3d p(N(rng), N(rng), 5.0 + N(rng)), q;
      for (int a = 0; a < 3; ++a) q[a] = s0 * (R0(a, 0) * p[0] + R0(a, 1) * p[1] + R0(a, 2) * p[2]) + t0[a];
      matches.push_back({p, q});
    }
    Camera cam;
    Matrix3d K;
    K(0, 0) = K(1, 1) = 525; K(0, 2) = 319.5; K(1, 2) = 239.5;
    cam.SetK(K); cam.SetRT(Matrix3d(), Vector3d()); cam.SetW(640); cam.SetH(480);
    SRTSolver solver(200);
    solver.SetInput(matches, cam, cam);
    solver.SetPrintFlag(false);
    double scale; Matrix3d R; Vector3d t;
    solver.EstimateTransform(scale, R, t);
    double err = std::fabs(scale - s0);
    for (int a = 0; a < 3; ++a) {
      err = std::fmax(err, std::fabs(t[a] - t0[a]));
      for (int b = 0; b < 3; ++b) err = std::fmax(err, std::fabs(R(a, b) - R0(a, b)));
    }
    std::printf("SRT max abs error vs planted: %.3e\n", err);
    if (err > 1e-9) return 1;
    // ---- Deformation: Processor.cpp:1135-1137 call sequence on an icosahedron-ish blob
    std::vector<Vector3d> pts, nrm;
    std::vector<int> facets;
    const int nu = 48, nv = 32;
    for (int i = 1; i < nv; ++i)
      for (int j = 0; j < nu; ++j) {
        double th = M_PI * i / nv, ph = 2 * M_PI * j / nu;
        pts.push_back(Vector3d(std::sin(th) * std::cos(ph), std::sin(th) * std::sin(ph), std::cos(th)));
      }
    nrm = pts;
    for (int i = 0; i + 1 < nv - 1; ++i)
      for (int j = 0; j < nu; ++j) {
        int a = i * nu + j, b = i * nu + (j + 1) % nu, c = (i + 1) * nu + j, d = (i + 1) * nu + (j + 1) % nu;
        facets.insert(facets.end(), {a, c, b, b, c, d});
      }
    std::vector<Vector3d> tgt = pts, tn = nrm;
    for (auto& p : tgt)
      for (int a = 0; a < 3; ++a) p[a] *= 1.06;
    Deformation deform(pts, nrm, facets);
    deform.gn.gn_iters = 5;
    deform.gn.pcg_iters = 30;
    deform.Deform(tgt, tn, 100.0, 100.0);
    auto out = deform.Vertices();
    double rmean = 0;
    for (auto& p : out) rmean += std::sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
    rmean /= out.size();
    std::printf("deformed mean radius %.4f (template 1.00, target 1.06)\n", rmean);
    if (!(rmean > 1.03 && rmean < 1.07)) return 1;
    // the same call with the solver the reference runs through CGAL (as-rigid-as-possible)
    Deformation deform2(pts, nrm, facets);
    deform2.solver = Deformation::ARAP;
    deform2.Deform(tgt, tn, 100.0, 100.0);
    auto out2 = deform2.Vertices();
    double rmean2 = 0;
    for (auto& p : out2) rmean2 += std::sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
    rmean2 /= out2.size();
    std::printf("ARAP deformed mean radius %.4f\n", rmean2);
    if (!(rmean2 > 1.03 && rmean2 < 1.07)) return 1;
    // ---- Model2Depth: fronto-parallel plane renders 1/z
    std::vector<Vector3d> plane = {Vector3d(-10, -10, 4), Vector3d(10, -10, 4), Vector3d(10, 10, 4), Vector3d(-10, 10, 4)};
    Model2Depth::SetInput({plane}, {{}}, {{0, 1, 2, 0, 2, 3}}, {{cam}}, {""});
    Model2Depth::Run();
    const auto& img = Model2Depth::Images()[0];
    for (float v : img)
      if (v != 0.25f) return 1;
    std::printf("Model2Depth plane ok (%zu px = 0.25)\n", img.size());
    // ---- StitchPairs (Processor.cpp:629-826 pair loop): two views of a sphere 5 units away, view B's
    // frame = planted similarity of view A's; lattice matches; both pairs must recover (s, R, t)
    {
      const int W = 640, H = 480;
      auto render_sphere = [&](const Matrix3d& Rw, const Vector3d& tw, double sw, const Camera& c) {
        // sphere of radius sw centred at sw * Rw * (0,0,5) + tw in the camera's world frame, camera
        // [I | 0]: inverse depth of the front intersection per pixel
        std::vector<float> d((size_t)W * H, 0.f);
        double ctr[3];
        for (int a = 0; a < 3; ++a) ctr[a] = sw * Rw(a, 2) * 5.0 + tw[a];
        for (int v = 0; v < H; ++v)
          for (int u = 0; u < W; ++u) {
            double dir[3] = {(u - 319.5) / 525.0, (v - 239.5) / 525.0, 1.0};
            double a2 = dir[0] * dir[0] + dir[1] * dir[1] + 1.0;
            double b = dir[0] * ctr[0] + dir[1] * ctr[1] + ctr[2];
            double c2 = ctr[0] * ctr[0] + ctr[1] * ctr[1] + ctr[2] * ctr[2] - sw * sw;
            double disc = b * b - a2 * c2;
            if (disc <= 0) continue;
            double z = (b - std::sqrt(disc)) / a2;
            if (z > 0) d[(size_t)v * W + u] = (float)(1.0 / z);
          }
        (void)c;
        return d;
      };
      Matrix3d I3;
      I3(0, 0) = I3(1, 1) = I3(2, 2) = 1.0;
      Camera ca;
      ca.SetK(K); ca.SetRT(I3, Vector3d()); ca.SetW(W); ca.SetH(H);
      const double sp = 1.1;
      Matrix3d Rp = axis_angle(0, 1, 0, 0.05);
      Vector3d tp(0.05, -0.03, 0.2);
      std::vector<float> da = render_sphere(I3, Vector3d(), 1.0, ca);
      std::vector<float> db = render_sphere(Rp, tp, sp, ca);
      ViewPair vp;
      vp.depth_a = &da; vp.depth_b = &db; vp.cam_a = ca; vp.cam_b = ca;
      for (int v = 10; v < H; v += 20)
        for (int u = 10; u < W; u += 20) {
          float d = da[(size_t)v * W + u];
          if (!(d >= 0.0025f && d <= 0.3f)) continue;
          double z = 1.0 / d, x = (u - 319.5) / 525.0 * z, y = (v - 239.5) / 525.0 * z, q[3];
          for (int a = 0; a < 3; ++a) q[a] = sp * (Rp(a, 0) * x + Rp(a, 1) * y + Rp(a, 2) * z) + tp[a];
          int u2 = (int)(525.0 * q[0] / q[2] + 319.5 + 0.5), v2 = (int)(525.0 * q[1] / q[2] + 239.5 + 0.5);
          if (u2 < 0 || v2 < 0 || u2 >= W || v2 >= H) continue;
          float d2 = db[(size_t)v2 * W + u2];
          if (!(d2 >= 0.0025f && d2 <= 0.3f)) continue;
          vp.matches.insert(vp.matches.end(), {u, v, u2, v2});
        }
      std::vector<ViewPair> pairs(2, vp);
      pairs[1].seed = 7;
      mvs_pair_params pp;
      mvs_pair_default_params(&pp);
      pp.n_nodes = 128;
      pp.workers = 2;
      auto res = StitchPairs(pairs, &pp);
      for (const auto& r : res) {
        if (std::fabs(r.scale - sp) > 0.03 * sp || r.points.empty() || r.deformed.size() != r.points.size())
          return 1;
        size_t covered = 0;
        for (float v : r.render) covered += v > 0.f;
        if (covered < 1000) return 1;
      }
      std::printf("StitchPairs ok: %zu matches, scale %.4f (planted %.2f), %zu vertices\n",
                  vp.matches.size() / 4, res[0].scale, sp, res[0].points.size());
    }
    return 0;
  } catch (const Error& e) {
    std::printf("error: %s\n", e.what());
    return e.status == MVS_ERR_NO_DEVICE ? 3 : 1;
  }
}
