// stitch.cu — mvs_stitch_pairs: the hot path of a list of independent view pairs in ONE library
// call (the reference's independent `for k` pair loop, Processor/Processor.cpp:
// 629-826, and per pair the sequence :751-765 RemoveOutliers, :806-823 SRTSolver, :1025-1026 apply,
// :1130-1137 Deformation, :1190-1191 Model2Depth).
//
// Host side only: a pool of worker threads, each with its OWN context (own stream, side stream,
// stream-ordered allocation pool, pinned staging), works through the jobs k, k + W, k + 2W, ...
// Pairs are independent, so one pair's host glue and stream synchronisations overlap the other
// workers' kernels; nothing here is Python, so no interpreter lock serialises the glue.  Every
// worker context lives as long as the parent context (created once, reused by later calls).
//
// Per pair, in the reference's call order:
//   matched pixels --unproject--> p, q                      [Camera/Camera.cpp:40-44,61-67]
//   RemoveOutliers (rounds x RANSAC + symmetric inlier test) [Processor/Processor.cpp:177-269]
//   SRTSolver::EstimateTransform + ResidualError             [Processor/Processor.cpp:814-818]
//   Depth2Model mesh of view 1, s R p + t into frame 2       [Depth2Model.cpp:7-82, Processor.cpp:1025]
//   Deformation onto the unprojected view-2 points (GN)      [Processor/Processor.cpp:1135-1137]
//   Model2Depth of the deformed mesh into camera 2           [Processor/Processor.cpp:1190-1191]
#include <atomic>
#include <chrono>
#include <mutex>
#include <thread>
#include <stdlib.h>

#include <vector>

#include "deform.cuh"

namespace mvs {

// host scratch of one worker, kept between pairs and calls: a fresh 7 MB std::vector per pair costs a
// zero fill plus an mmap/munmap pair, and with a dozen workers those serialise on the process's
// address-space lock
struct PairScratch {
  std::vector<int32_t> uv, okf, keep;
  std::vector<double> p, q, pk, qk;
  // device bump arena of this worker (common.cuh AllocStream): every temporary of a pair
  char* arena = nullptr;
  size_t arena_cap = 0;
};

constexpr size_t kArenaStart = (size_t)96 << 20;

// every device buffer created by this thread until arena_off() comes from the worker's arena
static int arena_on(mvs_ctx* ctx, PairScratch& S) {
  if (getenv("MVS_NO_ARENA")) return MVS_OK;
  AllocStream& as = alloc_stream();
  size_t start = kArenaStart;
  if (const char* env = getenv("MVS_ARENA_MB")) {  // test hook: a small arena exercises the fallback
    long mb = atol(env);
    if (mb > 0) start = (size_t)mb << 20;
  }
  size_t want = std::max(S.arena_cap, start);
  if (as.arena_missed > 0) want = std::max(want, S.arena_cap + 2 * as.arena_missed + (start >> 1));
  as.arena_missed = 0;
  if (want > S.arena_cap) {
    if (S.arena) cudaFree(S.arena);
    S.arena = nullptr;
    S.arena_cap = 0;
    if (cudaMalloc((void**)&S.arena, want) != cudaSuccess) {
      cudaGetLastError();
      return MVS_OK;  // no arena: the stream-ordered pool serves the pair
    }
    S.arena_cap = want;
  }
  as.arena = S.arena;
  as.arena_cap = S.arena_cap;
  as.arena_off = 0;
  return MVS_OK;
}
static void arena_off() {
  AllocStream& as = alloc_stream();
  as.arena = nullptr;
  as.arena_cap = as.arena_off = 0;
}

struct StitchPool {
  std::mutex mu;
  std::vector<mvs_ctx*> workers;
  std::vector<PairScratch*> scratch;
};

template <class T>
static inline void ensure(std::vector<T>& v, size_t n) {
  if (v.size() < n) v.resize(n);
}

static std::mutex g_pools_mu;
static std::map<mvs_ctx*, StitchPool*> g_pools;

static StitchPool* pool_for(mvs_ctx* parent) {
  std::lock_guard<std::mutex> lk(g_pools_mu);
  auto it = g_pools.find(parent);
  if (it != g_pools.end()) return it->second;
  StitchPool* p = new StitchPool();
  g_pools[parent] = p;
  return p;
}

// called by mvs_ctx_destroy of the parent
void stitch_pool_release(mvs_ctx* parent) {
  StitchPool* p = nullptr;
  {
    std::lock_guard<std::mutex> lk(g_pools_mu);
    auto it = g_pools.find(parent);
    if (it == g_pools.end()) return;
    p = it->second;
    g_pools.erase(it);
  }
  for (mvs_ctx* w : p->workers) mvs_ctx_destroy(w);
  for (PairScratch* sc : p->scratch) {
    if (sc->arena) cudaFree(sc->arena);
    delete sc;
  }
  delete p;
}

// wall-clock stage accounting of a worker (MVS_STITCH_TRACE=1 prints the sums per call to stderr)
enum { ST_UNPROJ, ST_OUTLIERS, ST_SRT, ST_D2M, ST_APPLY, ST_CREATE, ST_FPS, ST_TARGET, ST_SOLVE,
       ST_GET, ST_RENDER, ST_COUNT };
static const char* kStageNames[ST_COUNT] = {"unproject", "remove_outliers", "srt", "depth2model",
                                            "apply_srt", "deform_create", "sample_fps", "set_target",
                                            "solve(graph+gn)", "get_vertices", "model2depth"};
struct StageClock {
  double* acc;  // [ST_COUNT] milliseconds, then [ST_COUNT] kernel launches
  mvs_ctx* ctx;
  int64_t l0;
  std::chrono::steady_clock::time_point t0;
  StageClock(double* a, mvs_ctx* c) : acc(a), ctx(c), l0(c->launch_count), t0(std::chrono::steady_clock::now()) {}
  void lap(int stage) {
    auto t1 = std::chrono::steady_clock::now();
    if (acc) {
      acc[stage] += std::chrono::duration<double, std::milli>(t1 - t0).count();
      acc[ST_COUNT + stage] += (double)(ctx->launch_count - l0);
    }
    l0 = ctx->launch_count;
    t0 = t1;
  }
};

static int run_pair(mvs_ctx* ctx, PairScratch& S, const mvs_pair_job& job, const mvs_pair_params& prm,
                    mvs_pair_result& out, double* stage_ms) {
  StageClock clk(stage_ms, ctx);
  const mvs_camera* ca = job.cam_a;
  const mvs_camera* cb = job.cam_b;
  const int64_t nm = job.n_matches;
  const int64_t npx = (int64_t)ca->w * ca->h;
  out.status = MVS_OK;
  out.n_kept = out.n_template = out.n_facets = out.n_target = 0;
  // ---- matched pixels -> 3-D matches
  std::vector<int32_t>&uv = S.uv, &okf = S.okf, &keep = S.keep;
  std::vector<double>&p = S.p, &q = S.q, &pk = S.pk, &qk = S.qk;
  ensure(uv, (size_t)nm * 2);
  ensure(p, (size_t)nm * 3);
  ensure(q, (size_t)nm * 3);
  ensure(okf, (size_t)nm);
  ensure(keep, (size_t)nm);
  ensure(pk, (size_t)nm * 3);
  ensure(qk, (size_t)nm * 3);
  for (int64_t i = 0; i < nm; ++i) {
    uv[2 * i] = job.matches[4 * i];
    uv[2 * i + 1] = job.matches[4 * i + 1];
  }
  MVS_TRY(mvs_unproject_pixels(ctx, job.depth_a, ca, prm.min_dsp, prm.max_dsp, uv.data(), nm,
                               p.data(), okf.data()));
  for (int64_t i = 0; i < nm; ++i) {
    uv[2 * i] = job.matches[4 * i + 2];
    uv[2 * i + 1] = job.matches[4 * i + 3];
  }
  MVS_TRY(mvs_unproject_pixels(ctx, job.depth_b, cb, prm.min_dsp, prm.max_dsp, uv.data(), nm,
                               q.data(), okf.data()));
  clk.lap(ST_UNPROJ);
  // ---- RemoveOutliers
  int64_t nk = 0;
  mvs_outlier_params op = prm.outliers;
  op.seed = job.seed;
  MVS_TRY(mvs_remove_outliers(ctx, p.data(), q.data(), nm, ca, cb, &op, keep.data(), &nk, &out.err,
                              nullptr));
  out.n_kept = nk;
  clk.lap(ST_OUTLIERS);
  if (job.kept_idx) memcpy(job.kept_idx, keep.data(), sizeof(int32_t) * (size_t)nk);
  if (nk < 3) {
    set_error("stitch_pairs: fewer than 3 matches survive RemoveOutliers");
    return MVS_ERR_DEGENERATE;
  }
  // ---- closed-form SRT on the inliers + residual
  for (int64_t i = 0; i < nk; ++i) {
    memcpy(&pk[3 * i], &p[3 * (size_t)keep[i]], sizeof(double) * 3);
    memcpy(&qk[3 * i], &q[3 * (size_t)keep[i]], sizeof(double) * 3);
  }
  {
    mvs_srt* s = nullptr;
    MVS_TRY(mvs_srt_create(ctx, &s));
    int rc = mvs_srt_set_iteration_num(s, op.iter_num);
    if (rc == MVS_OK) rc = mvs_srt_set_print(s, 0);
    if (rc == MVS_OK) rc = mvs_srt_set_input(s, pk.data(), qk.data(), nk, ca, cb);
    if (rc == MVS_OK) rc = mvs_srt_estimate(s, &out.scale, out.R, out.t);
    if (rc == MVS_OK) rc = mvs_srt_residual(s, out.scale, out.R, out.t, &out.residual);
    mvs_srt_destroy(s);
    MVS_TRY(rc);
  }
  clk.lap(ST_SRT);
  // ---- template = view-1 mesh, moved into the frame of view 2: stays on the device from the depth
  // map to the deformation handle (the stage-by-stage form moves it through host vectors three times)
  int64_t np = 0, nf = 0;
  DevBuf<double> dxyz;
  DevBuf<int> dfac;
  MVS_TRY(depth2model_dev(ctx, job.depth_a, ca, prm.min_dsp, prm.max_dsp, prm.smooth_threshold, &dxyz,
                          nullptr, &dfac, &np, &nf));
  clk.lap(ST_D2M);
  out.n_template = np;
  out.n_facets = nf;
  if (np <= 0) {
    set_error("stitch_pairs: view 1 has no valid depth");
    return MVS_ERR_DEGENERATE;
  }
  MVS_TRY(apply_srt_dev_f64(ctx, dxyz.p, np, out.scale, out.R, out.t));
  if (job.template_xyz) MVS_TRY(d2h_async(ctx, job.template_xyz, dxyz.p, sizeof(double) * 3 * np));
  if (job.facets && nf > 0) MVS_TRY(d2h_async(ctx, job.facets, dfac.p, sizeof(int32_t) * 3 * nf));
  clk.lap(ST_APPLY);
  // ---- deformation onto view 2
  mvs_deform* d = nullptr;
  MVS_TRY(deform_create_dev_f64(ctx, dxyz.p, np, dfac.p, nf, &d));
  clk.lap(ST_CREATE);
  int rc = mvs_deform_sample_fps(d, std::min<int64_t>(prm.n_nodes, np), 0);
  clk.lap(ST_FPS);
  if (rc == MVS_OK)
    rc = mvs_deform_set_target_depth(d, job.depth_b, cb, prm.min_dsp, prm.max_dsp, &out.n_target);
  clk.lap(ST_TARGET);
  if (rc == MVS_OK) rc = mvs_deform_solve(d, &prm.gn);
  clk.lap(ST_SOLVE);
  if (rc == MVS_OK && job.deformed_xyz) rc = mvs_deform_get_vertices(d, job.deformed_xyz);
  clk.lap(ST_GET);
  // ---- re-projection of the deformed mesh into camera 2, from the device-resident vertices
  if (rc == MVS_OK && job.render) {
    const float4* cur = nullptr;
    deform_vertices_dev(d, &cur);
    DevBuf<float> img;
    rc = img.reserve((size_t)npx);
    if (rc == MVS_OK)
      rc = mvs_model2depth_dev(ctx, reinterpret_cast<const float*>(cur), np, dfac.p, nf, cb, 1, 0.f, 0.f,
                               img.p);
    if (rc == MVS_OK) rc = d2h_sync(ctx, job.render, img.p, sizeof(float) * npx);
  }
  mvs_deform_destroy(d);
  MVS_TRY(rc);
  clk.lap(ST_RENDER);
  return MVS_OK;
}

}  // namespace mvs

using namespace mvs;

extern "C" {

int mvs_pair_default_params(mvs_pair_params* p) {
  if (!p) return MVS_ERR_INVALID_ARG;
  p->min_dsp = 0.0025;         // config.txt:23
  p->max_dsp = 0.3;            // config.txt:24
  p->smooth_threshold = 0.12;  // config.txt:38
  mvs_outlier_default_params(&p->outliers);
  p->outliers.pixel_err = 60.0;    // config.txt
  p->outliers.adapt_ratio = 0.75;  // config.txt
  p->n_nodes = 512;
  mvs_gn_default_params(&p->gn);
  p->gn.gn_iters = 4;
  p->gn.corr_mode = 1;
  p->workers = 0;
  return MVS_OK;
}

int mvs_stitch_pairs(mvs_ctx* ctx, const mvs_pair_job* jobs, int64_t n_jobs,
                     const mvs_pair_params* params, mvs_pair_result* results) {
  if (!ctx) ctx = default_ctx();
  if (!ctx) return MVS_ERR_NO_DEVICE;
  MVS_TRY(ctx_ensure_device(ctx));
  if (n_jobs < 0 || (n_jobs > 0 && (!jobs || !results))) {
    set_error("stitch_pairs: bad argument");
    return MVS_ERR_INVALID_ARG;
  }
  mvs_pair_params def;
  if (!params) {
    mvs_pair_default_params(&def);
    params = &def;
  }
  for (int64_t i = 0; i < n_jobs; ++i) {
    const mvs_pair_job& j = jobs[i];
    if (!j.depth_a || !j.depth_b || !j.cam_a || !j.cam_b || j.n_matches < 0 ||
        (j.n_matches > 0 && !j.matches) || j.cam_a->w <= 1 || j.cam_a->h <= 1 ||
        j.cam_b->w != j.cam_a->w || j.cam_b->h != j.cam_a->h) {
      set_error("stitch_pairs: job %lld is malformed", (long long)i);
      return MVS_ERR_INVALID_ARG;
    }
  }
  if (n_jobs == 0) return MVS_OK;
  int want = params->workers > 0 ? params->workers
                                 : (int)std::min<unsigned>(8u, std::max(1u, std::thread::hardware_concurrency() / 2));
  int W = (int)std::min<int64_t>(std::max(1, want), n_jobs);
  StitchPool* pool = pool_for(ctx);
  std::lock_guard<std::mutex> lk(pool->mu);  // one stitch call per parent context at a time
  while ((int)pool->workers.size() < W) {
    mvs_ctx* w = nullptr;
    MVS_TRY(mvs_ctx_create(&w, ctx->device, nullptr));
    pool->workers.push_back(w);
    pool->scratch.push_back(new PairScratch());
  }
  // the workers start after everything already queued on the parent's stream
  MVS_CUDA(cudaStreamSynchronize(ctx->stream));
  std::atomic<int64_t> next{0};
  std::vector<std::string> errs((size_t)W);
  for (int64_t i = 0; i < n_jobs; ++i) results[i].status = MVS_ERR_NO_DEVICE;  // until a worker ran it
  std::vector<int64_t> launches((size_t)W, 0);
  const char* trace_env = getenv("MVS_STITCH_TRACE");
  const bool trace = trace_env && trace_env[0] == '1';
  std::vector<double> stage_ms(trace ? (size_t)W * 2 * ST_COUNT : 0, 0.0);
  auto body = [&](int k) {
    mvs_ctx* w = pool->workers[k];
    if (ctx_ensure_device(w) != MVS_OK) {  // the other workers take the jobs; if none can, the
      errs[k] = mvs_last_error();          // results keep the status set below
      return;
    }
    int64_t before = w->launch_count;
    for (;;) {
      int64_t i = next.fetch_add(1);
      if (i >= n_jobs) break;
      arena_on(w, *pool->scratch[k]);
      int rc = run_pair(w, *pool->scratch[k], jobs[i], *params, results[i], trace ? &stage_ms[(size_t)k * 2 * ST_COUNT] : nullptr);
      arena_off();  // run_pair destroyed everything it created; the stream is idle
      results[i].status = rc;
      if (rc != MVS_OK && errs[k].empty()) errs[k] = mvs_last_error();
    }
    cudaStreamSynchronize(w->stream);
    launches[k] = w->launch_count - before;
  };
  if (W == 1) {
    body(0);
  } else {
    std::vector<std::thread> th;
    for (int k = 0; k < W; ++k) th.emplace_back(body, k);
    for (auto& t : th) t.join();
  }
  MVS_TRY(ctx_ensure_device(ctx));
  if (trace) {
    fprintf(stderr, "[mvs_stitch_pairs] %lld pairs, %d workers, ms per pair (wall, per worker thread):",
            (long long)n_jobs, W);
    for (int s = 0; s < ST_COUNT; ++s) {
      double sum = 0, ln = 0;
      for (int k = 0; k < W; ++k) {
        sum += stage_ms[(size_t)k * 2 * ST_COUNT + s];
        ln += stage_ms[(size_t)k * 2 * ST_COUNT + ST_COUNT + s];
      }
      fprintf(stderr, " %s %.3f (%.0f launches)", kStageNames[s], sum / (double)n_jobs, ln / (double)n_jobs);
    }
    fprintf(stderr, "\n");
  }
  for (int k = 0; k < W; ++k) ctx->launch_count += launches[k];  // the call's launches, on the caller's context
  for (int64_t i = 0; i < n_jobs; ++i)
    if (results[i].status != MVS_OK) {
      for (int k = 0; k < W; ++k)
        if (!errs[k].empty()) {
          set_error("stitch_pairs: pair %lld failed: %s", (long long)i, errs[k].c_str());
          break;
        }
      return results[i].status;
    }
  return MVS_OK;
}

}  // extern "C"
