// ctx.cu — context lifetime, error string, profiling scopes, launch accounting.
#include <algorithm>
#include <stdarg.h>
#include <stdlib.h>

#include <mutex>

#include "common.cuh"

namespace mvs {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

AllocStream& alloc_stream() {
  static thread_local AllocStream as;
  return as;
}

int ctx_ensure_device(mvs_ctx* ctx) {
  if (!ctx) {
    set_error("null context");
    return MVS_ERR_INVALID_ARG;
  }
  MVS_CUDA(cudaSetDevice(ctx->device));
  AllocStream& as = alloc_stream();
  as.stream = ctx->stream;
  as.valid = true;
  return MVS_OK;
}

static mvs_ctx* g_default = nullptr;
static std::mutex g_default_mu;

mvs_ctx* default_ctx() {
  std::lock_guard<std::mutex> lk(g_default_mu);
  if (!g_default) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    mvs_ctx* c = nullptr;
    if (mvs_ctx_create(&c, dev, nullptr) != MVS_OK) return nullptr;
    g_default = c;
  }
  return g_default;
}

static cudaEvent_t take_event(mvs_ctx* ctx) {
  if (!ctx->event_pool.empty()) {
    cudaEvent_t e = ctx->event_pool.back();
    ctx->event_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
  return e;
}

ProfScope::ProfScope(mvs_ctx* c, const char* n, int64_t nlaunch, cudaStream_t on)
    : ctx(c), name(n), launches(nlaunch), st(on ? on : c->stream) {
  ctx->launch_count += nlaunch;
  if (ctx->prof_on) {
    start = take_event(ctx);
    stop = take_event(ctx);
    if (start) cudaEventRecord(start, st);
  }
}

ProfScope::~ProfScope() {
  if (ctx->prof_on && start && stop) {
    cudaEventRecord(stop, st);
    PendingProf p;
    p.name = name;
    p.start = start;
    p.stop = stop;
    p.launches = launches;
    ctx->pending.push_back(p);
  }
}

int ctx_side_fork(mvs_ctx* ctx) {
  if (!ctx->side) {
    MVS_CUDA(cudaStreamCreateWithFlags(&ctx->side, cudaStreamNonBlocking));
    MVS_CUDA(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    MVS_CUDA(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
  }
  MVS_CUDA(cudaEventRecord(ctx->ev_fork, ctx->stream));
  MVS_CUDA(cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0));
  return MVS_OK;
}

int ctx_side_join(mvs_ctx* ctx) {
  MVS_CUDA(cudaEventRecord(ctx->ev_join, ctx->side));
  MVS_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
  return MVS_OK;
}

int ctx_resolve_prof(mvs_ctx* ctx) {
  if (ctx->pending.empty()) return MVS_OK;
  MVS_CUDA(cudaStreamSynchronize(ctx->stream));
  if (ctx->side) MVS_CUDA(cudaStreamSynchronize(ctx->side));
  for (auto& p : ctx->pending) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, p.start, p.stop) == cudaSuccess) {
      ProfEntry& e = ctx->prof[p.name];
      e.total_ms += ms;
      e.launches += p.launches;
    }
    ctx->event_pool.push_back(p.start);
    ctx->event_pool.push_back(p.stop);
  }
  ctx->pending.clear();
  return MVS_OK;
}

int ensure_dyn_smem(const void* kernel, int device, int bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, int> limit;
  std::lock_guard<std::mutex> lk(mu);
  int& cur = limit[std::make_pair(kernel, device)];
  if (bytes <= cur || bytes <= 48 * 1024) return MVS_OK;
  MVS_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  cur = bytes;
  return MVS_OK;
}

int pinned_reserve(mvs_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->pinned_bytes) return MVS_OK;
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  ctx->pinned = nullptr;
  ctx->pinned_bytes = 0;
  size_t want = bytes < 4096 ? 4096 : bytes;
  MVS_CUDA(cudaMallocHost(&ctx->pinned, want));
  ctx->pinned_bytes = want;
  return MVS_OK;
}

int h2d(mvs_ctx* ctx, void* dst, const void* src, size_t bytes) {
  if (bytes == 0) return MVS_OK;
  MVS_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return MVS_OK;
}

// Large copies into pageable (unpinned) caller memory go through two pinned 1 MiB chunks: the DMA of
// chunk k+1 runs while the CPU copies chunk k out.  cudaMemcpyAsync straight into pageable memory
// that was never touched (a fresh result vector, the normal case behind this ABI) pins and faults
// the destination page by page inside the driver: 2.6-3.2 ms for the 8.6 MB of one 640x480
// unprojection against 0.6 ms into warm memory.
constexpr size_t kStageChunk = (size_t)1 << 20;
constexpr size_t kStageMin = (size_t)1 << 19;

bool host_is_pageable(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return true;
  }
  return a.type == cudaMemoryTypeUnregistered;
}

int d2h_async(mvs_ctx* ctx, void* dst, const void* src, size_t bytes) {
  if (bytes >= kStageMin && host_is_pageable(dst)) {
    if (!ctx->stage[0]) {
      for (int k = 0; k < 2; ++k) {
        MVS_CUDA(cudaMallocHost(&ctx->stage[k], kStageChunk));
        MVS_CUDA(cudaEventCreateWithFlags(&ctx->stage_ev[k], cudaEventDisableTiming));
      }
    }
    const size_t nchunks = (bytes + kStageChunk - 1) / kStageChunk;
    auto issue = [&](size_t k) -> cudaError_t {
      size_t off = k * kStageChunk, len = std::min(kStageChunk, bytes - off);
      cudaError_t e = cudaMemcpyAsync(ctx->stage[k & 1], (const char*)src + off, len,
                                      cudaMemcpyDeviceToHost, ctx->stream);
      if (e != cudaSuccess) return e;
      return cudaEventRecord(ctx->stage_ev[k & 1], ctx->stream);
    };
    MVS_CUDA(issue(0));
    for (size_t k = 0; k < nchunks; ++k) {
      if (k + 1 < nchunks) MVS_CUDA(issue(k + 1));  // its buffer was copied out at step k - 1
      MVS_CUDA(cudaEventSynchronize(ctx->stage_ev[k & 1]));
      size_t off = k * kStageChunk, len = std::min(kStageChunk, bytes - off);
      memcpy((char*)dst + off, ctx->stage[k & 1], len);
    }
    return MVS_OK;  // complete on return
  }
  if (bytes) MVS_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return MVS_OK;  // complete after the caller's stream synchronisation
}

int d2h_sync(mvs_ctx* ctx, void* dst, const void* src, size_t bytes) {
  MVS_TRY(d2h_async(ctx, dst, src, bytes));
  MVS_CUDA(cudaStreamSynchronize(ctx->stream));
  return MVS_OK;
}

}  // namespace mvs

namespace mvs {
static std::mutex g_ref_mu;
void ctx_retain(mvs_ctx* ctx) {
  std::lock_guard<std::mutex> lk(g_ref_mu);
  ++ctx->refs;
}
void stitch_pool_release(mvs_ctx* parent);  // stitch.cu: worker contexts of mvs_stitch_pairs

static void ctx_free(mvs_ctx* ctx) {
  stitch_pool_release(ctx);
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  mvs_comm_destroy(ctx);
  for (auto& p : ctx->pending) {
    cudaEventDestroy(p.start);
    cudaEventDestroy(p.stop);
  }
  for (auto e : ctx->event_pool) cudaEventDestroy(e);
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  if (ctx->solver_flag) cudaFreeHost(ctx->solver_flag);
  for (int k = 0; k < 2; ++k) {
    if (ctx->stage[k]) cudaFreeHost(ctx->stage[k]);
    if (ctx->stage_ev[k]) cudaEventDestroy(ctx->stage_ev[k]);
  }
  if (ctx->side) {
    cudaStreamSynchronize(ctx->side);
    cudaStreamDestroy(ctx->side);
    cudaEventDestroy(ctx->ev_fork);
    cudaEventDestroy(ctx->ev_join);
  }
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

void ctx_release(mvs_ctx* ctx) {
  {
    std::lock_guard<std::mutex> lk(g_ref_mu);
    if (--ctx->refs > 0) return;
  }
  ctx_free(ctx);
}
}  // namespace mvs

extern "C" {

const char* mvs_last_error(void) { return mvs::g_err; }

int mvs_version(void) { return 100; }

int mvs_device_count(int* count) {
  if (!count) return MVS_ERR_INVALID_ARG;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  *count = n;
  return MVS_OK;
}

int mvs_ctx_create(mvs_ctx** out, int device, void* stream) {
  if (!out) return MVS_ERR_INVALID_ARG;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    mvs::set_error(
        "no CUDA device visible: this library has no CPU fallback (cudaGetDeviceCount: %s)",
        e == cudaSuccess ? "0 devices" : cudaGetErrorString(e));
    return MVS_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= n) {
    mvs::set_error("device %d out of range (0..%d)", device, n - 1);
    return MVS_ERR_INVALID_ARG;
  }
  MVS_CUDA(cudaSetDevice(device));
  mvs_ctx* c = new mvs_ctx();
  c->device = device;
  if (stream) {
    c->stream = (cudaStream_t)stream;
    c->own_stream = false;
  } else {
    cudaError_t es = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (es != cudaSuccess) {
      mvs::set_error("cudaStreamCreate: %s", cudaGetErrorString(es));
      delete c;
      return MVS_ERR_CUDA;
    }
    c->own_stream = true;
  }
  int sms = 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0)
    c->num_sms = sms;
  // Staging buffers of the host-buffer entry points come from the stream-ordered pool
  // (cudaMallocAsync); keep freed blocks cached instead of returning them to the OS at every
  // synchronisation (a 60 MB upload otherwise pays milliseconds of map/unmap per call).
  {
    cudaMemPool_t pool = nullptr;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && pool) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
      if (getenv("MVS_POOL_NO_INTERNAL_DEPS")) {
        int off = 0;
        cudaMemPoolSetAttribute(pool, cudaMemPoolReuseAllowInternalDependencies, &off);
      }
    }
    cudaGetLastError();
  }
  if (cudaHostAlloc((void**)&c->solver_flag, 64, cudaHostAllocMapped) == cudaSuccess && c->solver_flag)
    memset(c->solver_flag, 0, 64);
  else
    c->solver_flag = nullptr;
  cudaGetLastError();
  *out = c;
  return MVS_OK;
}

int mvs_ctx_destroy(mvs_ctx* ctx) {
  if (!ctx) return MVS_OK;
  mvs::ctx_release(ctx);
  return MVS_OK;
}

int mvs_ctx_sync(mvs_ctx* ctx) {
  MVS_TRY(mvs::ctx_ensure_device(ctx));
  MVS_CUDA(cudaStreamSynchronize(ctx->stream));
  return MVS_OK;
}

int mvs_ctx_stream(mvs_ctx* ctx, void** stream) {
  if (!ctx || !stream) return MVS_ERR_INVALID_ARG;
  *stream = (void*)ctx->stream;
  return MVS_OK;
}

int mvs_prof_enable(mvs_ctx* ctx, int on) {
  if (!ctx) return MVS_ERR_INVALID_ARG;
  if (!on) MVS_TRY(mvs::ctx_resolve_prof(ctx));
  ctx->prof_on = on != 0;
  return MVS_OK;
}

int mvs_prof_reset(mvs_ctx* ctx) {
  if (!ctx) return MVS_ERR_INVALID_ARG;
  MVS_TRY(mvs::ctx_resolve_prof(ctx));
  ctx->prof.clear();
  return MVS_OK;
}

int mvs_prof_get(mvs_ctx* ctx, const char* name, double* total_ms, int64_t* launches) {
  if (!ctx || !name) return MVS_ERR_INVALID_ARG;
  MVS_TRY(mvs::ctx_resolve_prof(ctx));
  auto it = ctx->prof.find(name);
  double ms = 0.0;
  int64_t n = 0;
  if (it != ctx->prof.end()) {
    ms = it->second.total_ms;
    n = it->second.launches;
  }
  if (total_ms) *total_ms = ms;
  if (launches) *launches = n;
  return MVS_OK;
}

int mvs_prof_names(mvs_ctx* ctx, char* buf, int64_t cap, int64_t* needed) {
  if (!ctx) return MVS_ERR_INVALID_ARG;
  MVS_TRY(mvs::ctx_resolve_prof(ctx));
  std::string s;
  for (auto& kv : ctx->prof) {
    if (!s.empty()) s += "\n";
    s += kv.first;
  }
  if (needed) *needed = (int64_t)s.size() + 1;
  if (buf && cap > 0) {
    size_t ncopy = s.size() < (size_t)(cap - 1) ? s.size() : (size_t)(cap - 1);
    memcpy(buf, s.data(), ncopy);
    buf[ncopy] = 0;
  }
  return MVS_OK;
}

int mvs_launch_count(mvs_ctx* ctx, int64_t* launches) {
  if (!ctx || !launches) return MVS_ERR_INVALID_ARG;
  *launches = ctx->launch_count;
  return MVS_OK;
}

int mvs_launch_count_reset(mvs_ctx* ctx) {
  if (!ctx) return MVS_ERR_INVALID_ARG;
  ctx->launch_count = 0;
  return MVS_OK;
}

}  // extern "C"
