// common.cuh — context, error handling, device buffers, launch accounting shared by all kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/mvs_b200.h"

namespace mvs {

void set_error(const char* fmt, ...);

#define MVS_CUDA(expr)                                                                     \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      ::mvs::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return MVS_ERR_CUDA;                                                                 \
    }                                                                                      \
  } while (0)

#define MVS_TRY(expr)          \
  do {                         \
    int _s = (expr);           \
    if (_s != MVS_OK) return _s; \
  } while (0)

#define MVS_REQUIRE(cond, msg)                         \
  do {                                                 \
    if (!(cond)) {                                     \
      ::mvs::set_error("%s: %s", __func__, msg);       \
      return MVS_ERR_INVALID_ARG;                      \
    }                                                  \
  } while (0)

constexpr int kNumSMsB200 = 148;

struct ProfEntry {
  double total_ms = 0.0;
  int64_t launches = 0;
};

struct PendingProf {
  std::string name;
  cudaEvent_t start, stop;
  int64_t launches;
};

}  // namespace mvs

struct mvs_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  // forked side stream for work that is independent of the main stream's current kernel
  // (ctx_side_fork / ctx_side_join); created on first use
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int num_sms = mvs::kNumSMsB200;
  int64_t launch_count = 0;
  bool prof_on = false;
  std::map<std::string, mvs::ProfEntry> prof;
  std::vector<mvs::PendingProf> pending;
  std::vector<cudaEvent_t> event_pool;
  // pinned staging for small readbacks (reductions, scalars)
  void* pinned = nullptr;
  size_t pinned_bytes = 0;
  // two pinned chunks + events for large device -> pageable-host copies (d2h_sync)
  void* stage[2] = {nullptr, nullptr};
  cudaEvent_t stage_ev[2] = {nullptr, nullptr};
  // one pinned, device-visible word the solver kernels raise when a PCG partial sum is not finite or
  // leaves the fixed-point range (a diverging solve); read by the host after its stream synchronise
  int* solver_flag = nullptr;
  // NCCL communicator (opaque here; see comm.cu)
  void* comm = nullptr;
  int nranks = 1;
  int rank = 0;
  // lifetime: handles created on a context (NN index, SRT solver, deformation) keep it alive, so
  // destroying the context first (e.g. garbage collection order in a binding) is safe
  int refs = 1;
};

namespace mvs {

int ctx_resolve_prof(mvs_ctx* ctx);
int ctx_ensure_device(mvs_ctx* ctx);
mvs_ctx* default_ctx();
void ctx_retain(mvs_ctx* ctx);
void ctx_release(mvs_ctx* ctx);

// RAII scope that (optionally) records CUDA events around a kernel family and counts launches.
struct ProfScope {
  mvs_ctx* ctx;
  const char* name;
  int64_t launches;
  cudaEvent_t start = nullptr, stop = nullptr;
  cudaStream_t st = nullptr;  // the stream the family is launched on
  ProfScope(mvs_ctx* c, const char* n, int64_t nlaunch = 1, cudaStream_t on = nullptr);
  ~ProfScope();
};

// Fork: the side stream starts after everything queued on the main stream so far.  Join: the main
// stream waits for everything queued on the side stream so far.
int ctx_side_fork(mvs_ctx* ctx);
int ctx_side_join(mvs_ctx* ctx);

// The stream of the context the calling thread last entered (ctx_ensure_device): device buffers
// are allocated and freed in that stream's order from the cached stream-ordered pool, so the many
// temporaries of the host-buffer entry points cost microseconds instead of a cudaMalloc/cudaFree
// pair (each of which synchronises the device).
struct AllocStream {
  cudaStream_t stream = nullptr;
  bool valid = false;
  // bump arena of the calling thread (mvs_stitch_pairs: one per worker, reset per pair): while it is
  // set, device buffers are carved from it without any driver call — a pair allocates and frees
  // ~150 temporaries, and every cudaMallocAsync / cudaFreeAsync takes the process-wide driver lock
  // the workers' launches contend for
  char* arena = nullptr;
  size_t arena_cap = 0, arena_off = 0, arena_missed = 0;
};
AllocStream& alloc_stream();

// Simple owning device buffer (typed), grows but never shrinks.
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;  // elements
  cudaStream_t st = nullptr;  // stream the block was allocated on (async pool) ...
  bool async = false;         // ... or plain cudaMalloc
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  bool external = false;      // points into memory owned elsewhere (adopt): never freed here
  bool from_arena = false;    // carved from the thread's bump arena: never freed, may be re-carved
  void release() {
    if (p && !external) {
      if (async)
        cudaFreeAsync(p, st);
      else
        cudaFree(p);
    }
    p = nullptr;
    cap = 0;
    external = false;
    from_arena = false;
  }
  // view of `n` elements inside a block somebody else owns (the IPC exchange block of a sharded solve)
  void adopt(T* ptr, size_t n) {
    release();
    p = ptr;
    cap = n;
    external = true;
  }
  int reserve(size_t n) {
    if (n <= cap) return MVS_OK;
    if (external && !from_arena) {
      set_error("internal: an adopted buffer of %zu elements cannot grow to %zu", cap, n);
      return MVS_ERR_STATE;
    }
    release();
    size_t bytes = (n ? n : 1) * sizeof(T);
    AllocStream& as = alloc_stream();
    if (as.arena) {
      size_t off = (as.arena_off + 255) & ~(size_t)255;
      if (off + bytes <= as.arena_cap) {
        p = reinterpret_cast<T*>(as.arena + off);
        as.arena_off = off + bytes;
        cap = n ? n : 1;
        external = true;
        from_arena = true;
        return MVS_OK;
      }
      as.arena_missed += bytes;  // does not fit: fall through to the pool (the arena grows next time)
    }
    cudaError_t e;
    if (as.valid) {
      e = cudaMallocAsync((void**)&p, bytes, as.stream);
      st = as.stream;
      async = true;
    } else {
      e = cudaMalloc((void**)&p, bytes);
      async = false;
    }
    if (e != cudaSuccess) {
      set_error("device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
      p = nullptr;
      return MVS_ERR_CUDA;
    }
    cap = n ? n : 1;
    return MVS_OK;
  }
};

inline int div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Grid size for a grid-stride kernel: a multiple of the SM count, capped by the work available.
inline int grid_for(const mvs_ctx* ctx, int64_t work_items, int block, int ctas_per_sm) {
  int64_t need = (work_items + block - 1) / block;
  int64_t full = (int64_t)ctx->num_sms * ctas_per_sm;
  if (need >= full) return (int)full;
  return (int)(need < 1 ? 1 : need);
}

// Raises a kernel's dynamic shared-memory limit to at least `bytes` on `device` — ONCE: the limit
// only ever grows and is remembered per (kernel, device).  cudaFuncSetAttribute waits for the device
// (measured on B200, scratch/cluster_overlap2.cu: eight host threads that each launch a 1 ms kernel on
// their own stream run in 1.0 ms per round, and in 7.3 ms when each sets the attribute before its
// launch), so calling it per launch serialises every stream of the process.
int ensure_dyn_smem(const void* kernel, int device, int bytes);

int pinned_reserve(mvs_ctx* ctx, size_t bytes);
// true for ordinary (unregistered) host memory: cudaMemcpyAsync from it has staged the data when it
// returns, so the caller's buffer may be reused at once; pinned memory is read by the DMA engine later
bool host_is_pageable(const void* p);

// Host <-> device helpers on the context stream.
int h2d(mvs_ctx* ctx, void* dst, const void* src, size_t bytes);
int d2h_sync(mvs_ctx* ctx, void* dst, const void* src, size_t bytes);
// device -> host on the context stream; large copies into pageable memory are staged through pinned
// chunks (and are then complete on return), everything else completes at the next stream sync
int d2h_async(mvs_ctx* ctx, void* dst, const void* src, size_t bytes);

}  // namespace mvs
