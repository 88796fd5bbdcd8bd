// comm.cu — one-process-per-GPU NCCL plumbing (NVLink 5 / NVSwitch).  Only two collectives exist on
// this path: the all-reduce of the node-block J^T J / J^T r partial sums of a
// vertex-sharded Gauss-Newton step and of the PCG scalars.  View-pair sharding needs no data-path
// collective at all.  NCCL is resolved lazily with dlopen so the library loads (and every
// single-GPU entry point works) on hosts without libnccl.
#include <dlfcn.h>
#include <nccl.h>

#include <vector>

#include "comm.cuh"

namespace mvs {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) =
      nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;

static int load_nccl() {
  if (g_nccl.lib) return MVS_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* lib = nullptr;
  for (const char* n : names) {
    lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (lib) break;
  }
  if (!lib) {
    set_error("NCCL not found: dlopen(libnccl.so.2) failed: %s", dlerror());
    return MVS_ERR_COMM;
  }
#define MVS_SYM(field, name)                                   \
  *(void**)(&g_nccl.field) = dlsym(lib, name);                 \
  if (!g_nccl.field) {                                         \
    set_error("NCCL symbol %s missing", name);                 \
    dlclose(lib);                                              \
    return MVS_ERR_COMM;                                       \
  }
  MVS_SYM(GetUniqueId, "ncclGetUniqueId");
  MVS_SYM(CommInitRank, "ncclCommInitRank");
  MVS_SYM(CommDestroy, "ncclCommDestroy");
  MVS_SYM(AllReduce, "ncclAllReduce");
  MVS_SYM(AllGather, "ncclAllGather");
  MVS_SYM(GroupStart, "ncclGroupStart");
  MVS_SYM(GroupEnd, "ncclGroupEnd");
  MVS_SYM(GetErrorString, "ncclGetErrorString");
#undef MVS_SYM
  g_nccl.lib = lib;
  return MVS_OK;
}

#define MVS_NCCL(expr)                                                                  \
  do {                                                                                  \
    ncclResult_t _r = (expr);                                                           \
    if (_r != ncclSuccess) {                                                            \
      set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, g_nccl.GetErrorString(_r)); \
      return MVS_ERR_COMM;                                                              \
    }                                                                                   \
  } while (0)

int comm_allreduce_sum_f32(mvs_ctx* ctx, float* buf_dev, size_t count) {
  if (ctx->nranks <= 1 || count == 0) return MVS_OK;
  if (!ctx->comm) {
    set_error("all-reduce on %d ranks without a communicator (mvs_comm_init failed?)", ctx->nranks);
    return MVS_ERR_COMM;
  }
  ctx->launch_count += 1;
  MVS_NCCL(g_nccl.AllReduce(buf_dev, buf_dev, count, ncclFloat32, ncclSum, (ncclComm_t)ctx->comm,
                            ctx->stream));
  return MVS_OK;
}

int comm_allreduce_sum_f64(mvs_ctx* ctx, double* buf_dev, size_t count) {
  if (ctx->nranks <= 1 || count == 0) return MVS_OK;
  if (!ctx->comm) {
    set_error("all-reduce on %d ranks without a communicator (mvs_comm_init failed?)", ctx->nranks);
    return MVS_ERR_COMM;
  }
  ctx->launch_count += 1;
  MVS_NCCL(g_nccl.AllReduce(buf_dev, buf_dev, count, ncclFloat64, ncclSum, (ncclComm_t)ctx->comm,
                            ctx->stream));
  return MVS_OK;
}

int comm_allreduce_sum_i64(mvs_ctx* ctx, long long* buf_dev, size_t count) {
  if (ctx->nranks <= 1 || count == 0) return MVS_OK;
  if (!ctx->comm) {
    set_error("all-reduce on %d ranks without a communicator (mvs_comm_init failed?)", ctx->nranks);
    return MVS_ERR_COMM;
  }
  ctx->launch_count += 1;
  MVS_NCCL(g_nccl.AllReduce(buf_dev, buf_dev, count, ncclInt64, ncclSum, (ncclComm_t)ctx->comm,
                            ctx->stream));
  return MVS_OK;
}

// ------------------------------------------------------------------ peer-memory link
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int peer_block_alloc(mvs_ctx* ctx, size_t n64, size_t nf, PeerLink** out) {
  *out = nullptr;
  PeerLink* L = new PeerLink();
  L->off_sysf = align_up(n64 * sizeof(long long), 256);
  L->off_redf = align_up(L->off_sysf + nf * sizeof(float), 256);
  L->off_flags = align_up(L->off_redf + n64 * sizeof(float), 256);
  L->block_bytes = align_up(L->off_flags + kPeerMaxRanks * sizeof(unsigned), 1 << 21);
  cudaError_t e = cudaMalloc(&L->block, L->block_bytes);  // plain allocation: CUDA IPC needs one
  if (e != cudaSuccess) {
    set_error("peer exchange block of %zu bytes: %s", L->block_bytes, cudaGetErrorString(e));
    delete L;
    return MVS_ERR_CUDA;
  }
  MVS_CUDA(cudaMemsetAsync((char*)L->block + L->off_flags, 0, kPeerMaxRanks * sizeof(unsigned),
                           ctx->stream));
  L->dev.nranks = ctx->nranks;
  L->dev.rank = ctx->rank;
  for (int r = 0; r < kPeerMaxRanks; ++r) {
    L->dev.sys64[r] = nullptr;
    L->dev.sysf[r] = nullptr;
    L->dev.redf[r] = nullptr;
    L->dev.flags[r] = nullptr;
  }
  *out = L;
  return MVS_OK;
}

struct PeerHello {  // what the ranks swap
  cudaIpcMemHandle_t handle;
  unsigned long long block_bytes;
  unsigned long long offset;  // of the block inside the allocation the handle names
};

int peer_link_connect(mvs_ctx* ctx, PeerLink* L) {
  if (L->ready) return MVS_OK;
  if (!ctx->comm || ctx->nranks <= 1 || ctx->nranks > kPeerMaxRanks) {
    set_error("peer link: no communicator, or more than %d ranks", kPeerMaxRanks);
    return MVS_ERR_COMM;
  }
  const int n = ctx->nranks;
  PeerHello mine;
  memset(&mine, 0, sizeof(mine));
  // a rank that cannot export its block still takes part in the exchanges below (an early return
  // would leave the other ranks waiting in the all-gather): it announces 0 bytes, which every peer
  // reads as a mismatch, and the ranks then agree to stay on ncclAllReduce
  int bad = 0;
  if (cudaIpcGetMemHandle(&mine.handle, L->block) != cudaSuccess) {
    cudaGetLastError();
    memset(&mine.handle, 0, sizeof(mine.handle));
    bad = 1;
  }
  mine.block_bytes = bad ? 0 : L->block_bytes;
  // the handle names the whole underlying allocation: find the block's offset inside it
  mine.offset = 0;
  {
    typedef int (*GetRangeFn)(unsigned long long*, size_t*, unsigned long long);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qr) == cudaSuccess && fn) {
      unsigned long long base = 0;
      size_t size = 0;
      if (((GetRangeFn)fn)(&base, &size, (unsigned long long)(uintptr_t)L->block) == 0)
        mine.offset = (unsigned long long)(uintptr_t)L->block - base;
    }
    cudaGetLastError();
  }
  PeerHello* d_all = nullptr;
  MVS_CUDA(cudaMalloc((void**)&d_all, sizeof(PeerHello) * (n + 1)));
  MVS_CUDA(cudaMemcpyAsync(d_all + n, &mine, sizeof(mine), cudaMemcpyHostToDevice, ctx->stream));
  ncclResult_t r = g_nccl.AllGather(d_all + n, d_all, sizeof(PeerHello), ncclChar, (ncclComm_t)ctx->comm,
                                    ctx->stream);
  if (r != ncclSuccess) {
    cudaFree(d_all);
    set_error("peer link: ncclAllGather -> %s", g_nccl.GetErrorString(r));
    return MVS_ERR_COMM;
  }
  std::vector<PeerHello> all((size_t)n);
  MVS_CUDA(cudaMemcpyAsync(all.data(), d_all, sizeof(PeerHello) * n, cudaMemcpyDeviceToHost, ctx->stream));
  MVS_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(d_all);
  for (int k = 0; k < n && !bad; ++k) {
    char* base = nullptr;
    if (k == ctx->rank) {
      base = (char*)L->block;
    } else {
      if (all[k].block_bytes != L->block_bytes) {
        bad = 1;
        break;
      }
      void* p = nullptr;
      cudaError_t e = cudaIpcOpenMemHandle(&p, all[k].handle, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) {
        cudaGetLastError();
        bad = 1;
        break;
      }
      L->opened[k] = p;
      base = (char*)p + all[k].offset;
    }
    L->dev.sys64[k] = reinterpret_cast<long long*>(base);
    L->dev.sysf[k] = reinterpret_cast<float*>(base + L->off_sysf);
    L->dev.redf[k] = reinterpret_cast<float*>(base + L->off_redf);
    L->dev.flags[k] = reinterpret_cast<unsigned*>(base + L->off_flags);
  }
  // every rank must take the same path: agree on the outcome (sum of failures over the ranks)
  double* d_bad = nullptr;
  MVS_CUDA(cudaMalloc((void**)&d_bad, sizeof(double)));
  double hb = (double)bad;
  MVS_CUDA(cudaMemcpyAsync(d_bad, &hb, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  int rc = comm_allreduce_sum_f64(ctx, d_bad, 1);
  if (rc == MVS_OK) {
    MVS_CUDA(cudaMemcpyAsync(&hb, d_bad, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    MVS_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  cudaFree(d_bad);
  MVS_TRY(rc);
  if (hb > 0.0) {
    for (int k = 0; k < kPeerMaxRanks; ++k)
      if (L->opened[k]) {
        cudaIpcCloseMemHandle(L->opened[k]);
        L->opened[k] = nullptr;
      }
    L->failed = true;  // stay on ncclAllReduce (the systems simply live in the exchange block)
    return MVS_OK;
  }
  L->ready = true;
  return MVS_OK;
}

void peer_link_free(PeerLink* L) {
  if (!L) return;
  for (int k = 0; k < kPeerMaxRanks; ++k)
    if (L->opened[k]) cudaIpcCloseMemHandle(L->opened[k]);
  if (L->block) cudaFree(L->block);
  delete L;
}

// One thread per rank: publish "this rank reached `epoch`" in every rank's flag array (release, system
// scope), then wait until every rank has published it here.  Epochs only grow, so a fast rank that is
// already one barrier ahead cannot be mistaken for a late one.
__global__ void k_peer_barrier(PeerLinkDev L, unsigned epoch) {
  int r = threadIdx.x;
  if (r >= L.nranks) return;
  __threadfence_system();
  unsigned* remote = L.flags[r] + L.rank;
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(epoch) : "memory");
  const unsigned* local = L.flags[L.rank] + r;
  unsigned v = 0, spins = 0;
  do {
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(local) : "memory");
    if (++spins > (1u << 28)) __trap();  // a lost rank becomes an error, not a hang
  } while ((int)(v - epoch) < 0);
}

int peer_barrier(mvs_ctx* ctx, PeerLink* L) {
  ++L->epoch;
  ctx->launch_count += 1;
  k_peer_barrier<<<1, 32, 0, ctx->stream>>>(L->dev, L->epoch);
  MVS_CUDA(cudaGetLastError());
  return MVS_OK;
}

}  // namespace mvs

using namespace mvs;

extern "C" {

int mvs_comm_unique_id(char id[128]) {
  if (!id) return MVS_ERR_INVALID_ARG;
  MVS_TRY(load_nccl());
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  ncclUniqueId u;
  MVS_NCCL(g_nccl.GetUniqueId(&u));
  memcpy(id, &u, 128);
  return MVS_OK;
}

int mvs_comm_init(mvs_ctx* ctx, int nranks, int rank, const char id[128]) {
  MVS_REQUIRE(ctx && id && nranks >= 1 && rank >= 0 && rank < nranks, "bad argument");
  MVS_TRY(ctx_ensure_device(ctx));
  MVS_TRY(mvs_comm_destroy(ctx));
  if (nranks == 1) return MVS_OK;  // mvs_comm_destroy left nranks = 1, rank = 0
  MVS_TRY(load_nccl());
  ncclUniqueId u;
  memcpy(&u, id, 128);
  ncclComm_t comm = nullptr;
  MVS_NCCL(g_nccl.CommInitRank(&comm, nranks, u, rank));
  // the context becomes a member of the job only once the communicator exists
  ctx->comm = (void*)comm;
  ctx->nranks = nranks;
  ctx->rank = rank;
  return MVS_OK;
}

int mvs_comm_destroy(mvs_ctx* ctx) {
  if (!ctx) return MVS_OK;
  if (ctx->comm && g_nccl.lib) {
    cudaStreamSynchronize(ctx->stream);
    g_nccl.CommDestroy((ncclComm_t)ctx->comm);
  }
  ctx->comm = nullptr;
  ctx->nranks = 1;
  ctx->rank = 0;
  return MVS_OK;
}

int mvs_partition_units(int64_t n_units, int nranks, int rank, int64_t* bsc) {
  if (!bsc || nranks < 1 || rank < 0 || rank >= nranks || n_units < 0) {
    set_error("mvs_partition_units: bad argument");
    return MVS_ERR_INVALID_ARG;
  }
  // round-robin: unit u belongs to rank u % nranks
  bsc[0] = rank;
  bsc[1] = nranks;
  bsc[2] = n_units > rank ? (n_units - rank + nranks - 1) / nranks : 0;
  return MVS_OK;
}

}  // extern "C"
