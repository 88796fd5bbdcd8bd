// comm.cu — one-process-per-GPU NCCL plumbing (NVLink 5 / NVSwitch).  Only two collectives exist on
// this path: the all-reduce of the node-block J^T J / J^T r partial sums of a
// vertex-sharded Gauss-Newton step and of the PCG scalars.  View-pair sharding needs no data-path
// collective at all.  NCCL is resolved lazily with dlopen so the library loads (and every
// single-GPU entry point works) on hosts without libnccl.
#include <dlfcn.h>
#include <nccl.h>

#include "comm.cuh"

namespace mvs {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
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
