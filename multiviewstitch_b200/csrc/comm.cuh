// comm.cuh — internal multi-GPU helpers (see comm.cu): NCCL collectives and the peer-memory link of
// a vertex-sharded solve.
#pragma once
#include "common.cuh"
namespace mvs {
int comm_allreduce_sum_f32(mvs_ctx* ctx, float* buf_dev, size_t count);
int comm_allreduce_sum_f64(mvs_ctx* ctx, double* buf_dev, size_t count);
int comm_allreduce_sum_i64(mvs_ctx* ctx, long long* buf_dev, size_t count);

// Peer-memory link of one vertex-sharded deformation handle: every rank allocates ONE exchange block
// [fixed-point system | float system | reduced float upper triangle | flags] with cudaMalloc, the ranks swap CUDA IPC handles (one
// ncclAllGather of 64 bytes per rank) and map each other's blocks, so that a kernel of rank r can load
// the partial sums of every rank and store the reduced float system into every rank over NVLink.
constexpr int kPeerMaxRanks = 16;
struct PeerLinkDev {  // passed to kernels by value
  int nranks, rank;
  long long* sys64[kPeerMaxRanks];  // [r]: rank r's fixed-point system as mapped in this process
  float* sysf[kPeerMaxRanks];       // [r]: rank r's float system
  float* redf[kPeerMaxRanks];       // [r]: rank r's reduced upper-triangle system in float (n64 words)
  unsigned* flags[kPeerMaxRanks];   // [r]: rank r's flag array (kPeerMaxRanks words)
};
struct PeerLink {
  PeerLinkDev dev{};
  void* block = nullptr;        // this rank's exchange block (cudaMalloc)
  size_t block_bytes = 0;
  size_t off_sysf = 0, off_redf = 0, off_flags = 0;
  void* opened[kPeerMaxRanks] = {nullptr};  // peers' blocks (cudaIpcOpenMemHandle), null for self
  unsigned epoch = 0;           // barriers issued so far
  bool ready = false;
  bool failed = false;          // the ranks agreed to stay on the NCCL path (IPC unavailable somewhere)
};
// allocates the block (n64 long longs, nf floats, flags); no communication
int peer_block_alloc(mvs_ctx* ctx, size_t n64, size_t nf, PeerLink** out);
// collective over the context's communicator: swap IPC handles, map the peers
int peer_link_connect(mvs_ctx* ctx, PeerLink* L);
void peer_link_free(PeerLink* L);
// grid-less barrier over the ranks on the context's stream (flags in peer memory, system scope)
int peer_barrier(mvs_ctx* ctx, PeerLink* L);
}  // namespace mvs
