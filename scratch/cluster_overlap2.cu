// Host threads, one stream each: launch a ~1 ms cluster kernel then synchronise the stream.
#include <cooperative_groups.h>
#include <chrono>
#include <cstdio>
#include <thread>
#include <vector>
namespace cg = cooperative_groups;
__global__ void __launch_bounds__(512, 1) k_spin(long long cycles, int* out) {
  extern __shared__ float sm[];
  cg::cluster_group cl = cg::this_cluster();
  long long t0 = clock64();
  int it = 0;
  while (clock64() - t0 < cycles) { cl.sync(); ++it; }
  if (threadIdx.x == 0 && blockIdx.x == 0) *out = it;
}
static void launch(cudaStream_t s, int cluster, int* out, int setattr) {
  if (setattr) cudaFuncSetAttribute(k_spin, cudaFuncAttributeMaxDynamicSharedMemorySize, 150 * 1024);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cluster); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = 52 * 1024; cfg.stream = s;
  cudaLaunchAttribute a[1];
  a[0].id = cudaLaunchAttributeClusterDimension; a[0].val.clusterDim.x = cluster; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
  cfg.attrs = a; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, k_spin, (long long)1900000, out);
}
int main() {
  cudaFuncSetAttribute(k_spin, cudaFuncAttributeMaxDynamicSharedMemorySize, 150 * 1024);
  for (int mode = 0; mode < 4; ++mode) {
    int setattr = mode & 1, use_malloc = mode >> 1;
    for (int n : {1, 8}) {
      std::vector<cudaStream_t> st(n);
      std::vector<int*> out(n);
      for (int k = 0; k < n; ++k) { cudaStreamCreateWithFlags(&st[k], cudaStreamNonBlocking); cudaMalloc(&out[k], 4); }
      auto body = [&](int k, int reps) {
        cudaSetDevice(0);
        for (int r = 0; r < reps; ++r) {
          int* tmp = nullptr;
          if (use_malloc) cudaMallocAsync((void**)&tmp, 1 << 20, st[k]);
          launch(st[k], 8, out[k], setattr);
          int h;
          cudaMemcpyAsync(&h, out[k], 4, cudaMemcpyDeviceToHost, st[k]);
          if (use_malloc) cudaFreeAsync(tmp, st[k]);
          cudaStreamSynchronize(st[k]);
        }
      };
      { std::vector<std::thread> th; for (int k = 0; k < n; ++k) th.emplace_back(body, k, 2); for (auto& t : th) t.join(); }
      auto t0 = std::chrono::steady_clock::now();
      { std::vector<std::thread> th; for (int k = 0; k < n; ++k) th.emplace_back(body, k, 10); for (auto& t : th) t.join(); }
      float ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
      printf("setattr-per-launch %d mallocasync %d threads %d: %.2f ms per round\n", setattr, use_malloc, n, ms / 10);
      for (int k = 0; k < n; ++k) { cudaStreamDestroy(st[k]); cudaFree(out[k]); }
    }
  }
  return 0;
}
