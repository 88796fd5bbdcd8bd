// Do cluster kernels launched on different streams overlap on B200?
#include <cooperative_groups.h>
#include <cstdio>
#include <chrono>
#include <vector>
namespace cg = cooperative_groups;
__global__ void __launch_bounds__(512, 1) k_spin(long long cycles, int use_sync, int* out) {
  extern __shared__ float sm[];
  cg::cluster_group cl = cg::this_cluster();
  long long t0 = clock64();
  int it = 0;
  while (clock64() - t0 < cycles) {
    if (use_sync) cl.sync();
    ++it;
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) *out = it;
}
static float run(int nstreams, int cluster, size_t smem, int use_sync, int maxdyn) {
  std::vector<cudaStream_t> st(nstreams);
  for (auto& s : st) cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  int* out; cudaMalloc(&out, 4);
  if (maxdyn) cudaFuncSetAttribute(k_spin, cudaFuncAttributeMaxDynamicSharedMemorySize, maxdyn);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto launch = [&](cudaStream_t s) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cluster); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute a[1];
    a[0].id = cudaLaunchAttributeClusterDimension; a[0].val.clusterDim.x = cluster; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
    cfg.attrs = a; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, k_spin, (long long)1900000, use_sync, out);  // ~1 ms
  };
  for (auto s : st) launch(s);
  cudaDeviceSynchronize();
  cudaEventRecord(e0, 0);
  cudaDeviceSynchronize();
  auto t0 = std::chrono::steady_clock::now();
  for (auto s : st) launch(s);
  cudaDeviceSynchronize();
  auto t1 = std::chrono::steady_clock::now();
  float ms = std::chrono::duration<float, std::milli>(t1 - t0).count();
  printf("streams %2d cluster %d smem %6zu sync %d maxdyn %6d: %.2f ms  (%s)\n", nstreams, cluster, smem, use_sync, maxdyn, ms, cudaGetErrorString(cudaGetLastError()));
  for (auto s : st) cudaStreamDestroy(s);
  cudaFree(out);
  return ms;
}
#include <chrono>
int main() {
  for (int cl : {1, 4, 8}) {
    run(1, cl, 0, 1, 0);
    run(8, cl, 0, 1, 0);
    run(8, cl, 52 * 1024, 1, 150 * 1024);
    run(8, cl, 52 * 1024, 0, 150 * 1024);
  }
  return 0;
}
