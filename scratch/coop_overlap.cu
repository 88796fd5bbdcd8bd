// Do cooperative launches from different host threads / streams overlap?
#include <cooperative_groups.h>
#include <chrono>
#include <cstdio>
#include <thread>
#include <vector>
namespace cg = cooperative_groups;
__global__ void __launch_bounds__(512, 1) k_spin(long long cycles, int sync, int* out) {
  cg::grid_group g = cg::this_grid();
  int it = 0;
  for (int r = 0; r < 100; ++r) {  // fixed trip count: every thread reaches every grid barrier
    long long t0 = clock64();
    while (clock64() - t0 < cycles / 100) {}
    if (sync) g.sync();
    ++it;
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) *out = it;
}
int main() {
  for (int coop : {0, 1}) {
    for (int n : {1, 8}) {
      std::vector<cudaStream_t> st(n);
      std::vector<int*> out(n);
      for (int k = 0; k < n; ++k) { cudaStreamCreateWithFlags(&st[k], cudaStreamNonBlocking); cudaMalloc(&out[k], 4); }
      auto body = [&](int k, int reps) {
        cudaSetDevice(0);
        for (int r = 0; r < reps; ++r) {
          long long cyc = 1900000; int sync = coop;
          void* args[] = {&cyc, &sync, &out[k]};
          if (coop) cudaLaunchCooperativeKernel((void*)k_spin, dim3(16), dim3(512), args, 0, st[k]);
          else cudaLaunchKernel((void*)k_spin, dim3(16), dim3(512), args, 0, st[k]);
          cudaStreamSynchronize(st[k]);
        }
      };
      { std::vector<std::thread> th; for (int k = 0; k < n; ++k) th.emplace_back(body, k, 2); for (auto& t : th) t.join(); }
      auto t0 = std::chrono::steady_clock::now();
      { std::vector<std::thread> th; for (int k = 0; k < n; ++k) th.emplace_back(body, k, 10); for (auto& t : th) t.join(); }
      float ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
      printf("cooperative %d threads %d: %.2f ms per round (%s)\n", coop, n, ms / 10, cudaGetErrorString(cudaGetLastError()));
      for (int k = 0; k < n; ++k) { cudaStreamDestroy(st[k]); cudaFree(out[k]); }
    }
  }
  return 0;
}
