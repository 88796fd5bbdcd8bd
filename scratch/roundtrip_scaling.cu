// Host round trips (tiny kernel -> 4-byte D2H -> stream sync) from N host threads, one stream each:
// microseconds per round trip.  Is the driver's per-process serialisation what stretches the
// synchronisation-heavy stages of mvs_stitch_pairs?
#include <chrono>
#include <cstdio>
#include <thread>
#include <vector>
__global__ void k_tiny(int* out, int v) { if (threadIdx.x == 0) *out = v; }
int main() {
  for (int pinned : {0, 1}) {
    for (int n : {1, 2, 4, 8, 12}) {
      std::vector<cudaStream_t> st(n);
      std::vector<int*> out(n), host(n);
      for (int k = 0; k < n; ++k) {
        cudaStreamCreateWithFlags(&st[k], cudaStreamNonBlocking);
        cudaMalloc(&out[k], 4);
        if (pinned) cudaMallocHost(&host[k], 4); else host[k] = new int;
      }
      auto body = [&](int k, int reps) {
        cudaSetDevice(0);
        for (int r = 0; r < reps; ++r) {
          k_tiny<<<1, 32, 0, st[k]>>>(out[k], r);
          cudaMemcpyAsync(host[k], out[k], 4, cudaMemcpyDeviceToHost, st[k]);
          cudaStreamSynchronize(st[k]);
        }
      };
      { std::vector<std::thread> th; for (int k = 0; k < n; ++k) th.emplace_back(body, k, 200); for (auto& t : th) t.join(); }
      auto t0 = std::chrono::steady_clock::now();
      const int reps = 2000;
      { std::vector<std::thread> th; for (int k = 0; k < n; ++k) th.emplace_back(body, k, reps); for (auto& t : th) t.join(); }
      float us = std::chrono::duration<float, std::micro>(std::chrono::steady_clock::now() - t0).count();
      printf("pinned %d threads %2d: %.1f us per round trip per thread, %.0f round trips/s in total\n", pinned, n, us / reps, 1e6 * n * reps / us);
      for (int k = 0; k < n; ++k) { cudaStreamDestroy(st[k]); cudaFree(out[k]); if (pinned) cudaFreeHost(host[k]); else delete host[k]; }
    }
  }
  return 0;
}
