#include <cooperative_groups.h>
#include <cstdio>
namespace cg = cooperative_groups;
__global__ void k_sync(int n, float* out) {
  cg::grid_group g = cg::this_grid();
  float v = threadIdx.x;
  for (int i = 0; i < n; ++i) { g.sync(); v += 1.f; }
  if (out) out[blockIdx.x * blockDim.x + threadIdx.x] = v;
}
// hand-rolled sense-reversing barrier: one atomic counter, thread 0 spins
__device__ __forceinline__ void my_barrier(unsigned* ctr, unsigned ncta, unsigned& phase) {
  __syncthreads();
  if (threadIdx.x == 0) {
    phase += ncta;
    __threadfence();
    atomicAdd(ctr, 1u);
    while (*((volatile unsigned*)ctr) < phase) {}
    __threadfence();
  }
  __syncthreads();
}
__global__ void k_mysync(int n, unsigned* ctr, float* out) {
  unsigned phase = 0;
  float v = threadIdx.x;
  for (int i = 0; i < n; ++i) { my_barrier(ctr, gridDim.x, phase); v += 1.f; }
  if (out) out[blockIdx.x * blockDim.x + threadIdx.x] = v;
}
int main() {
  float* out; cudaMalloc(&out, 4 * 1024 * 1024);
  unsigned* ctr; cudaMalloc(&ctr, 4);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int grid : {148, 256, 296}) for (int threads : {256, 512}) {
    int n = 1000; void* args[] = {&n, &out};
    cudaLaunchCooperativeKernel((void*)k_sync, dim3(grid), dim3(threads), args, 0, 0);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    cudaLaunchCooperativeKernel((void*)k_sync, dim3(grid), dim3(threads), args, 0, 0);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    cudaMemset(ctr, 0, 4);
    void* args2[] = {&n, &ctr, &out};
    cudaEventRecord(a);
    cudaLaunchCooperativeKernel((void*)k_mysync, dim3(grid), dim3(threads), args2, 0, 0);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms2; cudaEventElapsedTime(&ms2, a, b);
    printf("grid %d x %d: cg grid.sync %.2f us, hand-rolled %.2f us  (%s)\n", grid, threads, ms, ms2, cudaGetErrorString(cudaGetLastError()));
  }
}
