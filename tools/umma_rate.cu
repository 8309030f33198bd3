// Microbenchmark: issue rate of tcgen05.mma kind::tf32 (SS operands, M = 128) for several N, and of
// dependent vs independent accumulators.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
//   -I ignnition_b200/csrc tools/umma_rate.cu -o gpurun_out/umma_rate ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_common.cuh"
void ign_set_error(const char*, ...) {}
void ign_count_launch() {}
using namespace ign_tc;

__global__ void __launch_bounds__(128, 1) rate_kernel(int n, int iters, int n_acc, long long* cycles) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.001f * (i & 7);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) tmem_alloc(&tmem_base_s, 512);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t a = smem_u32(smem), b = a + 16384;
    const uint32_t idesc = umma_idesc(n);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t d = tmem_base_s + (it % n_acc) * 256;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) umma_tf32(d, umma_desc(a + kk * 32), umma_desc(b + kk * 32), idesc, 1u);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) *cycles = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem_base_s, 512);
}

int main() {
  long long* d_c; cudaMalloc(&d_c, 8);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 2000;
  for (int grid : {1, 148}) for (int n_acc : {1, 2}) for (int n : {32, 64, 96, 128, 256}) {
    rate_kernel<<<grid, 128, 64 * 1024>>>(n, iters, n_acc, d_c);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0; cudaMemcpy(&c, d_c, 8, cudaMemcpyDeviceToHost);
    double per = (double)c / (iters * 4);
    printf("grid %3d accumulators %d N %3d : %8.1f cycles per tcgen05.mma (M128 N%d K8 tf32) -> %.0f MAC/cycle/SM  %s\n",
           grid, n_acc, n, per, n, 128.0 * n * 8 / per, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
