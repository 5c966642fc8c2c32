// Microbenchmark: cycles per tcgen05.mma instruction issued back to back by one thread (kind::f16, M=128, K=16),
// for N = 64 / 128 / 256, A operand from shared memory or from tensor memory.  One CTA per SM (all SMs busy, so the
// clocks are the loaded ones).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../av-separation-transformer_b200/csrc
#include "common.cuh"
#include <cstdio>
using namespace avsep;

__global__ void __launch_bounds__(64, 1) mma_rate_kernel(int N, int a_in_tmem, int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 64) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc(&tmem_slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc(1u, 128, N);
    const uint64_t adesc = umma_desc_kmajor_sw128(smem_u32(smem), 1024);
    const uint64_t bdesc = umma_desc_kmajor_sw128(smem_u32(smem + 16384), 1024);
    // warm-up
    for (int i = 0; i < 8; ++i) umma_f16(tm, adesc, bdesc, idesc, 1);
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t0 = clock64();
    if (a_in_tmem) {
      for (int i = 0; i < reps; ++i) umma_f16_ts(tm, tm + 256 + 8 * (i & 3), bdesc + 2 * (i & 3), idesc, 1);
    } else {
      for (int i = 0; i < reps; ++i) umma_f16(tm, adesc + 2 * (i & 3), bdesc + 2 * (i & 3), idesc, 1);
    }
    const long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 1);
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

int main() {
  long long* out;
  cudaMallocManaged(&out, 16);
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int reps = 2048;
  for (int a_in_tmem = 0; a_in_tmem < 2; ++a_in_tmem)
    for (int N : {32, 64, 128, 256}) {
      mma_rate_kernel<<<148, 64, 64 * 1024>>>(N, a_in_tmem, reps, out);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
      const double math = 128.0 * N * 16 * 2 / 8192.0;     // cycles at 8192 dense bf16 FLOP/clk/SM
      printf("N=%3d A=%s: issue %.1f clk/MMA, issue+drain %.1f clk/MMA (tensor-pipe time at 8192 FLOP/clk: %.0f)\n", N,
             a_in_tmem ? "tmem" : "smem", double(out[0]) / reps, double(out[1]) / reps, math);
    }
  return 0;
}
