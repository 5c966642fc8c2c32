// Microbenchmark: cycles per tcgen05.mma instruction issued back to back by one thread (kind::f16, M=128, K=16),
// for N = 64 / 128 / 256, A operand from shared memory or from tensor memory.  One CTA per SM (all SMs busy, so the
// clocks are the loaded ones).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../av-separation-transformer_b200/csrc
#include "common.cuh"
#include <cudaTypedefs.h>
#include <cstdio>
using namespace avsep;

__global__ void __launch_bounds__(64, 1) mma_rate_kernel(const __grid_constant__ CUtensorMap tmap, int N, int a_in_tmem,
                                                         int stream_tma, int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint64_t tfull[4];
  __shared__ volatile int stop;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 64) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int i = 0; i < 4; ++i) mbar_init(&tfull[i], 1); stop = 0; fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc(&tmem_slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_slot;
  if (threadIdx.x == 32 && stream_tma) {
    // concurrent TMA fill of a 4 x 16 KB ring in the same shared memory (the weight stream of the real kernels)
    for (int i = 0; !stop; ++i) {
      const int slot = i & 3;
      if (i >= 4) mbar_wait(&tfull[slot], ((i >> 2) - 1) & 1);
      mbar_arrive_expect_tx(&tfull[slot], 16384);
      tma_load_2d(smem + 65536 + slot * 16384, &tmap, &tfull[slot], 0, ((blockIdx.x * 7 + i) & 63) * 128);
    }
  }
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
    stop = 1;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

int main() {
  long long* out;
  cudaMallocManaged(&out, 16);
  void* buf;
  cudaMalloc(&buf, 1 << 20);
  cudaMemset(buf, 0, 1 << 20);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  auto enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  CUtensorMap tm;
  cuuint64_t dims[2] = {64, 8192};
  cuuint64_t strides[1] = {128};
  cuuint32_t box[2] = {64, 128}, es[2] = {1, 1};
  enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  const int reps = 2048;
  for (int stream_tma = 0; stream_tma < 2; ++stream_tma)
    for (int a_in_tmem = 0; a_in_tmem < 2; ++a_in_tmem)
      for (int N : {64, 128, 256}) {
        mma_rate_kernel<<<148, 64, 160 * 1024>>>(tm, N, a_in_tmem, stream_tma, reps, out);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        const double math = 128.0 * N * 16 * 2 / 8192.0;     // cycles at 8192 dense bf16 FLOP/clk/SM
        printf("N=%3d A=%s %s: %.1f clk/MMA (tensor-pipe time at 8192 FLOP/clk: %.0f)\n", N, a_in_tmem ? "tmem" : "smem",
               stream_tma ? "with a concurrent TMA fill" : "alone                     ", double(out[1]) / reps, math);
      }
  return 0;
}
