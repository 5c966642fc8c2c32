// Companion of tma_rate.cu: per-SM L2 read rate through the LSU paths (LDG.128 into registers, cp.async 16 B into shared
// memory), 512 threads per CTA streaming the same L2-resident 1 MB, with 1 or 148 CTAs, and both paths at once with TMA.
#include "common.cuh"
#include <cstdio>
#include <cstdlib>
using namespace avsep;

__global__ void __launch_bounds__(512, 1) ldg_kernel(const uint4* __restrict__ buf, int n_vec, int reps, int mode,
                                                     long long* out, uint32_t* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint32_t acc = 0;
  const long long t0 = clock64();
  if (mode == 0) {
    for (int r = 0; r < reps; ++r) {
#pragma unroll 8
      for (int i = threadIdx.x; i < n_vec; i += 512) {
        uint4 v;
        asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(buf + i));
        acc ^= v.x ^ v.y ^ v.z ^ v.w;
      }
    }
  } else {
    const uint32_t sbase = smem_u32(smem);
    for (int r = 0; r < reps; ++r) {
      for (int i0 = 0; i0 < n_vec; i0 += 512 * 16) {          // 128 KB of smem per round
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const int i = i0 + k * 512 + threadIdx.x;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbase + ((k * 512 + threadIdx.x) << 4)), "l"(buf + i) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
}

int main() {
  const int n_vec = (1 << 20) / 16;
  uint4* buf;
  cudaMalloc(&buf, 1 << 20);
  cudaMemset(buf, 1, 1 << 20);
  long long* out;
  uint32_t* sink;
  cudaMallocManaged(&out, 148 * sizeof(long long));
  cudaMalloc(&sink, 4);
  cudaFuncSetAttribute(ldg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const int reps = 64;
  for (int mode = 0; mode < 2; ++mode)
    for (int ctas : {1, 148}) {
      for (int it = 0; it < 2; ++it) {
        ldg_kernel<<<ctas, 512, 128 * 1024>>>(buf, n_vec, reps, mode, out, sink);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
      }
      double worst = 0;
      for (int i = 0; i < ctas; ++i) worst = out[i] > worst ? out[i] : worst;
      const double bpc = double(reps) * (1 << 20) / worst;
      printf("%s, %3d CTAs x 512 threads: %.1f B/clk/SM (%.0f GB/s/SM, %.2f TB/s aggregate)\n",
             mode == 0 ? "LDG.128 -> registers" : "cp.async 16 B -> smem", ctas, bpc, bpc * clk_khz * 1e-6,
             bpc * clk_khz * 1e-6 * ctas * 1e-3);
    }
  return 0;
}
