// Microbenchmark: tensor-memory load / store throughput per SM as seen by epilogue warps (tcgen05.ld / tcgen05.st,
// 32x32b.x32: every thread moves 32 fp32 columns of its own lane), with 4, 8 and 16 warps issuing at once.
#include "common.cuh"
#include <cstdio>
#include <cstdlib>
using namespace avsep;

template <int MODE>   // 0 = ld, 1 = st, 2 = ld + st back
__global__ void __launch_bounds__(640, 1) tmem_rate_kernel(int warps, int reps, long long* out, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot;
  float acc = 0.f;
  long long t0 = 0, t1 = 0;
  if (warp < warps) {
    const int q = warp & 3, part = warp >> 2;               // lane quarter, column part
    const uint32_t addr = base + (static_cast<uint32_t>(q * 32) << 16) + (part & 15) * 32;
    uint32_t v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = lane + i;
    tmem_st_32x32b_x32(addr, v);
    tmem_st_wait();
    __syncwarp();
    t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      if (MODE == 0 || MODE == 2) {
        tmem_ld_32x32b_x32(addr, v);
        tmem_ld_wait();
        acc += __uint_as_float(v[r & 31]);
      }
      if (MODE == 1 || MODE == 2) {
        v[0] = r;
        tmem_st_32x32b_x32(addr, v);
        tmem_st_wait();
      }
    }
    t1 = clock64();
    if (lane == 0) out[warp] = t1 - t0;
  }
  if (acc == 123.456f) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(base, 512);
}

int main() {
  long long* out;
  float* sink;
  cudaMallocManaged(&out, 32 * sizeof(long long));
  cudaMalloc(&sink, 4);
  const int reps = 20000;
  const char* names[3] = {"ld x32", "st x32", "ld+st x32"};
  for (int mode = 0; mode < 3; ++mode)
    for (int warps : {1, 4, 8, 16}) {
      for (int it = 0; it < 2; ++it) {
        if (mode == 0) tmem_rate_kernel<0><<<1, 640>>>(warps, reps, out, sink);
        if (mode == 1) tmem_rate_kernel<1><<<1, 640>>>(warps, reps, out, sink);
        if (mode == 2) tmem_rate_kernel<2><<<1, 640>>>(warps, reps, out, sink);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
      }
      long long worst = 0;
      for (int i = 0; i < warps; ++i) worst = out[i] > worst ? out[i] : worst;
      const double bytes = double(reps) * warps * 32 * 32 * 4 * (mode == 2 ? 2 : 1);
      printf("%-10s %2d warps: %.1f clk per op per warp, %.1f B/clk/SM\n", names[mode], warps, double(worst) / reps, bytes / worst);
    }
  return 0;
}
