// Microbenchmark: legacy mma.sync (HMMA.16816 bf16) latency / throughput on sm_100a, and the cost of
// fence.proxy.async.shared::cta after a few shared-memory stores (both sit on the conv1 path of the visual CNN).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__global__ void hmma_kernel(int reps, int mode, long long* out, float* sink) {
  __shared__ uint32_t buf[4096];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float c0[4] = {0, 0, 0, 0}, c1[4] = {0, 0, 0, 0}, c2[4] = {0, 0, 0, 0}, c3[4] = {0, 0, 0, 0};
  uint32_t a = 0x3f803f80u + lane, b = 0x3f803f80u;
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    if (mode == 0) {                 // dependent chain
      mma16816(c0, a, a, a, a, b, b);
    } else if (mode == 1) {          // 4 independent accumulators
      mma16816(c0, a, a, a, a, b, b); mma16816(c1, a, a, a, a, b, b);
      mma16816(c2, a, a, a, a, b, b); mma16816(c3, a, a, a, a, b, b);
    } else {                         // 8 stores + fence.proxy.async
      for (int i = 0; i < 8; ++i) buf[(threadIdx.x + i * 128) & 4095] = r + i;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
  }
  const long long t1 = clock64();
  if (lane == 0) out[warp] = t1 - t0;
  if (c0[0] + c1[0] + c2[0] + c3[0] == 123.f) sink[0] = buf[lane];
}

int main() {
  long long* out; float* sink;
  cudaMallocManaged(&out, 64 * 8); cudaMalloc(&sink, 4);
  const int reps = 4000;
  const char* names[3] = {"HMMA.16816 dependent chain", "HMMA.16816 x4 independent", "8 x STS + fence.proxy.async"};
  for (int mode = 0; mode < 3; ++mode)
    for (int warps : {1, 4, 8, 16}) {
      hmma_kernel<<<1, warps * 32>>>(reps, mode, out, sink);
      cudaDeviceSynchronize();
      long long w = 0; for (int i = 0; i < warps; ++i) w = out[i] > w ? out[i] : w;
      printf("%-30s %2d warps: %.1f clk per iteration\n", names[mode], warps, double(w) / reps);
    }
  return 0;
}
