// Microbenchmark: sustained TMA load rate from an L2-resident buffer into a shared-memory ring (16 KB boxes,
// 128B swizzle), per SM, with 1 .. 148 CTAs streaming at once, every CTA reading the SAME 1 MB (the weight-stream
// pattern of the fused FFN / GEMM kernels) or its OWN 1 MB.
#include "common.cuh"
#include <cudaTypedefs.h>
#include <cstdio>
#include <cstdlib>
using namespace avsep;

constexpr int BOX = 16384;

template <int SLOTS>
__global__ void __launch_bounds__(64, 1) tma_rate_kernel(const __grid_constant__ CUtensorMap tm, int boxes_per_mb,
                                                         int own, int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[SLOTS];
  if (threadIdx.x == 0) {
    for (int i = 0; i < SLOTS; ++i) mbar_init(&full[i], 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int base = own ? blockIdx.x * boxes_per_mb : 0;
    const long long t0 = clock64();
    for (int i = 0; i < reps + SLOTS; ++i) {
      const int slot = i % SLOTS;
      if (i >= SLOTS) mbar_wait(&full[slot], ((i / SLOTS) - 1) & 1);      // previous fill of this slot has landed
      if (i < reps) {
        mbar_arrive_expect_tx(&full[slot], BOX);
        tma_load_2d(smem + slot * BOX, &tm, &full[slot], 0, (base + (i % boxes_per_mb)) * 128);
      }
    }
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
}

// Same ring, filled by 1-D bulk copies (cp.async.bulk, no tensor map): the source is a contiguous image of the
// shared-memory tile (weights can be prepacked pre-swizzled, so the copy does not need the TMA's address generation).
template <int SLOTS, int BYTES>
__global__ void __launch_bounds__(64, 1) bulk1d_rate_kernel(const uint8_t* src, int boxes_per_mb, int own, int reps,
                                                            long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[SLOTS];
  if (threadIdx.x == 0) {
    for (int i = 0; i < SLOTS; ++i) mbar_init(&full[i], 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const size_t base = own ? static_cast<size_t>(blockIdx.x) * boxes_per_mb * BOX : 0;
    const int per_mb = boxes_per_mb * BOX / BYTES;
    const long long t0 = clock64();
    for (int i = 0; i < reps + SLOTS; ++i) {
      const int slot = i % SLOTS;
      if (i >= SLOTS) mbar_wait(&full[slot], ((i / SLOTS) - 1) & 1);
      if (i < reps) {
        mbar_arrive_expect_tx(&full[slot], BYTES);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(smem + slot * BYTES)), "l"(src + base + static_cast<size_t>(i % per_mb) * BYTES),
                       "r"(BYTES), "r"(smem_u32(&full[slot]))
                     : "memory");
      }
    }
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
}

// P producer warps, each streaming its own ring of SLOTS x BYTES with 1-D bulk copies (is the per-instruction cost of
// ~370 clk a property of the issuing thread or of the SM's copy engine?)
template <int P, int SLOTS, int BYTES>
__global__ void __launch_bounds__(32 * P, 1) bulk1d_multi_kernel(const uint8_t* src, int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[P * SLOTS];
  if (threadIdx.x == 0) {
    for (int i = 0; i < P * SLOTS; ++i) mbar_init(&full[i], 1);
    fence_mbar_init();
  }
  __syncthreads();
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
    uint8_t* ring = smem + w * SLOTS * BYTES;
    uint64_t* bar = full + w * SLOTS;
    const int per_mb = (1 << 20) / BYTES;
    const long long t0 = clock64();
    for (int i = 0; i < reps + SLOTS; ++i) {
      const int slot = i % SLOTS;
      if (i >= SLOTS) mbar_wait(&bar[slot], ((i / SLOTS) - 1) & 1);
      if (i < reps) {
        mbar_arrive_expect_tx(&bar[slot], BYTES);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(ring + slot * BYTES)), "l"(src + static_cast<size_t>((i + w * 7) % per_mb) * BYTES),
                       "r"(BYTES), "r"(smem_u32(&bar[slot]))
                     : "memory");
      }
    }
    const long long t1 = clock64();
    out[blockIdx.x * P + w] = t1 - t0;
  }
}

int main() {
  const int n_sm = 148, boxes_per_mb = 64;                 // 64 boxes of [128 rows x 64 bf16] = 1 MB
  const size_t rows = static_cast<size_t>(n_sm) * boxes_per_mb * 128;
  void* buf;
  cudaMalloc(&buf, rows * 128);
  cudaMemset(buf, 0, rows * 128);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  auto enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  CUtensorMap tm;
  cuuint64_t dims[2] = {64, rows};
  cuuint64_t strides[1] = {128};
  cuuint32_t box[2] = {64, 128}, es[2] = {1, 1};
  if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
    printf("encode failed\n");
    return 1;
  }
  long long* out;
  cudaMallocManaged(&out, n_sm * sizeof(long long));
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const int reps = 4096;
  auto run = [&](auto kern, int slots, int own, int ctas) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, slots * BOX + 1024);
    for (int it = 0; it < 2; ++it) {       // first pass warms L2
      kern<<<ctas, 64, slots * BOX + 1024>>>(tm, boxes_per_mb, own, reps, out);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); exit(1); }
    }
    double worst = 0;
    for (int i = 0; i < ctas; ++i) worst = out[i] > worst ? out[i] : worst;
    const double bpc = double(reps) * BOX / worst;
    printf("%2d slots in flight, %s 1 MB, %3d CTAs: %.1f B/clk/SM  (%.0f GB/s/SM, %.2f TB/s aggregate at %.2f GHz)\n", slots,
           own ? "own " : "same", ctas, bpc, bpc * clk_khz * 1e-6, bpc * clk_khz * 1e-6 * ctas * 1e-3, clk_khz * 1e-6);
  };
  for (int ctas : {1, 148}) {
    run(tma_rate_kernel<2>, 2, 0, ctas);
    run(tma_rate_kernel<4>, 4, 0, ctas);
    run(tma_rate_kernel<8>, 8, 0, ctas);
    run(tma_rate_kernel<12>, 12, 0, ctas);
  }
  run(tma_rate_kernel<8>, 8, 1, 74);
  run(tma_rate_kernel<8>, 8, 1, 148);
  auto run1d = [&](auto kern, int slots, int bytes, int own, int ctas) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, slots * bytes + 1024);
    const int r1 = reps * BOX / bytes;
    for (int it = 0; it < 2; ++it) {
      kern<<<ctas, 64, slots * bytes + 1024>>>(static_cast<const uint8_t*>(buf), boxes_per_mb, own, r1, out);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); exit(1); }
    }
    double worst = 0;
    for (int i = 0; i < ctas; ++i) worst = out[i] > worst ? out[i] : worst;
    const double bpc = double(r1) * bytes / worst;
    printf("1-D bulk copies of %5d B, %2d in flight, %s 1 MB, %3d CTAs: %.1f B/clk/SM  (%.0f GB/s/SM, %.2f TB/s aggregate)\n",
           bytes, slots, own ? "own " : "same", ctas, bpc, bpc * clk_khz * 1e-6, bpc * clk_khz * 1e-6 * ctas * 1e-3);
  };
  for (int ctas : {1, 148}) {
    run1d(bulk1d_rate_kernel<4, 16384>, 4, 16384, 0, ctas);
    run1d(bulk1d_rate_kernel<8, 16384>, 8, 16384, 0, ctas);
    run1d(bulk1d_rate_kernel<4, 32768>, 4, 32768, 0, ctas);
    run1d(bulk1d_rate_kernel<16, 4096>, 16, 4096, 0, ctas);
  }
  run1d(bulk1d_rate_kernel<8, 16384>, 8, 16384, 1, 148);
  for (int ctas : {1, 148}) {
    run1d(bulk1d_rate_kernel<3, 65536>, 3, 65536, 0, ctas);
    run1d(bulk1d_rate_kernel<2, 65536>, 2, 65536, 0, ctas);
    run1d(bulk1d_rate_kernel<6, 32768>, 6, 32768, 0, ctas);
    run1d(bulk1d_rate_kernel<2, 32768>, 2, 32768, 0, ctas);
  }
  long long* outm;
  cudaMallocManaged(&outm, n_sm * 4 * sizeof(long long));
  auto runm = [&](auto kern, int producers, int slots, int bytes, int ctas) {
    const int smem_bytes = producers * slots * bytes + 1024;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    const int r1 = reps * BOX / bytes / producers;
    for (int it = 0; it < 2; ++it) {
      kern<<<ctas, 32 * producers, smem_bytes>>>(static_cast<const uint8_t*>(buf), r1, outm);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); exit(1); }
    }
    double worst = 0;
    for (int i = 0; i < ctas * producers; ++i) worst = outm[i] > worst ? outm[i] : worst;
    const double bpc = double(r1) * bytes * producers / worst;
    printf("%d producer warps x %d copies of %5d B in flight, %3d CTAs: %.1f B/clk/SM  (%.2f TB/s aggregate)\n", producers,
           slots, bytes, ctas, bpc, bpc * clk_khz * 1e-6 * ctas * 1e-3);
  };
  for (int ctas : {1, 148}) {
    runm(bulk1d_multi_kernel<2, 4, 16384>, 2, 4, 16384, ctas);
    runm(bulk1d_multi_kernel<4, 2, 16384>, 4, 2, 16384, ctas);
    runm(bulk1d_multi_kernel<2, 2, 32768>, 2, 2, 32768, ctas);
    runm(bulk1d_multi_kernel<4, 3, 8192>, 4, 3, 8192, ctas);
  }
  return 0;
}
