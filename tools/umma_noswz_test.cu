// Bring-up check for the shifted-view implicit GEMM of the visual CNN: tcgen05.mma with an A operand in the
// NO-SWIZZLE K-major layout, stored "chunk-major" (for every 16-byte K chunk an array of rows, 16 B per row), so that an
// 8-row x 16-byte core matrix starting at ANY row is 128 contiguous bytes and a row-shifted view of the activation is
// just a different start address.  Prints the max error for both readings of (leading, stride) byte offsets.
#include "common.cuh"
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
using namespace avsep;

constexpr int ROWS = 160, KTOT = 32, NB = 64;

__device__ __forceinline__ uint64_t desc_noswz(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}

__global__ void __launch_bounds__(128, 1) test_kernel(const uint16_t* a_chunks, const uint16_t* b_sw, float* out, int shift,
                                                     int swap) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                       // [4 chunks][ROWS][16 B]
  uint8_t* sB = smem + 16384;               // [64 rows][128 B] SW128
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 4 * ROWS * 16 / 4; i += 128) reinterpret_cast<uint32_t*>(sA)[i] = reinterpret_cast<const uint32_t*>(a_chunks)[i];
  for (int i = threadIdx.x; i < NB * 128 / 4; i += 128) reinterpret_cast<uint32_t*>(sB)[i] = reinterpret_cast<const uint32_t*>(b_sw)[i];
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc(&slot, 64);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc(1u, 128, 64);
    for (int ks = 0; ks < 2; ++ks) {
      const uint32_t a_addr = smem_u32(sA) + ((ks * 2) * ROWS + shift) * 16;
      const uint32_t lbo = ROWS * 16, sbo = 128;
      const uint64_t ad = swap ? desc_noswz(a_addr, sbo, lbo) : desc_noswz(a_addr, lbo, sbo);
      const uint64_t bd = umma_desc_kmajor_sw128(smem_u32(sB), 1024) + 2 * ks;
      umma_f16(tm, ad, bd, idesc, ks != 0 ? 1u : 0u);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t v[32];
  for (int c = 0; c < 2; ++c) {
    tmem_ld_32x32b_x32(tm + (static_cast<uint32_t>(warp * 32) << 16) + c * 32, v);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * 64 + c * 32 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 64);
}

static uint16_t f2bf(float f) { uint32_t u; memcpy(&u, &f, 4); u += 0x7fffu + ((u >> 16) & 1u); return u >> 16; }
static float bf2f(uint16_t b) { uint32_t u = static_cast<uint32_t>(b) << 16; float f; memcpy(&f, &u, 4); return f; }

int main() {
  std::vector<float> A(ROWS * KTOT), B(NB * KTOT);
  srand(1);
  for (auto& x : A) x = bf2f(f2bf((rand() % 2001 - 1000) / 500.0f));
  for (auto& x : B) x = bf2f(f2bf((rand() % 2001 - 1000) / 500.0f));
  std::vector<uint16_t> ac(4 * ROWS * 8), bs(NB * 64, 0);
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < ROWS; ++r)
      for (int e = 0; e < 8; ++e) ac[(c * ROWS + r) * 8 + e] = f2bf(A[r * KTOT + c * 8 + e]);
  for (int n = 0; n < NB; ++n)
    for (int k = 0; k < KTOT; ++k) bs[n * 64 + (((k >> 3) ^ (n & 7)) << 3) + (k & 7)] = f2bf(B[n * KTOT + k]);
  uint16_t *da, *db;
  float* dout;
  cudaMalloc(&da, ac.size() * 2); cudaMalloc(&db, bs.size() * 2); cudaMalloc(&dout, 128 * 64 * 4);
  cudaMemcpy(da, ac.data(), ac.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, bs.data(), bs.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  std::vector<float> out(128 * 64);
  for (int shift : {0, 8, 5, 13})
    for (int swap = 0; swap < 1; ++swap) {
      cudaMemset(dout, 0, 128 * 64 * 4);
      test_kernel<<<1, 128, 32768>>>(da, db, dout, shift, swap);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("shift %d swap %d: error %s\n", shift, swap, cudaGetErrorString(cudaGetLastError())); return 1; }
      cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
      double err = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < NB; ++n) {
          double ref = 0;
          for (int k = 0; k < KTOT; ++k) ref += double(A[(m + shift) * KTOT + k]) * B[n * KTOT + k];
          err = fmax(err, fabs(ref - out[m * 64 + n]));
        }
      printf("shift %2d, %s: max err %.4g\n", shift, swap ? "desc(LBO = 8-row group stride, SBO = K-chunk stride)" : "desc(LBO = K-chunk stride, SBO = 8-row group stride)", err);
    }
  return 0;
}
