// TMA-fed tcgen05 GEMM for sm_100a with fused epilogues.
//
//   acc[m, n] = sum_{tap < taps} sum_{k < K} A[m + row_shift + tap, k] * W[n, tap*tap_stride + k]
//
// taps = 1 is every nn.Linear of the path (QKV / out-proj / FFN / frame_proj / decoder);
// taps = 3 is nn.Conv1d(k=3, padding=1) as an implicit GEMM: three row-shifted TMA views of the
// zero-haloed activation accumulate into the same TMEM tile (reference: model.py:37-42,56).
//
// Structure (one 128 x n_tile output tile per CTA, 192 threads):
//   warp 0 / lane 0 : TMA producer  - cp.async.bulk.tensor A/W tiles (128B swizzle) into a STAGES-deep ring
//   warp 1 / lane 0 : MMA issuer    - tcgen05.mma (M=128, N=n_tile, K=16|8) accumulating in TMEM,
//                                     tcgen05.commit releases smem stages / signals the epilogue
//   warps 2..5      : epilogue      - tcgen05.ld (thread = accumulator row), bias / ReLU / exact GELU /
//                                     positional-encoding add / Conv1d halo handling, or the decoder tail
//                                     (sigmoid, x mixed_spec, (B,S,F,T) stores coalesced along T)
// Two or more CTAs are resident per SM (<= 113 KB smem each), so one CTA's epilogue overlaps another's main loop.
#include "common.cuh"
#include "kernels.h"

#include <cudaTypedefs.h>
#include <stdio.h>

namespace avsep {

namespace {

constexpr int BM = 128;
constexpr int TILE_K_BYTES = 128;            // one 128-byte swizzle span per row per stage
constexpr int A_STAGE_BYTES = BM * TILE_K_BYTES;
constexpr int NUM_THREADS = 192;

struct GemmDev {
  int M, N, K, taps, tap_stride, row_shift, n_tile;
  GemmEpilogue e;
};

template <int BN_MAX, int STAGES>
struct SmemLayout {
  static constexpr int B_STAGE_BYTES = BN_MAX * TILE_K_BYTES;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;   // barriers + alignment slack
};

template <bool TF32>
__device__ __forceinline__ void store_op_chunk(void* out_op, size_t off, const float (&val)[32], int ncols, bool vec) {
  if constexpr (TF32) {
    float* o = reinterpret_cast<float*>(out_op) + off;
    if (vec) {
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(o + j) = make_float4(val[j], val[j + 1], val[j + 2], val[j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < ncols) o[j] = val[j];
    }
  } else {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out_op) + off;
    if (vec) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint4 u;
        u.x = pack_bf16x2(val[j], val[j + 1]);
        u.y = pack_bf16x2(val[j + 2], val[j + 3]);
        u.z = pack_bf16x2(val[j + 4], val[j + 5]);
        u.w = pack_bf16x2(val[j + 6], val[j + 7]);
        *reinterpret_cast<uint4*>(o + j) = u;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < ncols) o[j] = __float2bfloat16_rn(val[j]);
    }
  }
}

template <int BN_MAX, int STAGES, bool TF32>
__global__ void __launch_bounds__(NUM_THREADS, (BN_MAX <= 128 ? 2 : 1))
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const GemmDev p) {
  using L = SmemLayout<BN_MAX, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * p.n_tile;
  constexpr int ELEMS_PER_TILE = TF32 ? 32 : 64;                 // elements per 128-byte K span
  const int kt_per_tap = (p.K + ELEMS_PER_TILE - 1) / ELEMS_PER_TILE;
  const int num_kt = kt_per_tap * p.taps;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, BN_MAX);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer ----------------
      const uint32_t stage_bytes = A_STAGE_BYTES + static_cast<uint32_t>(p.n_tile) * TILE_K_BYTES;
      for (int kt = 0; kt < num_kt; ++kt) {
        const int s = kt % STAGES;
        const uint32_t ph = (kt / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        const int tap = kt / kt_per_tap;
        const int kk = (kt - tap * kt_per_tap) * ELEMS_PER_TILE;
        uint8_t* a_dst = smem + s * L::STAGE_BYTES;
        uint8_t* b_dst = a_dst + A_STAGE_BYTES;
        mbar_arrive_expect_tx(&full_bar[s], stage_bytes);
        tma_load_2d(a_dst, &tmA, &full_bar[s], kk, m0 + p.row_shift + tap);
        tma_load_2d(b_dst, &tmW, &full_bar[s], tap * p.tap_stride + kk, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- MMA issuer ----------------
      const uint32_t idesc = umma_idesc(TF32 ? 2u : 1u, BM, static_cast<uint32_t>(p.n_tile));
      for (int kt = 0; kt < num_kt; ++kt) {
        const int s = kt % STAGES;
        const uint32_t ph = (kt / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * L::STAGE_BYTES);
        const uint32_t b_addr = a_addr + A_STAGE_BYTES;
        const uint64_t adesc = umma_desc_kmajor_sw128(a_addr, 1024);
        const uint64_t bdesc = umma_desc_kmajor_sw128(b_addr, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k) {   // 4 x 32 bytes of K per stage: advance the start address by 32 B (>>4 = 2)
          const uint32_t acc = (kt | k) != 0 ? 1u : 0u;
          if constexpr (TF32) umma_tf32(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, acc);
          else                umma_f16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, acc);
        }
        umma_commit(&empty_bar[s]);      // smem stage reusable once these MMAs have read it
      }
      umma_commit(accum_bar);            // accumulator complete
    }
  } else {
    // ---------------- epilogue: warps 2..5, TMEM lane quarter = warp % 4 ----------------
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int m = m0 + r;
    const GemmEpilogue& e = p.e;
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const uint32_t taddr_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);

    bool valid = m < p.M;
    bool zero_row = false;
    long long orow = m;
    int pos = 0;
    if (e.kind == EPI_STD) {
      if (e.rowmap == ROW_IDENT) {
        if (e.pe_period > 0) pos = m % e.pe_period;
      } else {
        const int b = m / e.Lp;
        const int tp = m - b * e.Lp;
        const bool halo = (tp == 0) || (tp == e.Lp - 1);
        pos = tp - 1;
        if (e.rowmap == ROW_PAD2PAD) {
          zero_row = halo;
        } else {
          valid = valid && !halo;
          orow = static_cast<long long>(b) * (e.Lp - 2) + (tp - 1);
        }
      }
    }
    int tb = 0, tt = 0;
    if (e.kind == EPI_TAIL) {
      tb = m / e.T;
      tt = m - tb * e.T;
    }

    for (int c0 = 0; c0 < p.n_tile; c0 += 32) {
      const int col0 = n0 + c0;
      if (col0 >= p.N) break;                      // uniform across the CTA
      uint32_t v[32];
      tmem_ld_32x32b_x32(taddr_row + static_cast<uint32_t>(c0), v);
      tmem_ld_wait();
      const int ncols = min(32, min(p.N - col0, p.n_tile - c0));
      if (!valid) continue;
      float val[32];
      const bool full = (ncols == 32);
      if (full && e.bias != nullptr && (p.N & 3) == 0) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(e.bias + col0 + j));
          val[j] = __uint_as_float(v[j]) + b4.x;
          val[j + 1] = __uint_as_float(v[j + 1]) + b4.y;
          val[j + 2] = __uint_as_float(v[j + 2]) + b4.z;
          val[j + 3] = __uint_as_float(v[j + 3]) + b4.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float bj = (e.bias != nullptr && j < ncols) ? __ldg(e.bias + col0 + j) : 0.0f;
          val[j] = __uint_as_float(v[j]) + bj;
        }
      }

      if (e.kind == EPI_TAIL) {
        // SeparationDecoder head (model.py:204-207,220): column c = s*F + f; lanes are consecutive t, so the
        // (B,S,F,T) stores and the mixed_spec (B,F,T) loads are both contiguous across the warp.
        int s = col0 / e.F;
        int f = col0 - s * e.F;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (j < ncols) {
            const float mk = 1.0f / (1.0f + __expf(-val[j]));
            const size_t o = ((static_cast<size_t>(tb) * e.S + s) * e.F + f) * e.T + tt;
            const float mx = __ldg(e.mixed + (static_cast<size_t>(tb) * e.F + f) * e.T + tt);
            e.masks[o] = mk;
            e.separated[o] = mk * mx;
            if (++f == e.F) { f = 0; ++s; }
          }
        }
        continue;
      }

      if (e.act == ACT_RELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) val[j] = fmaxf(val[j], 0.0f);
      } else if (e.act == ACT_GELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) val[j] = gelu_erf(val[j]);
      }
      if (e.pe != nullptr) {
        const float* pr = e.pe + static_cast<size_t>(pos) * p.N + col0;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < ncols) val[j] += __ldg(pr + j);
      }
      if (zero_row) {
#pragma unroll
        for (int j = 0; j < 32; ++j) val[j] = 0.0f;
      }
      if (e.out_f32 != nullptr) {
        float* o = e.out_f32 + static_cast<size_t>(orow) * e.ld_f32 + col0;
        if (full && (e.ld_f32 & 3) == 0) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(o + j) = make_float4(val[j], val[j + 1], val[j + 2], val[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < ncols) o[j] = val[j];
        }
      }
      if (e.out_op != nullptr) {
        const bool vec = full && ((e.ld_op & 7) == 0);
        store_op_chunk<TF32>(e.out_op, static_cast<size_t>(orow) * e.ld_op + col0, val, ncols, vec);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN_MAX);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;

const char* encode_2d(CUtensorMap* map, bool tf32, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld_elems,
                      uint32_t box_inner, uint32_t box_outer) {
  const uint64_t esz = tf32 ? 4 : 2;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) return "gemm: operand pointer not 16-byte aligned";
  if (((ld_elems * esz) & 15) != 0) return "gemm: operand leading dimension not a multiple of 16 bytes";
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld_elems * esz};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(map, tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                        const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return "gemm: cuTensorMapEncodeTiled failed";
  return nullptr;
}

template <int BN_MAX, int STAGES, bool TF32>
const char* launch_cfg(cudaStream_t s, const CUtensorMap& ta, const CUtensorMap& tw, const GemmDev& d) {
  using L = SmemLayout<BN_MAX, STAGES>;
  static bool attr_done = false;
  auto kern = gemm_tcgen05_kernel<BN_MAX, STAGES, TF32>;
  if (!attr_done) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL) != cudaSuccess)
      return "gemm: cudaFuncSetAttribute(smem) failed";
    attr_done = true;
  }
  dim3 grid((d.M + BM - 1) / BM, (d.N + d.n_tile - 1) / d.n_tile, 1);
  kern<<<grid, NUM_THREADS, L::TOTAL, s>>>(ta, tw, d);
  if (cudaGetLastError() != cudaSuccess) return "gemm: kernel launch failed";
  return nullptr;
}

}  // namespace

const char* gemm_init() {
  if (g_encode != nullptr) return nullptr;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess || fn == nullptr)
    return "gemm: cuTensorMapEncodeTiled entry point not found (driver too old?)";
  g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  return nullptr;
}

const char* launch_gemm(cudaStream_t s, int prec, const GemmProblem& p, const GemmEpilogue& e, int force_bn) {
  if (const char* err = gemm_init()) return err;
  if (p.M <= 0 || p.N <= 0 || p.K <= 0) return "gemm: empty problem";
  const bool tf32 = (prec == PREC_TF32);
  // n_tile: split N evenly over ceil(N / bn_max) tiles, rounded up to the UMMA N granularity (16).
  int bn_max = force_bn > 0 ? force_bn : (p.N > 128 && (p.N % 128 != 0 || p.N >= 512) ? 256 : 128);
  if (e.kind == EPI_TAIL) bn_max = force_bn > 0 ? force_bn : 256;
  const int ntiles = (p.N + bn_max - 1) / bn_max;
  int n_tile = ((p.N + ntiles - 1) / ntiles + 15) / 16 * 16;
  if (n_tile > bn_max) n_tile = bn_max;
  GemmDev d;
  d.M = p.M; d.N = p.N; d.K = p.K; d.taps = p.taps; d.tap_stride = p.tap_stride; d.row_shift = p.row_shift;
  d.n_tile = n_tile;
  d.e = e;
  CUtensorMap ta, tw;
  const uint32_t box_k = tf32 ? 32 : 64;
  if (const char* err = encode_2d(&ta, tf32, p.A, p.K, p.rowsA, p.lda, box_k, BM)) return err;
  if (const char* err = encode_2d(&tw, tf32, p.W, static_cast<uint64_t>(p.ldw), p.N, p.ldw, box_k, n_tile)) return err;
  if (bn_max == 128) {
    return tf32 ? launch_cfg<128, 3, true>(s, ta, tw, d) : launch_cfg<128, 3, false>(s, ta, tw, d);
  } else if (bn_max == 256) {
    return tf32 ? launch_cfg<256, 4, true>(s, ta, tw, d) : launch_cfg<256, 4, false>(s, ta, tw, d);
  }
  return "gemm: unsupported tile width";
}

}  // namespace avsep
