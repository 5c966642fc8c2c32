// VisualEncoder.conv (model.py:81-92,106-107) for 32x32 frames as shifted-view implicit GEMMs on tcgen05.
//
// The stride-2 3x3 convolutions conv2 (32 -> 64 channels, 16x16 -> 8x8) and conv3 (64 -> 128, 8x8 -> 4x4) never
// build an im2col matrix.  Their input activation is kept in shared memory as four PARITY PLANES (pixel (Y, X) goes to
// plane (Y & 1, X & 1) at (Y >> 1, X >> 1)), each with one zero row above and one zero column to the left, flattened
// with the same pitch as the OUTPUT raster.  Output pixel (y, x), tap (ky, kx) then reads plane ((ky != 1), (kx != 1))
// at the output's own flattened index plus a constant (-pitch if ky == 0, -1 if kx == 0): every tap is the same GEMM
// on a row-shifted view of one plane.  The planes are stored "chunk-major" (for every 16-byte channel chunk an array of
// rows, 16 B per row): the no-swizzle K-major UMMA layout, in which an 8-row core matrix starting at ANY row is 128
// contiguous bytes, so a shifted view is just another start address in the shared-memory descriptor
// (tools/umma_noswz_test.cu).  The raster rows that belong to the zero border are computed and thrown away (81 rows
// per 64 pixels in conv2, 32 per 16 in conv3): the tensor pipe has the time, the CUDA cores do not -- the previous
// kernel (visual_cnn_tc.cu) spent half of its time gathering im2col slabs.
//
//   conv2, per frame : D[128 raster rows x 64 ch]  = act1 views (A, smem) x W2 (B, smem, 128B-swizzled slabs), 18 MMAs
//   conv3, per 3 frames, TRANSPOSED: D[128 ch x 96 raster rows] = W3 (A, resident in TENSOR MEMORY: 128 lanes x 288
//          columns, loaded once per CTA) x act2 views (B, smem), 36 MMAs.  The weights never pass through shared memory
//          (these MMAs are bound by shared-memory operand bandwidth), and the epilogue thread owns one channel: bias +
//          ReLU + the mean over a frame's 16 pixels is a plain sum over columns, no shuffles.
//
// Per CTA (persistent over groups of 3 frames), 17 warps in three decoupled roles that meet only through mbarriers:
// warps 0..7 stage the frame and run conv1 (mma.sync, K = 9) into the act1 planes (double buffered per frame); warps
// 8..15 run the conv2 epilogue (bias + ReLU -> act2 planes, double buffered per group) and the conv3 epilogue; warp 16
// lane 0 issues every tcgen05.mma.  BatchNorm (eval) is folded into weights /
// bias on the host.  Only the 4 KB frame and the 256 B feature row touch HBM.
#include "common.cuh"
#include "kernels.h"

#include <stdlib.h>
#include <string.h>

namespace avsep {

namespace {

constexpr int IG_BUILDERS = 512;
constexpr int IG_THREADS = IG_BUILDERS + 32;
constexpr int GROUP = 3;                          // frames per conv3 tile
constexpr int PITCH1 = 9, FR1 = 81, PAD1 = 16, ROWS1 = 104;                   // act1 plane arrays (one frame): 16 + 81 -> 104
constexpr int PITCH2 = 5, FR2 = 32, PAD2 = 8, ROWS2 = PAD2 + GROUP * FR2;     // act2 plane arrays (one group): 104
constexpr int ACT1_BYTES = 16 * ROWS1 * 16;       // [4 planes][4 chunks][ROWS1][16 B]
constexpr int ACT2_BYTES = 32 * ROWS2 * 16;       // [4 planes][8 chunks][ROWS2][16 B]
constexpr int OFF_W2 = 0;                                   // 5 x 8 KB    conv2 weight slabs (resident, 128B swizzle)
constexpr int OFF_ACT1 = OFF_W2 + 5 * 8192;                 // 2 buffers
constexpr int OFF_ACT2 = OFF_ACT1 + 2 * ACT1_BYTES;         // 2 buffers
constexpr int OFF_IN = OFF_ACT2 + 2 * ACT2_BYTES;           // 2 x 34 x 34 fp32 (zero border), double buffered
constexpr int OFF_MISC = OFF_IN + 2 * 34 * 34 * 4;
constexpr int IG_SMEM = OFF_MISC + 256;
static_assert(OFF_MISC % 8 == 0, "barrier alignment");
static_assert(IG_SMEM <= 227 * 1024, "visual_cnn_ig: shared memory budget exceeded");
constexpr uint32_t TM_W3 = 0, TM_ACC3 = 288, TM_ACC2 = 384;   // tensor-memory columns: W3 | conv3 acc (96) | conv2 acc 2 x 64

struct CnnIgDev {
  const float* frames;
  __nv_bfloat16* pooled;
  const uint32_t* w1; const float* b1;
  const uint8_t* w2_slabs; const float* b2;
  const uint4* w3_rows;  // [128 ch][576 k] bf16, k = tap * 64 + c
  const float* b3;
  int M, num_groups;
  int dbg;               // timing ablations (AVSEP_CNN_DBG): 1 skip conv1 stores, 2 skip conv1 loads, 4 skip conv1 MMAs
  long long* trace;      // optional [grid][64] clock64 stamps of the CTA's second group (debug): builder thread 0 in
                         // [0,32), MMA thread in [32,64)
};
#define ITRACE(slot) do { if (p.trace != nullptr && trace_on && tbase < 64) p.trace[blockIdx.x * 64 + tbase + (slot)] = clock64(); } while (0)

__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// Shared-memory descriptor, K-major, no swizzle: core matrices of 8 rows x 16 B; LBO = byte distance between the two
// 16-byte K chunks of one MMA, SBO = byte distance between consecutive 8-row groups (tools/umma_noswz_test.cu).
__device__ __forceinline__ uint64_t desc_noswz(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}

__global__ void __launch_bounds__(IG_THREADS, 1) visual_cnn_ig_kernel(const CnnIgDev p) {
  extern __shared__ __align__(1024) uint8_t smem_raw_ig[];
  uint8_t* const smem = smem_raw_ig;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* w2s = smem + OFF_W2;
  uint8_t* act1 = smem + OFF_ACT1;
  uint8_t* act2 = smem + OFF_ACT2;
  float* sIn = reinterpret_cast<float*>(smem + OFF_IN);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_MISC);
  uint64_t* act1_ready = bars;          // [2] (16) conv1 of the frame written
  uint64_t* acc2_full = bars + 2;       // [2]  conv2 MMAs of the frame complete (its act1 buffer may be overwritten)
  uint64_t* acc2_empty = bars + 4;      // [2] (16) conv2 epilogue drained the accumulator
  uint64_t* act2_ready = bars + 6;      // (16) act2 of the group complete
  uint64_t* acc3_full = bars + 7;       // conv3 MMAs of the group complete
  uint64_t* acc3_empty = bars + 8;      // (16) conv3 epilogue drained the accumulator
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < 5 * 8192 / 16; i += IG_THREADS)
    reinterpret_cast<uint4*>(w2s)[i] = reinterpret_cast<const uint4*>(p.w2_slabs)[i];
  // the zero borders of the parity planes (and of the staged frame) are written once: later stores touch only interiors
  for (int i = tid; i < (OFF_MISC - OFF_ACT1) / 16; i += IG_THREADS)
    reinterpret_cast<uint4*>(act1)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&act1_ready[i], 8);
      mbar_init(&acc2_full[i], 1);
      mbar_init(&acc2_empty[i], 8);
    }
    mbar_init(act2_ready, 8);
    mbar_init(acc3_full, 1);
    mbar_init(acc3_empty, 8);
    fence_mbar_init();
  }
  if (warp == 16) tmem_alloc(tmem_slot, 512);
  fence_proxy_async_smem();     // generic stores above are read by the tensor core
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // conv3 weights -> tensor memory, once: lane = output channel, column c holds the bf16 pair k = 2c, 2c + 1
  if (warp < 4) {
    const uint4* src = p.w3_rows + static_cast<size_t>(warp * 32 + lane) * 72;        // 576 bf16 = 72 x 16 B
    const uint32_t dst = tmem_base + TM_W3 + (static_cast<uint32_t>(warp * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < 9; ++c) {
      uint32_t v[32];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint4 w = __ldg(src + c * 8 + i);
        v[4 * i] = w.x; v[4 * i + 1] = w.y; v[4 * i + 2] = w.z; v[4 * i + 3] = w.w;
      }
      tmem_st_32x32b_x32(dst + c * 32, v);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  griddep_launch_dependents();
  griddep_wait();

  const int n_my_groups = (p.num_groups - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int n_my_frames = n_my_groups * GROUP;

  if (warp == 16) {
    // =========================== MMA issuer (one lane) ===========================
    if (lane == 0 && !(p.dbg & 16)) {
      constexpr uint32_t ID2 = umma_idesc(1u, 128, 64);
      constexpr uint32_t ID3 = umma_idesc(1u, 128, 96);
      const uint32_t a1 = smem_u32(act1), a2 = smem_u32(act2);
      constexpr int tbase = 32;
      uint32_t n3 = 0;
      const uint64_t w2desc = umma_desc_kmajor_sw128(smem_u32(w2s), 1024);     // slab j at + j * 8192 B = + j * 512 units
      auto conv3 = [&](uint32_t g) {            // g = CTA-local group index
        const bool trace_on = g == 1;
        ITRACE(8);
        mbar_wait(act2_ready, g & 1);
        mbar_wait(acc3_empty, (n3 & 1) ^ 1);
        tc_fence_after();
        ITRACE(9);                              // act2 seen, accumulator free
        // one descriptor per buffer; every (tap, k-step) view is that descriptor plus a constant in its address field
        const uint64_t d0 = desc_noswz(a2 + (g & 1) * ACT2_BYTES + PAD2 * 16, ROWS2 * 16, 128);
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int ky = tap / 3, kx = tap % 3;
          const int plane = (ky != 1 ? 2 : 0) + (kx != 1 ? 1 : 0);
          const int delta = (ky == 0 ? -PITCH2 : 0) + (kx == 0 ? -1 : 0);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const int off16 = (plane * 8 + 2 * ks) * ROWS2 + delta;      // in 16-byte units
            umma_f16_ts(tmem_base + TM_ACC3, tmem_base + TM_W3 + tap * 32 + ks * 8,
                        d0 + static_cast<uint64_t>(static_cast<int64_t>(off16)), ID3, (tap | ks) != 0 ? 1u : 0u);
          }
        }
        umma_commit(acc3_full);
        ITRACE(10);                             // conv3 issued
        ++n3;
      };
      for (int i = 0; i < n_my_frames; ++i) {
        const uint32_t b = i & 1;
        const bool trace_on = (i / GROUP) == 1;
        const int ts = (i % GROUP) * 2;
        mbar_wait(&act1_ready[b], (i >> 1) & 1);
        mbar_wait(&acc2_empty[b], ((i >> 1) & 1) ^ 1);
        tc_fence_after();
        ITRACE(ts);                             // act1 of the frame seen
        const uint64_t d0 = desc_noswz(a1 + b * ACT1_BYTES + PAD1 * 16, ROWS1 * 16, 128);
        const uint32_t d_t = tmem_base + TM_ACC2 + b * 64;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int ky = tap / 3, kx = tap % 3;
          const int plane = (ky != 1 ? 2 : 0) + (kx != 1 ? 1 : 0);
          const int delta = (ky == 0 ? -PITCH1 : 0) + (kx == 0 ? -1 : 0);
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const int off16 = (plane * 4 + 2 * ks) * ROWS1 + delta;
            umma_f16(d_t, d0 + static_cast<uint64_t>(static_cast<int64_t>(off16)),
                     w2desc + ((tap >> 1) * 512 + ((tap & 1) * 2 + ks) * 2), ID2, (tap | ks) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&acc2_full[b]);
        ITRACE(ts + 1);                         // conv2 issued
        // conv3 of the previous group goes after the first conv2 of this one, so that the CUDA-core warps (which are
        // waiting for that conv2 to release an act1 buffer) are never stalled behind it
        if (i % GROUP == 0 && i > 0) conv3(static_cast<uint32_t>(i / GROUP - 1));
      }
      if (n_my_groups > 0) conv3(static_cast<uint32_t>(n_my_groups - 1));
    }
  } else if (warp < 8) {
    // =========================== group A (8 warps): frame staging + conv1 ===========================
    const int gid = lane >> 2, tig = lane & 3;
    const int tbase = tid == 0 ? 0 : 4096;
    // conv1 B fragments and bias (constant, registers)
    // The folded BatchNorm bias rides in the unused k = 9 row of the weight fragment (the A fragment carries a constant 1
    // there), so the epilogue is a packed ReLU only: the CUDA-core warps are instruction-issue bound.
    uint32_t bw1[4][2];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      bw1[nt][0] = __ldg(p.w1 + nt * 64 + lane * 2);
      bw1[nt][1] = __ldg(p.w1 + nt * 64 + lane * 2 + 1);
      if (tig == 0) {
        const __nv_bfloat16 bb = __float2bfloat16(__ldg(p.b1 + nt * 8 + gid));
        bw1[nt][1] = (bw1[nt][1] & 0xffffu) | (static_cast<uint32_t>(__bfloat16_as_ushort(bb)) << 16);
      }
    }
    const int k0 = 2 * tig, k1 = 2 * tig + 1;
    const int off0 = (k0 / 3) * 34 + (k0 % 3), off1 = (k1 / 3) * 34 + (k1 % 3), off8 = 2 * 34 + 2;
    // conv1: this warp's two m-tiles of 16 output pixels r = t * 16 + gid (+ 8): source offset in sIn, destination in act1
    int c1_src[2][2], c1_dst[2][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int r = (warp + 8 * mt) * 16 + gid + 8 * hf;
        const int Y = r >> 4, X = r & 15;
        c1_src[mt][hf] = (2 * Y) * 34 + 2 * X;
        const int plane = (Y & 1) * 2 + (X & 1);
        c1_dst[mt][hf] = (plane * 4 * ROWS1 + PAD1 + ((Y >> 1) + 1) * PITCH1 + (X >> 1) + 1) * 16 + tig * 4;
      }
    // Input prefetch: the load is unconditional (clamped address) and its result is not touched until the next slot's
    // staging, so that its latency (a frame comes from HBM) hides behind this slot's conv1; validity is a separate flag
    float4 pre;
    bool pre_valid;
    auto prefetch_frame = [&](int local_i) {
      const int li = local_i < n_my_frames ? local_i : 0;
      const int fr = (static_cast<int>(blockIdx.x) + (li / GROUP) * static_cast<int>(gridDim.x)) * GROUP + li % GROUP;
      pre_valid = local_i < n_my_frames && fr < p.M;
      pre = __ldg(reinterpret_cast<const float4*>(p.frames + static_cast<size_t>(fr < p.M ? fr : 0) * 1024) + tid);
    };
    prefetch_frame(0);
    for (int i = 0; i < n_my_frames; ++i) {
      const bool trace_on = (i / GROUP) == 1;
      const int ts = (i % GROUP) * 8;
      ITRACE(ts);                                     // slot start
      float* sin = sIn + (i & 1) * (34 * 34);
      // ---- stage the frame (interior of the zero-bordered 34x34 tile) from the prefetch registers ----
      {
        const int y = tid >> 3, x4 = (tid & 7) * 4;
        float* d = sin + (y + 1) * 34 + (x4 + 1);
        d[0] = pre_valid ? pre.x : 0.f; d[1] = pre_valid ? pre.y : 0.f;
        d[2] = pre_valid ? pre.z : 0.f; d[3] = pre_valid ? pre.w : 0.f;
      }
      ITRACE(ts + 3);                                 // staging stores issued
      named_bar_sync(1, 256);       // staged; also: every warp of the group has finished conv1 of frame i - 1 (other tile)
      ITRACE(ts + 4);                                 // barrier passed
      prefetch_frame(i + 1);
      // ---- this frame's act1 buffer was last read by the conv2 MMAs of frame i - 2 ----
      if (i >= 2 && !(p.dbg & 16)) {
        mbar_wait(&acc2_full[i & 1], ((i - 2) >> 1) & 1);
        tc_fence_after();
      }
      ITRACE(ts + 1);                                 // staged, act1 buffer free
      // ---- conv1: 256 output pixels, K = 9 (padded to 16), N = 32 -> act1 parity planes; two m-tiles per warp ----
      uint8_t* abuf = act1 + (i & 1) * ACT1_BYTES;
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        uint32_t a[4] = {0, 0, 0, 0};
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const float* base = sin + c1_src[mt][hf];
          if (!(p.dbg & 2)) {
            a[hf] = pack_bf16x2(base[off0], base[off1]);
            if (tig == 0) a[2 + hf] = pack_bf16x2(base[off8], 1.0f);      // k = 8: tap 8; k = 9: constant 1 (x bias row)
          }
        }
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          float cacc[4] = {0.f, 0.f, 0.f, 0.f};
          if (!(p.dbg & 4)) mma16816(cacc, a[0], a[1], a[2], a[3], bw1[nt][0], bw1[nt][1]);
          if (!(p.dbg & 1)) {
            const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
            __nv_bfloat162 lo = __hmax2(__floats2bfloat162_rn(cacc[0], cacc[1]), zero2);
            __nv_bfloat162 hi = __hmax2(__floats2bfloat162_rn(cacc[2], cacc[3]), zero2);
            *reinterpret_cast<__nv_bfloat162*>(abuf + c1_dst[mt][0] + nt * ROWS1 * 16) = lo;
            *reinterpret_cast<__nv_bfloat162*>(abuf + c1_dst[mt][1] + nt * ROWS1 * 16) = hi;
          } else if (cacc[0] + cacc[1] + cacc[2] + cacc[3] == 12345.f) {
            abuf[0] = 1;
          }
        }
      }
      ITRACE(ts + 5);                                 // conv1 stores issued
      fence_proxy_async_smem();
      ITRACE(ts + 6);                                 // proxy fence done
      __syncwarp();
      if (lane == 0) mbar_arrive(&act1_ready[i & 1]);
      ITRACE(ts + 2);                                 // conv1 done
    }
  } else {
    // =========================== group B (8 warps): conv2 and conv3 epilogues ===========================
    const int q = warp & 3, part = (warp - 8) >> 2;
    const uint32_t lane_sel = static_cast<uint32_t>(q * 32) << 16;
    const int tbase = tid == 256 ? 48 : 4096;
    // conv2 epilogue: this thread's raster row (TMEM lane) -> act2 destination inside a (plane, chunk) array
    int e2_dst;             // byte offset without the buffer / frame-slot terms; < 0: border or junk row
    {
      const int m = q * 32 + lane;
      const int yy = m / PITCH1, xx = m - yy * PITCH1;
      const bool ok = m < FR1 && yy >= 1 && xx >= 1;
      const int y = yy - 1, x = xx - 1;
      const int plane = (y & 1) * 2 + (x & 1);
      e2_dst = ok ? ((plane * 8 + part * 4) * ROWS2 + PAD2 + ((y >> 1) + 1) * PITCH2 + (x >> 1) + 1) * 16 : -1;
    }
    const float b3v = __ldg(p.b3 + q * 32 + lane);          // conv3 epilogue: thread = output channel

    // conv2 epilogue of CTA-local frame j: accumulator buffer j & 1 -> bias + ReLU -> act2 planes of the frame's group
    auto conv2_epilogue = [&](int j) {
      const uint32_t b = j & 1;
      mbar_wait_sleep(&acc2_full[b], (j >> 1) & 1);       // (suspended, not spinning: group A shares the issue slots)
      tc_fence_after();
      if (q < 3) {                               // raster rows 96..127 of the tile are junk: nothing to read
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + TM_ACC2 + b * 64 + lane_sel + part * 32, v);
        tmem_ld_wait();
        if (e2_dst >= 0) {
          uint8_t* dst = act2 + ((j / GROUP) & 1) * ACT2_BYTES + (j % GROUP) * FR2 * 16 + e2_dst;
          const float* bb = p.b2 + part * 32;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float4 ba = __ldg(reinterpret_cast<const float4*>(bb + c * 8));
            const float4 bc = __ldg(reinterpret_cast<const float4*>(bb + c * 8 + 4));
            uint4 u;
            u.x = pack_bf16x2(fmaxf(__uint_as_float(v[c * 8 + 0]) + ba.x, 0.f), fmaxf(__uint_as_float(v[c * 8 + 1]) + ba.y, 0.f));
            u.y = pack_bf16x2(fmaxf(__uint_as_float(v[c * 8 + 2]) + ba.z, 0.f), fmaxf(__uint_as_float(v[c * 8 + 3]) + ba.w, 0.f));
            u.z = pack_bf16x2(fmaxf(__uint_as_float(v[c * 8 + 4]) + bc.x, 0.f), fmaxf(__uint_as_float(v[c * 8 + 5]) + bc.y, 0.f));
            u.w = pack_bf16x2(fmaxf(__uint_as_float(v[c * 8 + 6]) + bc.z, 0.f), fmaxf(__uint_as_float(v[c * 8 + 7]) + bc.w, 0.f));
            *reinterpret_cast<uint4*>(dst + c * ROWS2 * 16) = u;
          }
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&acc2_empty[b]);
        if (j % GROUP == GROUP - 1) mbar_arrive(act2_ready);
      }
    };
    // conv3 epilogue of CTA-local group g: thread = output channel q * 32 + lane; part 0 takes frame slots 0 and 1, part 1
    // slot 2; bias + ReLU + mean over the frame's 16 valid raster rows = columns (y+1) * 5 + (x+1), y, x in 0..3
    auto conv3_epilogue = [&](int g) {
      mbar_wait_sleep(acc3_full, g & 1);
      tc_fence_after();
      const int s0 = part == 0 ? 0 : 2, s1 = part == 0 ? 2 : 3;
      for (int sl = s0; sl < s1; ++sl) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + TM_ACC3 + lane_sel + sl * FR2, v);
        tmem_ld_wait();
        float s = 0.f;
#pragma unroll
        for (int y = 1; y <= 4; ++y)
#pragma unroll
          for (int x = 1; x <= 4; ++x) s += fmaxf(__uint_as_float(v[y * PITCH2 + x]) + b3v, 0.f);
        const int fr = (static_cast<int>(blockIdx.x) + g * static_cast<int>(gridDim.x)) * GROUP + sl;
        if (fr < p.M) p.pooled[static_cast<size_t>(fr) * 128 + q * 32 + lane] = __float2bfloat16(s * (1.f / 16.f));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc3_empty);
    };
    for (int j = 0; j < ((p.dbg & 16) ? 0 : n_my_frames); ++j) {
      const bool trace_on = (j / GROUP) == 1;
      const int ts = (j % GROUP) * 4;
      ITRACE(ts);                                     // waiting for the frame's conv2 accumulator
      conv2_epilogue(j);
      ITRACE(ts + 1);                                 // conv2 epilogue done
      // the previous group's conv3 (issued after the first conv2 of this group) has had two frame slots to complete
      if (j % GROUP == GROUP - 1 && j >= GROUP) conv3_epilogue(j / GROUP - 1);
      ITRACE(ts + 2);
    }
    if (n_my_groups > 0 && !(p.dbg & 16)) conv3_epilogue(n_my_groups - 1);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 16) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

inline uint16_t f2bf_ig(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return static_cast<uint16_t>((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return static_cast<uint16_t>(u >> 16);
}

}  // namespace

size_t visual_cnn_ig_w3_bytes() { return 128 * 576 * 2; }

// w3 [128][9*64] fp32, BN-folded, K index = tap*64 + c -> the same matrix in bf16 (row = output channel): each row is
// copied verbatim into one lane of tensor memory, where it serves as the A operand of the transposed conv3 GEMM
void visual_cnn_ig_pack(const float* w3, uint8_t* w3_rows) {
  uint16_t* o = reinterpret_cast<uint16_t*>(w3_rows);
  for (int i = 0; i < 128 * 576; ++i) o[i] = f2bf_ig(w3[i]);
}

const char* launch_visual_cnn_ig(cudaStream_t s, const float* frames, int M, const CnnWeights& w, const uint8_t* w2_slabs,
                                 const uint8_t* w3_rows, void* pooled, int num_sms, long long* trace) {
  if (M <= 0) return "visual_cnn_ig: empty problem";
  if ((reinterpret_cast<uintptr_t>(w3_rows) & 15) != 0) return "visual_cnn_ig: weights not 16-byte aligned";
  static bool attr_done = false;
  if (!attr_done) {
    if (cudaFuncSetAttribute(visual_cnn_ig_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, IG_SMEM) != cudaSuccess)
      return "visual_cnn_ig: cudaFuncSetAttribute failed";
    attr_done = true;
  }
  CnnIgDev d;
  d.frames = frames; d.pooled = reinterpret_cast<__nv_bfloat16*>(pooled);
  d.w1 = w.w1; d.b1 = w.b1; d.w2_slabs = w2_slabs; d.b2 = w.b2; d.w3_rows = reinterpret_cast<const uint4*>(w3_rows); d.b3 = w.b3;
  d.M = M; d.num_groups = (M + GROUP - 1) / GROUP;
  d.trace = trace;
  static int dbg = -1;
  if (dbg < 0) { const char* e = getenv("AVSEP_CNN_DBG"); dbg = e ? atoi(e) : 0; }
  d.dbg = dbg;
  const int grid = d.num_groups < num_sms ? d.num_groups : num_sms;
  launch_pdl(visual_cnn_ig_kernel, dim3(grid), dim3(IG_THREADS), IG_SMEM, s, d);
  return cudaGetLastError() == cudaSuccess ? nullptr : "visual_cnn_ig: launch failed";
}

}  // namespace avsep
