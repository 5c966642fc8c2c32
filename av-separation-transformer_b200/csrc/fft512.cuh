// 512-point complex transform held in registers: 64 threads x 8 values, three radix-8 passes (decimation in
// frequency), two exchanges through shared memory instead of the nine read-modify-write sweeps of a radix-2
// transform - the shared-memory wavefront count is what bounds the STFT kernels.
//
//   n = t + 64 q                 pass 1: 8-point DFT over q  -> p,   times W_512^(t p)
//   t = t2 + 8 q'                pass 2: 8-point DFT over q' -> p2,  times W_64^(t2 p2)
//                                pass 3: 8-point DFT over t2 -> p3;  k = p + 8 p2 + 64 p3
//
// On return thread t64 = 8 p + p2 holds X[(t64 >> 3) + 8 (t64 & 7) + 64 i] in v[i].  SGN = -1 forward, +1 inverse
// (unscaled).  The exchange buffers use a pitch of 72 floats per 64-point sub-transform and 9 per 8-point one, which
// makes every access below bank-conflict free.  All 64-thread groups of the CTA must call this together (it uses
// __syncthreads).
#pragma once
#include <cuda_runtime.h>

namespace avsep {

constexpr int FFT512_XCH = 8 * 72;   // floats per exchange plane (re or im) per 64-thread group

template <int SGN>
__device__ __forceinline__ float2 mul_i(float2 z) {           // z * (SGN i)
  return make_float2(-SGN * z.y, SGN * z.x);
}

template <int SGN>
__device__ __forceinline__ void dft4(float2 b0, float2 b1, float2 b2, float2 b3, float2& y0, float2& y1, float2& y2,
                                     float2& y3) {
  const float2 s0 = make_float2(b0.x + b2.x, b0.y + b2.y), s1 = make_float2(b0.x - b2.x, b0.y - b2.y);
  const float2 s2 = make_float2(b1.x + b3.x, b1.y + b3.y), s3 = mul_i<SGN>(make_float2(b1.x - b3.x, b1.y - b3.y));
  y0 = make_float2(s0.x + s2.x, s0.y + s2.y);
  y2 = make_float2(s0.x - s2.x, s0.y - s2.y);
  y1 = make_float2(s1.x + s3.x, s1.y + s3.y);
  y3 = make_float2(s1.x - s3.x, s1.y - s3.y);
}

template <int SGN>
__device__ __forceinline__ void dft8(float2 (&a)[8]) {
  constexpr float R = 0.70710678118654752440f;
  // two 4-point transforms (even / odd inputs)
  float2 e[4], o[4];
  dft4<SGN>(a[0], a[2], a[4], a[6], e[0], e[1], e[2], e[3]);
  dft4<SGN>(a[1], a[3], a[5], a[7], o[0], o[1], o[2], o[3]);
  // W_8^1 = (1 + SGN i)/sqrt2, W_8^2 = SGN i, W_8^3 = (-1 + SGN i)/sqrt2
  const float2 o1 = make_float2(R * (o[1].x - SGN * o[1].y), R * (o[1].y + SGN * o[1].x));
  const float2 o2 = mul_i<SGN>(o[2]);
  const float2 o3 = make_float2(R * (-o[3].x - SGN * o[3].y), R * (-o[3].y + SGN * o[3].x));
  a[0] = make_float2(e[0].x + o[0].x, e[0].y + o[0].y);
  a[4] = make_float2(e[0].x - o[0].x, e[0].y - o[0].y);
  a[1] = make_float2(e[1].x + o1.x, e[1].y + o1.y);
  a[5] = make_float2(e[1].x - o1.x, e[1].y - o1.y);
  a[2] = make_float2(e[2].x + o2.x, e[2].y + o2.y);
  a[6] = make_float2(e[2].x - o2.x, e[2].y - o2.y);
  a[3] = make_float2(e[3].x + o3.x, e[3].y + o3.y);
  a[7] = make_float2(e[3].x - o3.x, e[3].y - o3.y);
}

// Twiddle tables in shared memory, laid out the way the passes read them (conflict free): pass 1 reads
// tw[p * 64 + t] = W_512^(t p) with consecutive t; pass 2 reads tw[512 + t2 * 9 + p2] = W_64^(t2 p2) (8 distinct
// entries per access, pitch 18 words).  Stored as (cos, sin) of the positive angle.
constexpr int FFT512_TW = 512 + 72;   // float2 entries

__device__ __forceinline__ void fft512_fill_twiddles(float2* tw, int tid, int nthreads) {
  for (int m = tid; m < FFT512_TW; m += nthreads) {
    int num;                                        // angle = 2 pi num / 512
    if (m < 512) num = (m >> 6) * (m & 63);
    else { const int r = m - 512; num = 8 * (r / 9) * (r % 9); }
    float sv, cv;
    sincospif(static_cast<float>(num & 511) / 256.0f, &sv, &cv);
    tw[m] = make_float2(cv, sv);
  }
}

// Natural-order index k -> padded position (4 floats of padding per 32): the register layout after the last pass
// (k = hi + 8 lo + 64 i) then lands on 32 distinct banks, and consecutive k stay conflict free.  < FFT512_XCH.
__device__ __forceinline__ int fft512_nat(int k) { return k + ((k >> 5) << 2); }

template <int SGN>
__device__ __forceinline__ float2 twiddle_mul(float2 z, float2 w) {   // z * (w.x + SGN i w.y)
  return make_float2(z.x * w.x - SGN * z.y * w.y, z.y * w.x + SGN * z.x * w.y);
}

// v[q] = x[t + 64 q] on entry (t = thread within the group, 0 .. 63); xr / xi: the group's exchange planes.
template <int SGN>
__device__ __forceinline__ void fft512_regs(float2 (&v)[8], int t, float* xr, float* xi, const float2* tw) {
  dft8<SGN>(v);
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    const float2 z = p == 0 ? v[0] : twiddle_mul<SGN>(v[p], tw[p * 64 + t]);
    xr[p * 72 + t] = z.x; xi[p * 72 + t] = z.y;
  }
  __syncthreads();
  const int hi = t >> 3, lo = t & 7;      // pass 2: sub-transform p = hi, t2 = lo
#pragma unroll
  for (int q = 0; q < 8; ++q) v[q] = make_float2(xr[hi * 72 + lo + 8 * q], xi[hi * 72 + lo + 8 * q]);
  __syncthreads();
  dft8<SGN>(v);
#pragma unroll
  for (int p2 = 0; p2 < 8; ++p2) {
    const float2 z = p2 == 0 ? v[0] : twiddle_mul<SGN>(v[p2], tw[512 + lo * 9 + p2]);
    xr[hi * 72 + p2 * 9 + lo] = z.x; xi[hi * 72 + p2 * 9 + lo] = z.y;
  }
  __syncthreads();
  // pass 3: sub-transform (p = hi, p2 = lo), its 8 points are contiguous (pitch 9 between sub-transforms)
#pragma unroll
  for (int q = 0; q < 8; ++q) v[q] = make_float2(xr[hi * 72 + lo * 9 + q], xi[hi * 72 + lo * 9 + q]);
  dft8<SGN>(v);
}

// index held in v[i] after fft512_regs
__device__ __forceinline__ int fft512_out_index(int t, int i) { return (t >> 3) + 8 * (t & 7) + 64 * i; }

}  // namespace avsep
