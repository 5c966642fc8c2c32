// The waveform side of the path: the reference stops at magnitude spectrograms and lists
// "STFT phase / iSTFT reconstruction" as missing (reference README.md:140), so a separated utterance never becomes
// audio there.  These two kernels close that gap on the device with the reference's own analysis framing:
//
//   stft_complex_kernel   the framing of SyntheticAVDataset._stft (dataset.py:122-135: frame i starts at i*hop, no
//                         centring, zero-padded past the end, np.hanning window, rfft) keeping the phase: complex64
//                         (B, F, T); |.| is bit-identical to the magnitudes of synth.cu's stft_mag_kernel.
//   istft_masked_kernel   y[b,s,n] = sum_i w[n-i*hop] * irfft(mask[b,s,:,i] * X[b,:,i])[n-i*hop] / sum_i w^2[n-i*hop]
//                         (weighted overlap-add, the least-squares inverse of that analysis; samples no window
//                         reaches with non-zero weight - n = 0 for a Hann window - come out 0).
//
// One CTA owns a run of consecutive output samples and transforms every frame that overlaps it (no atomics, results
// independent of the launch geometry); two Hermitian spectra share one complex inverse transform (z = A + iB ->
// ifft(z) = a + ib for real a, b).  fp32 shared-memory radix-2 transforms, fp64 window: HBM/latency-bound CUDA-core
// work, no tensor cores.
#include "common.cuh"
#include "kernels.h"

namespace avsep {

namespace {

constexpr double kPi = 3.141592653589793238462643383279502884;
constexpr int TGA = 8;   // analysis frames per CTA

__device__ __forceinline__ void fft_stages(float* re, float* im, const float* twr, const float* twi, int log2n,
                                           int tid) {
  for (int s = 1; s <= log2n; ++s) {
    __syncthreads();
    const int half = 1 << (s - 1);
    const int pos = tid & (half - 1);
    const int i0 = ((tid >> (s - 1)) << s) + pos;
    const int i1 = i0 + half;
    const int tw = pos << (log2n - s);
    const float wr = twr[tw], wi = twi[tw];
    const float xr = re[i1], xi = im[i1];
    const float pr = wr * xr - wi * xi, pi = wr * xi + wi * xr;
    const float ar = re[i0], ai = im[i0];
    re[i0] = ar + pr; im[i0] = ai + pi;
    re[i1] = ar - pr; im[i1] = ai - pi;
  }
  __syncthreads();
}

// np.hanning(M)[k] = 0.5 + 0.5*cos(pi*(1-M+2k)/(M-1)), fp64 (dataset.py:124)
__device__ __forceinline__ double hann(int k, int nfft) {
  return 0.5 + 0.5 * cos(kPi * static_cast<double>(1 - nfft + 2 * k) / static_cast<double>(nfft - 1));
}

// grid (ceil(T/TGA), B), nfft/2 threads.  spec: (B, F, T) float2; mag: (B, F, T) float or null.
__global__ void stft_complex_kernel(const float* __restrict__ waves, int n, int nfft, int log2n, int hop, int T, int F,
                                    float2* __restrict__ spec, float* __restrict__ mag) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  double* win = reinterpret_cast<double*>(sm_raw);          // [nfft]
  float* re = reinterpret_cast<float*>(win + nfft);         // [nfft]
  float* im = re + nfft;                                    // [nfft]
  float* twr = im + nfft;                                   // [nfft/2]
  float* twi = twr + nfft / 2;                              // [nfft/2]
  float2* stage = reinterpret_cast<float2*>(twi + nfft / 2);  // [F][TGA]
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * TGA;
  const int tid = threadIdx.x;
  const float* x = waves + static_cast<size_t>(b) * n;
  {
    float sv, cv;
    sincospif(-2.0f * static_cast<float>(tid) / static_cast<float>(nfft), &sv, &cv);
    twr[tid] = cv; twi[tid] = sv;
    win[tid] = hann(tid, nfft);
    win[tid + nfft / 2] = hann(tid + nfft / 2, nfft);
  }
  const int ng = min(TGA, T - t0);
  for (int g = 0; g < ng; g += 2) {
    const int ta = t0 + g;
    const bool has_b = g + 1 < ng;
    __syncthreads();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int k = tid + h * (nfft / 2);
      const int sa = ta * hop + k, sb = sa + hop;
      const float va = sa < n ? x[sa] : 0.f;
      const float vb = (has_b && sb < n) ? x[sb] : 0.f;
      const double w = win[k];
      const int r = static_cast<int>(__brev(static_cast<unsigned>(k)) >> (32 - log2n));
      re[r] = static_cast<float>(static_cast<double>(va) * w);     // float32 `frame *= window`, dataset.py:131
      im[r] = static_cast<float>(static_cast<double>(vb) * w);
    }
    fft_stages(re, im, twr, twi, log2n, tid);
    for (int f = tid; f < F; f += blockDim.x) {
      const int fn = (nfft - f) & (nfft - 1);
      const float zr = re[f], zi = im[f], yr = re[fn], yi = im[fn];
      stage[f * TGA + g] = make_float2(0.5f * (zr + yr), 0.5f * (zi - yi));                 // A = (Z + conj Z~)/2
      if (has_b) stage[f * TGA + g + 1] = make_float2(0.5f * (zi + yi), 0.5f * (yr - zr));  // B = (Z - conj Z~)/2i
    }
  }
  __syncthreads();
  for (int e = tid; e < F * TGA; e += blockDim.x) {
    const int f = e / TGA, g = e - f * TGA;
    if (g >= ng) continue;
    const size_t o = (static_cast<size_t>(b) * F + f) * T + t0 + g;
    const float2 v = stage[e];
    spec[o] = v;
    // same expression as stft_mag_kernel (0.5 * hypot of the unscaled sums) so the magnitudes agree bit for bit
    if (mag) mag[o] = 0.5f * hypotf(2.0f * v.x, 2.0f * v.y);
  }
}

// grid (ceil(L/CH), S, B), nfft/2 threads.  spec (B, F, T) float2, masks (B, S, F, T) float or null (S = 1),
// out (B, S, L).  CH = samples per CTA (a multiple of hop).  The frames that overlap the CTA's samples are processed
// in batches of NB: the masked half spectra of a batch are staged in shared memory with row-contiguous loads (the
// (F, T) layout puts the frames of one bin side by side, so a per-frame gather would touch one sector per element),
// then transformed two at a time.
constexpr int NB = 12;        // frames per staged batch (even)
constexpr int NBP = NB + 1;   // row pitch in float2: 26 words, conflict-free for 64-bit accesses down a column

__global__ void istft_masked_kernel(const float2* __restrict__ spec, const float* __restrict__ masks, int S, int nfft,
                                    int log2n, int hop, int T, int F, int L, int CH, float* __restrict__ out) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  double* win = reinterpret_cast<double*>(sm_raw);          // [nfft]
  double* acc = win + nfft;                                 // [CH]
  float2* tile = reinterpret_cast<float2*>(acc + CH);       // [F][NBP]
  float* re = reinterpret_cast<float*>(tile + static_cast<size_t>(F) * NBP);   // [nfft]
  float* im = re + nfft;                                    // [nfft]
  float* twr = im + nfft;                                   // [nfft/2]
  float* twi = twr + nfft / 2;                              // [nfft/2]
  const int b = blockIdx.z, s = blockIdx.y;
  const int n0 = blockIdx.x * CH;
  const int n1 = min(n0 + CH, L);
  const int tid = threadIdx.x;
  {
    float sv, cv;
    sincospif(2.0f * static_cast<float>(tid) / static_cast<float>(nfft), &sv, &cv);   // inverse: conj twiddles
    twr[tid] = cv; twi[tid] = sv;
    win[tid] = hann(tid, nfft);
    win[tid + nfft / 2] = hann(tid + nfft / 2, nfft);
  }
  for (int i = tid; i < CH; i += blockDim.x) acc[i] = 0.0;
  // frames overlapping [n0, n1): i*hop <= n1-1 and i*hop + nfft - 1 >= n0
  const int first = n0 - nfft + 1;
  const int i_lo = first > 0 ? (first + hop - 1) / hop : 0;
  const int i_hi = min(T - 1, (n1 - 1) / hop);
  const float2* X = spec + static_cast<size_t>(b) * F * T;
  const float* Mk = masks ? masks + (static_cast<size_t>(b) * S + s) * F * T : nullptr;
  const float inv_n = 1.0f / static_cast<float>(nfft);
  for (int ib = i_lo; ib <= i_hi; ib += NB) {
    const int nb = min(NB, i_hi - ib + 1);
    __syncthreads();
    for (int e = tid; e < F * NB; e += blockDim.x) {
      const int f = e / NB, j = e - f * NB;
      float2 v = make_float2(0.f, 0.f);
      if (j < nb) {
        const size_t o = static_cast<size_t>(f) * T + ib + j;
        v = X[o];
        if (Mk) { const float m = Mk[o]; v.x *= m; v.y *= m; }
        if (f == 0 || f == nfft / 2) v.y = 0.f;            // irfft ignores the imaginary part of DC and Nyquist
      }
      tile[f * NBP + j] = v;
    }
    for (int j = 0; j < nb; j += 2) {
      const int ia = ib + j;
      const bool has_b = j + 1 < nb;
      __syncthreads();
      // Z[k] = A[k] + i B[k] with A, B the Hermitian extensions of the two half spectra (frame b is zero when
      // absent); bit-reversed placement for the in-place transform.
      for (int k = tid; k < nfft; k += blockDim.x) {
        const int f = k <= nfft / 2 ? k : nfft - k;
        const float sgn = k <= nfft / 2 ? 1.f : -1.f;
        const float2 a = tile[f * NBP + j];
        const float2 bb = tile[f * NBP + j + 1];           // column NB is never written past nb: j + 1 <= NB - 1
        const int r = static_cast<int>(__brev(static_cast<unsigned>(k)) >> (32 - log2n));
        re[r] = a.x - sgn * bb.y;
        im[r] = sgn * a.y + bb.x;
      }
      fft_stages(re, im, twr, twi, log2n, tid);
      // overlap-add frame a, then frame b (a barrier between: the two frames hit the same samples hop apart)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k = tid + h * (nfft / 2);
        const int nn = ia * hop + k - n0;
        if (nn >= 0 && nn < n1 - n0) acc[nn] += win[k] * static_cast<double>(re[k] * inv_n);
      }
      if (has_b) {
        __syncthreads();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int k = tid + h * (nfft / 2);
          const int nn = (ia + 1) * hop + k - n0;
          if (nn >= 0 && nn < n1 - n0) acc[nn] += win[k] * static_cast<double>(im[k] * inv_n);
        }
      }
    }
  }
  __syncthreads();
  float* y = out + (static_cast<size_t>(b) * S + s) * L;
  for (int i = tid; i < n1 - n0; i += blockDim.x) {
    const int n = n0 + i;
    const int fl = n - nfft + 1;
    double wss = 0.0;
    for (int fr = fl > 0 ? (fl + hop - 1) / hop : 0; fr <= min(T - 1, n / hop); ++fr) {
      const double w = win[n - fr * hop];
      wss += w * w;
    }
    y[n] = wss > 1e-11 ? static_cast<float>(acc[i] / wss) : 0.f;
  }
}

bool fft_geometry_ok(int nfft, int hop, int* log2n) {
  if (nfft < 8 || nfft > 2048 || (nfft & (nfft - 1)) || hop < 1 || hop > nfft) return false;
  *log2n = 0;
  while ((1 << *log2n) < nfft) ++*log2n;
  return true;
}

}  // namespace

const char* launch_stft_complex(cudaStream_t s, const float* waves, int B, int L, int nfft, int hop, float* spec,
                                float* mag) {
  int log2n;
  if (B <= 0 || L <= 0) return "stft: empty problem";
  if (!fft_geometry_ok(nfft, hop, &log2n)) return "stft: n_fft must be a power of two in [8, 2048], 1 <= hop <= n_fft";
  const int T = 1 + L / hop, F = nfft / 2 + 1;
  const size_t smem = static_cast<size_t>(nfft) * sizeof(double) + 3 * static_cast<size_t>(nfft) * sizeof(float) +
                      static_cast<size_t>(F) * TGA * sizeof(float2);
  static size_t granted = 48 * 1024;
  if (smem > granted) {
    if (cudaFuncSetAttribute(stft_complex_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return "stft: cannot raise the shared-memory limit";
    granted = 227 * 1024;
  }
  stft_complex_kernel<<<dim3((T + TGA - 1) / TGA, B), nfft / 2, smem, s>>>(waves, L, nfft, log2n, hop, T, F,
                                                                          reinterpret_cast<float2*>(spec), mag);
  return cudaGetLastError() == cudaSuccess ? nullptr : "stft: launch failed";
}

const char* launch_istft_masked(cudaStream_t s, const float* spec, const float* masks, int B, int S, int T, int nfft,
                                int hop, int L, float* waves) {
  int log2n;
  if (B <= 0 || S <= 0 || T <= 0 || L <= 0) return "istft: empty problem";
  if (!fft_geometry_ok(nfft, hop, &log2n)) return "istft: n_fft must be a power of two in [8, 2048], 1 <= hop <= n_fft";
  if (S > 65535 || B > 65535) return "istft: batch or speaker count above 65535";
  if (masks == nullptr && S != 1) return "istft: without masks there is one output per utterance (S = 1)";
  if (static_cast<long long>(T - 1) * hop + nfft < L) return "istft: the frames do not reach the requested length";
  const int F = nfft / 2 + 1;
  int tg = 4096 / hop;
  tg = tg < 1 ? 1 : (tg > 8 ? 8 : tg);
  const int CH = tg * hop;
  const size_t smem = (static_cast<size_t>(nfft) + CH) * sizeof(double) + static_cast<size_t>(F) * NBP * sizeof(float2) +
                      3 * static_cast<size_t>(nfft) * sizeof(float);
  if (smem > 227 * 1024) return "istft: geometry needs more than 227 KB of shared memory";
  static size_t granted = 48 * 1024;
  if (smem > granted) {
    if (cudaFuncSetAttribute(istft_masked_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return "istft: cannot raise the shared-memory limit";
    granted = 227 * 1024;
  }
  istft_masked_kernel<<<dim3((L + CH - 1) / CH, S, B), nfft / 2, smem, s>>>(
      reinterpret_cast<const float2*>(spec), masks, S, nfft, log2n, hop, T, F, L, CH, waves);
  return cudaGetLastError() == cudaSuccess ? nullptr : "istft: launch failed";
}

}  // namespace avsep
