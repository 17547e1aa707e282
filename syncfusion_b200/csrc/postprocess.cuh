// Generation post-processing: /root/reference/main/generation.py:85-98 for a whole batch in ONE pass
// over the generated waveforms, so that only the cropped, 22.05 kHz result crosses PCIe:
//   cut_prefix : samples before the clip's first onset are zeroed        (:87-88, first_onset = nonzero(y[i][0])[0])
//   crop       : gen[i, :, :cut_length]                                   (:90, :96)
//   resample   : torchaudio.functional.resample(orig_freq -> new_freq)    (:90-92) - windowed-sinc polyphase FIR, Hann
//                window, lowpass_filter_width 6, rolloff 0.99; out[f * new + p] = sum_k table[p][k] x[f * orig + k - width]
// HBM-bound streaming kernel (4 B read + 1.8 B written per input sample; 348 MACs per output from shared memory): a CTA
// stages the input span of FPB output frames in shared memory once (mask and crop applied while staging) and every
// thread produces outputs from it; the [new x K] tap table (205 KB for 48 k -> 22.05 k) is read through the read-only
// cache, adjacent threads reading adjacent phases.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sfb {

// first non-zero sample of every onset track: out[b] = min n with y[b, n] != 0 (L if none).  out must be pre-set to L.
__global__ void __launch_bounds__(256) first_onset_kernel(const float* __restrict__ y, int* __restrict__ out, int L) {
  const int b = blockIdx.y;
  const int n0 = (blockIdx.x * 256 + threadIdx.x) * 4;
  if (n0 >= L) return;
  const float* row = y + (size_t)b * L;
  int first = L;
  if (n0 + 3 < L && (L & 3) == 0) {        // rows are 16-byte aligned only when L is a multiple of 4
    const float4 v = *reinterpret_cast<const float4*>(row + n0);
    first = v.x != 0.f ? n0 : v.y != 0.f ? n0 + 1 : v.z != 0.f ? n0 + 2 : v.w != 0.f ? n0 + 3 : L;
  } else {
    for (int n = n0; n < min(n0 + 4, L); ++n) if (row[n] != 0.f) { first = n; break; }
  }
  if (first < L) atomicMin(&out[b], first);
}
__global__ void fill_int_kernel(int* p, int n, int v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

struct PostParams {
  const float* gen;         // [B, L]
  const int* first_onset;   // [B] or null (no prefix cut)
  const float* table;       // [nw][K] or null (no resampling)
  float* out;               // [B, T]
  int L, cut, orig, nw, K, width, T, fpb;
};

// grid (ceil(frames / fpb), B), 256 threads, dynamic smem ((fpb - 1) * orig + K) floats
__global__ void __launch_bounds__(256) postprocess_resample_kernel(const PostParams p) {
  extern __shared__ float xs[];
  const int b = blockIdx.y;
  const int f0 = blockIdx.x * p.fpb;
  const int span = (p.fpb - 1) * p.orig + p.K;
  const int first = p.first_onset ? p.first_onset[b] : 0;
  const float* g = p.gen + (size_t)b * p.L;
  const int base = f0 * p.orig - p.width;
  for (int i = threadIdx.x; i < span; i += 256) {
    const int n = base + i;
    xs[i] = (n >= first && n >= 0 && n < p.cut) ? __ldg(g + n) : 0.f;     // zero padding, prefix mask and crop in one place
  }
  __syncthreads();
  float* o = p.out + (size_t)b * p.T;
  for (int j = threadIdx.x; j < p.fpb * p.nw; j += 256) {
    const int f = j / p.nw, ph = j - f * p.nw;
    const int t = (f0 + f) * p.nw + ph;
    if (t >= p.T) continue;
    const float* tab = p.table + (size_t)ph * p.K;
    const float* x = xs + f * p.orig;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int k = 0;
    for (; k + 3 < p.K; k += 4) {
      a0 = fmaf(__ldg(tab + k), x[k], a0); a1 = fmaf(__ldg(tab + k + 1), x[k + 1], a1);
      a2 = fmaf(__ldg(tab + k + 2), x[k + 2], a2); a3 = fmaf(__ldg(tab + k + 3), x[k + 3], a3);
    }
    for (; k < p.K; ++k) a0 = fmaf(__ldg(tab + k), x[k], a0);
    o[t] = (a0 + a1) + (a2 + a3);
  }
}

// no resampling: masked crop copy.  grid (ceil(T / 1024), B)
__global__ void __launch_bounds__(256) postprocess_copy_kernel(const PostParams p) {
  const int b = blockIdx.y;
  const int first = p.first_onset ? p.first_onset[b] : 0;
  for (int n = (blockIdx.x * 256 + threadIdx.x); n < p.T; n += gridDim.x * 256)
    p.out[(size_t)b * p.T + n] = n >= first ? p.gen[(size_t)b * p.L + n] : 0.f;
}

}  // namespace sfb
