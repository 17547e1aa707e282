// Onset encoder on the GPU: audio_encoders_pytorch.Encoder1d as configured at
// /root/reference/exp/model/diffusion.yaml:35-43 and called at main/generation.py:71 / main/module_diffusion.py:196
// (`_, y_latent = model.onsets_encoder(y, with_info=True)`), whose info['xs'][2:-1] pyramid is the sampling path's
// `channels` input.  It runs ONCE per batch, outside the sampling loop: 17 ResNet blocks + 8 strided convolutions with
// 2 ... 256 channels, ~1 GFLOP and ~60 MB per clip - pure bandwidth / latency, < 1 % of one sampling call.  So this is a
// small set of fp32 CUDA-core streaming kernels in the reference's own NCL layout (its outputs are handed to
// sfb_sample unchanged), each fusing what the reference runs as 3-5 passes:
//   enc_conv_kernel   GroupNorm apply + SiLU on the INPUT (statistics from the producer's epilogue, fp64 sums) ->
//                     Conv1d(k = 1 / 3 / 2f+1, stride f, zero padding) -> + bias -> + residual (identity or 1x1 shortcut
//                     conv) -> store + GroupNorm statistics of the OUTPUT for the next block (fp64 atomics per clip, group)
//   enc_stats_kernel  statistics of the raw onset track (the only tensor no kernel of ours produced)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sfb {

struct EncConvParams {
  const float* in;          // [B, Cin, Lin]
  const float* w;           // [Cout, Cin, K]
  const float* bias;        // [Cout]
  float* out;               // [B, Cout, Lout]
  const double* gn_stats;   // [B, G, 2] (sum, sum of squares) of `in`, or null: no input transform
  const float* gn_w;        // [Cin] GroupNorm affine
  const float* gn_b;
  const float* res;         // [B, Cres, Lout] residual source or null
  const float* sc_w;        // [Cout, Cres] 1x1 shortcut conv or null (identity residual: Cres == Cout)
  const float* sc_b;        // [Cout]
  double* out_stats;        // [B, Gout, 2] or null
  int G, Gout, Cin, Cout, Cres, Lin, Lout, K, stride, pad;
  float eps;
};

constexpr int kEncMaxC = 512;

// grid (ceil(Lout / 128), Cout, B), 128 threads: one output sample per thread.
__global__ void __launch_bounds__(128) enc_conv_kernel(const EncConvParams p) {
  __shared__ float sa[kEncMaxC], sb[kEncMaxC];     // per input channel: y = silu(x * a + b)
  __shared__ double red[2][4];
  const int b = blockIdx.z, co = blockIdx.y;
  const int l = blockIdx.x * 128 + threadIdx.x;
  if (p.gn_stats != nullptr) {
    const int gs = p.Cin / p.G;
    const double inv_cnt = 1.0 / ((double)gs * p.Lin);
    for (int ci = threadIdx.x; ci < p.Cin; ci += 128) {
      const int g = ci / gs;
      const double s1 = p.gn_stats[((size_t)b * p.G + g) * 2], s2 = p.gn_stats[((size_t)b * p.G + g) * 2 + 1];
      const double mean = s1 * inv_cnt;
      const double var = fmax(s2 * inv_cnt - mean * mean, 0.0);
      const float a = (float)(1.0 / sqrt(var + (double)p.eps)) * p.gn_w[ci];
      sa[ci] = a;
      sb[ci] = p.gn_b[ci] - (float)mean * a;
    }
    __syncthreads();
  }
  float acc = 0.f;
  const bool valid = l < p.Lout;
  if (valid) {
    const float* inb = p.in + (size_t)b * p.Cin * p.Lin;
    const float* wr = p.w + (size_t)co * p.Cin * p.K;
    const int pos0 = l * p.stride - p.pad;
    for (int ci = 0; ci < p.Cin; ++ci) {
      const float* row = inb + (size_t)ci * p.Lin;
      const float a = p.gn_stats ? sa[ci] : 1.f, c = p.gn_stats ? sb[ci] : 0.f;
      for (int k = 0; k < p.K; ++k) {
        const int pos = pos0 + k;
        if (pos < 0 || pos >= p.Lin) continue;                 // zero padding applies AFTER the activation
        float x = __ldg(row + pos);
        if (p.gn_stats) {
          x = fmaf(x, a, c);
          x = x / (1.f + expf(-x));                            // SiLU
        }
        acc = fmaf(__ldg(wr + ci * p.K + k), x, acc);
      }
    }
    acc += p.bias[co];
    if (p.res != nullptr) {
      const float* rb = p.res + (size_t)b * p.Cres * p.Lout;
      if (p.sc_w != nullptr) {
        float s = p.sc_b[co];
        for (int cr = 0; cr < p.Cres; ++cr) s = fmaf(p.sc_w[(size_t)co * p.Cres + cr], __ldg(rb + (size_t)cr * p.Lout + l), s);
        acc += s;
      } else {
        acc += __ldg(rb + (size_t)co * p.Lout + l);
      }
    }
    p.out[((size_t)b * p.Cout + co) * p.Lout + l] = acc;
  }
  if (p.out_stats != nullptr) {
    double s1 = valid ? (double)acc : 0.0, s2 = valid ? (double)acc * acc : 0.0;
    for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s1; red[1][threadIdx.x >> 5] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
      const int g = co / (p.Cout / p.Gout);
      atomicAdd(&p.out_stats[((size_t)b * p.Gout + g) * 2], red[0][0] + red[0][1] + red[0][2] + red[0][3]);
      atomicAdd(&p.out_stats[((size_t)b * p.Gout + g) * 2 + 1], red[1][0] + red[1][1] + red[1][2] + red[1][3]);
    }
  }
}

// statistics of a raw [B, C, L] tensor: grid (ceil(L / 1024), C, B), 256 threads
__global__ void __launch_bounds__(256) enc_stats_kernel(const float* __restrict__ in, double* __restrict__ stats, int C, int L, int G) {
  __shared__ double red[2][8];
  const int b = blockIdx.z, c = blockIdx.y;
  const float* row = in + ((size_t)b * C + c) * L;
  double s1 = 0.0, s2 = 0.0;
  for (int l = blockIdx.x * 1024 + threadIdx.x; l < min(L, (int)(blockIdx.x + 1) * 1024); l += 256) {
    const double v = row[l];
    s1 += v; s2 += v * v;
  }
  for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s1; red[1][threadIdx.x >> 5] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, q = 0.0;
    for (int i = 0; i < 8; ++i) { a += red[0][i]; q += red[1][i]; }
    const int g = c / (C / G);
    atomicAdd(&stats[((size_t)b * G + g) * 2], a);
    atomicAdd(&stats[((size_t)b * G + g) * 2 + 1], q);
  }
}

}  // namespace sfb
