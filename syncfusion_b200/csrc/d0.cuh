// Depth-0 (C = 8, L = full waveform length) streaming kernels.  At 8 channels the contractions have an arithmetic
// intensity of ~12 flop/B: they are HBM-bound, a 128 x N UMMA tile cannot be filled, so these run on
// the CUDA cores with one waveform position per thread, 128-bit accesses and fp32 math:
//   d0_down   : Down_0 = Conv1d(1 -> 8, k = 1) on the raw waveform (a11) (+ GN stats)
//   conv3_c8  : ResNet Conv1d(8 -> 8, k = 3, p = 1) (a6), epilogue A: +bias -> operand dtype (+ GN stats),
//               epilogue B: +bias +residual -> fp32 stream
//   inject_c8 : InjectChannels Conv1x1(cat[x, ctx]) + x (+ cross-attention bias) (a8, a10)
//   d0_up     : Up_0 (nearest x1 + conv3 8 -> 1, or transpose k = 1) fused with SkipModulate (a12, a5) -> v
#pragma once
#include "elementwise.cuh"

namespace sfb {

// per-channel (sum, sum of squares) of 8 channels over the block's 256 positions, accumulated in fp64: GroupNorm at
// depth 0 is a per-channel instance norm, and E[x^2] - mean^2 cancels catastrophically in fp32 whenever a channel is
// nearly constant (|mean| >> std).  Values go through smem; 64 threads each reduce 32 rows of one channel in double.
struct Stats8Smem {
  float y[256 * 9];
  double red[16];
};
__device__ __forceinline__ void stats8_block_reduce(const float* y, bool valid, double* stats_b, Stats8Smem& sm) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int j = 0; j < 8; ++j) sm.y[tid * 9 + j] = valid ? y[j] : 0.f;
  if (tid < 16) sm.red[tid] = 0.0;
  __syncthreads();
  if (tid < 64) {
    const int c = tid & 7, seg = tid >> 3;
    double a = 0.0, q = 0.0;
#pragma unroll 8
    for (int i = 0; i < 32; ++i) {
      const double v = (double)sm.y[(i * 8 + seg) * 9 + c];
      a += v;
      q += v * v;
    }
    atomicAdd(&sm.red[c * 2], a);
    atomicAdd(&sm.red[c * 2 + 1], q);
  }
  __syncthreads();
  if (tid < 16) atomicAdd(&stats_b[tid], sm.red[tid]);
}

// x [Bx, L] f32 (clip b % Bx) -> y [B, L, 8] f32 ; w [8], bias [8]
__global__ void __launch_bounds__(256) d0_down_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                      const float* __restrict__ bias, float* __restrict__ y,
                                                      double* __restrict__ stats, int L, int Bx) {
  pdl_trigger();
  pdl_wait();
  __shared__ Stats8Smem s_st;
  const int b = blockIdx.y;
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = l < L;
  float o[8];
  if (valid) {
    const float xv = x[(size_t)(b % Bx) * L + l];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = xv * __ldg(&w[j]) + __ldg(&bias[j]);
    Vec8<float>::store(y + ((size_t)b * L + l) * 8, o);
  }
  if (stats) stats8_block_reduce(o, valid, stats + (size_t)b * 16, s_st);
}

// in [B, L, 8] (T) ; w [24][8] f32 (k = tap * 8 + ci, co fastest) ; bias [8]
// mode A: out_t = acc + bias (+stats) ; mode B: out_r = acc + bias + resid
template <typename T>
__global__ void __launch_bounds__(256) conv3_c8_kernel(const T* __restrict__ in, const float* __restrict__ w,
                                                       const float* __restrict__ bias, const float* resid, float* out_r,
                                                       T* __restrict__ out_t, double* __restrict__ stats, int L) {
  pdl_trigger();
  pdl_wait();
  __shared__ float s_w[24 * 8];
  __shared__ Stats8Smem s_st;
  if (threadIdx.x < 192) s_w[threadIdx.x] = w[threadIdx.x];
  __syncthreads();
  const int b = blockIdx.y;
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = l < L;
  float acc[8];
  if (valid) {
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = __ldg(&bias[j]);
    const T* base = in + (size_t)b * L * 8;
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      const int ll = l + t - 1;
      if (ll < 0 || ll >= L) continue;
      float xv[8];
      Vec8<T>::load(base + (size_t)ll * 8, xv);
#pragma unroll
      for (int ci = 0; ci < 8; ++ci) {
        const float4 w0 = *reinterpret_cast<const float4*>(&s_w[(t * 8 + ci) * 8]);
        const float4 w1 = *reinterpret_cast<const float4*>(&s_w[(t * 8 + ci) * 8 + 4]);
        acc[0] += xv[ci] * w0.x; acc[1] += xv[ci] * w0.y; acc[2] += xv[ci] * w0.z; acc[3] += xv[ci] * w0.w;
        acc[4] += xv[ci] * w1.x; acc[5] += xv[ci] * w1.y; acc[6] += xv[ci] * w1.z; acc[7] += xv[ci] * w1.w;
      }
    }
    const size_t g = ((size_t)b * L + l) * 8;
    if (resid) {
      float rv[8];
      Vec8<float>::load(resid + g, rv);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += rv[j];
    }
    if (out_r) Vec8<float>::store(out_r + g, acc);
    if (out_t) store_operand8<T>(out_t + g, acc);
  }
  if (stats) stats8_block_reduce(acc, valid, stats + (size_t)b * 16, s_st);
}

// m_t [B, L, 8] (T operand copy), m_r [B, L, 8] f32 (residual), ctx [Bc, L, CTX] (T), w [8 + CTX][8] f32 (k, co), bias [8],
// xbias [B, 8] or null -> out = W [m, ctx] + bias + m_r + xbias
template <typename T, int CTX>
__global__ void __launch_bounds__(256) inject_c8_kernel(const T* __restrict__ m_t, const float* m_r, const T* __restrict__ ctx,
                                                        const float* __restrict__ w, const float* __restrict__ bias,
                                                        const float* __restrict__ xbias, float* out_r, T* __restrict__ out_t,
                                                        double* __restrict__ stats, int L, int Bc, int xb_stride) {
  pdl_trigger();
  pdl_wait();
  __shared__ float s_w[(8 + CTX) * 8];
  __shared__ Stats8Smem s_st;
  for (int i = threadIdx.x; i < (8 + CTX) * 8; i += blockDim.x) s_w[i] = w[i];
  __syncthreads();
  const int b = blockIdx.y;
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = l < L;
  float acc[8];
  if (valid) {
    const size_t g = ((size_t)b * L + l) * 8;
    float xv[8], rv[8];
    Vec8<T>::load(m_t + g, xv);
    Vec8<float>::load(m_r + g, rv);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = __ldg(&bias[j]) + rv[j] + (xbias ? __ldg(&xbias[(size_t)b * xb_stride + j]) : 0.f);
#pragma unroll
    for (int ci = 0; ci < 8; ++ci)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += xv[ci] * s_w[ci * 8 + j];
    const T* cp = ctx + ((size_t)(b % Bc) * L + l) * CTX;
#pragma unroll
    for (int ci = 0; ci < CTX; ++ci) {
      const float cv = to_f32(cp[ci]);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += cv * s_w[(8 + ci) * 8 + j];
    }
    if (out_r) Vec8<float>::store(out_r + g, acc);
    if (out_t) store_operand8<T>(out_t + g, acc);
  }
  if (stats) stats8_block_reduce(acc, valid, stats + (size_t)b * 16, s_st);
}

// c [B, L, 8] (T) ; w [taps][8] f32 ; v[b, l] = x[b % Bx, l] + s[b % smod] * (sum w c + bias)
template <typename T>
__global__ void __launch_bounds__(256) d0_up_kernel(const T* __restrict__ c, const float* __restrict__ w, float bias,
                                                    const float* __restrict__ skip_scale, int sstride, int smod,
                                                    const float* __restrict__ x, float* __restrict__ v, int L, int Bx,
                                                    int taps) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.y;
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= L) return;
  const int pad = taps == 3 ? 1 : 0;
  float acc = bias;
  const T* base = c + (size_t)b * L * 8;
  for (int t = 0; t < taps; ++t) {
    const int ll = l + t - pad;
    if (ll < 0 || ll >= L) continue;
    float xv[8];
    Vec8<T>::load(base + (size_t)ll * 8, xv);
#pragma unroll
    for (int ci = 0; ci < 8; ++ci) acc += xv[ci] * __ldg(&w[t * 8 + ci]);
  }
  const float s = __ldg(&skip_scale[(size_t)(b % smod) * sstride]);
  v[(size_t)b * L + l] = x[(size_t)(b % Bx) * L + l] + s * acc;
}

}  // namespace sfb
