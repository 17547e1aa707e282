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

// x [Bx, L] f32 (clip b % Bx) -> y [B, L, 8] f32 ; w [8], bias [8].  One position per thread (a warp stores 1 KB
// contiguous; four positions per thread put 32 different lines behind every store instruction and ran 10 % slower).
// The GroupNorm sums of y = w x + b follow from the block's (sum x, sum x^2, count) in fp64 - sum y = w sx + b n,
// sum y^2 = w^2 sxx + 2 w b sx + b^2 n - so the block reduces TWO values instead of sixteen.
__global__ void __launch_bounds__(256) d0_down_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                      const float* __restrict__ bias, float* __restrict__ y,
                                                      double* __restrict__ stats, int L, int Bx) {
  pdl_trigger();
  __shared__ double s_red[8][2];
  __shared__ int s_cnt[8];
  float wv[8], bv[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { wv[j] = __ldg(&w[j]); bv[j] = __ldg(&bias[j]); }
  pdl_wait();
  const int b = blockIdx.y;
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  float xv = 0.f;
  const bool valid = l < L;
  if (valid) {
    xv = x[(size_t)(b % Bx) * L + l];
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = xv * wv[j] + bv[j];
    Vec8<float>::store(y + ((size_t)b * L + l) * 8, o);
  }
  if (stats) {
    double a = (double)xv, q2 = (double)xv * (double)xv;
    int c = valid ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); q2 += __shfl_xor_sync(0xffffffffu, q2, o); c += __shfl_xor_sync(0xffffffffu, c, o); }
    if ((threadIdx.x & 31) == 0) { s_red[threadIdx.x >> 5][0] = a; s_red[threadIdx.x >> 5][1] = q2; s_cnt[threadIdx.x >> 5] = c; }
    __syncthreads();
    if (threadIdx.x < 16) {
      double bsx = 0.0, bsxx = 0.0, n = 0.0;
#pragma unroll
      for (int k = 0; k < 8; ++k) { bsx += s_red[k][0]; bsxx += s_red[k][1]; n += (double)s_cnt[k]; }
      const int ch = threadIdx.x >> 1;
      const double wc = (double)__ldg(&w[ch]), bc = (double)__ldg(&bias[ch]);
      const double v = (threadIdx.x & 1) ? wc * wc * bsxx + 2.0 * wc * bc * bsx + bc * bc * n : wc * bsx + bc * n;
      atomicAdd(&stats[(size_t)b * 16 + threadIdx.x], v);
    }
  }
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

// c [B, L, 8] (T) ; w [taps][8] f32 ; v[b, l] = x[b % Bx, l] + s[b % smod] * (sum w c + bias).  Four positions per thread
// (L % 4 == 0): the six rows l - 1 .. l + 4 are loaded once and shared by the four outputs, x and v move as 16-byte vectors.
template <typename T>
__global__ void __launch_bounds__(256) d0_up_kernel(const T* __restrict__ c, const float* __restrict__ w, float bias,
                                                    const float* __restrict__ skip_scale, int sstride, int smod,
                                                    const float* __restrict__ x, float* __restrict__ v, int L, int Bx,
                                                    int taps) {
  pdl_trigger();
  float wv[3][8];
#pragma unroll
  for (int t = 0; t < 3; ++t)
#pragma unroll
    for (int ci = 0; ci < 8; ++ci) wv[t][ci] = t < taps ? __ldg(&w[t * 8 + ci]) : 0.f;
  pdl_wait();
  const int b = blockIdx.y;
  const int l = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (l >= L) return;
  const int pad = taps == 3 ? 1 : 0;
  const T* base = c + (size_t)b * L * 8;
  float rows[6][8];                       // positions l - pad .. l - pad + 3 + (taps - 1)
  const int nrows = 4 + taps - 1;
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    const int ll = l - pad + r;
    if (r < nrows && ll >= 0 && ll < L) Vec8<T>::load(base + (size_t)ll * 8, rows[r]);
    else {
#pragma unroll
      for (int ci = 0; ci < 8; ++ci) rows[r][ci] = 0.f;
    }
  }
  const float s = __ldg(&skip_scale[(size_t)(b % smod) * sstride]);
  const float4 xq = *reinterpret_cast<const float4*>(x + (size_t)(b % Bx) * L + l);
  const float xv[4] = {xq.x, xq.y, xq.z, xq.w};
  float out[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float acc = bias;
#pragma unroll
    for (int t = 0; t < 3; ++t)
#pragma unroll
      for (int ci = 0; ci < 8; ++ci) acc += rows[q + t][ci] * wv[t][ci];     // taps == 1: wv[1], wv[2] are zero and pad == 0
    out[q] = xv[q] + s * acc;
  }
  *reinterpret_cast<float4*>(v + (size_t)b * L + l) = make_float4(out[0], out[1], out[2], out[3]);
}


// ---------------------------------------------------------------------------------------------------- fused depth-0 item
// The six reference ops of a depth-0 item (GN1+SiLU, conv1, GN2+SiLU, conv2 + x, Modulation, InjectChannels) in TWO
// launches - the GroupNorm statistics of conv1's output are the only full-length dependency inside the item:
//   d0_gn_conv1 : h = conv3(SiLU(GN1(x))) + b1                        x fp32 -> h operand dtype (+ GN2 sums)
//   d0_tail     : y = Inject(Mod(LN(conv3(SiLU(GN2(h))) + b2 + x)))   h, x -> y fp32 in place (+ operand copy, GN sums)
// Every tensor crosses HBM once per launch (700 -> ~400 MB per item at B = 16); activations between the fused steps
// stay in fp32 registers / shared memory, so the result is at least as close to the fp32 reference as the unfused path.
//
// Tiling: a block of 256 threads covers kD0Rows = 768 consecutive positions base - 1 .. base + 766 (three per thread,
// all loads issued up front), activates them into shared memory (row pitch 12 floats: conflict-free 128-bit
// accesses) and produces the 766 outputs base .. base + 765 - the conv halo costs no separate loads.  The conv keeps
// the three positions of a thread in registers so every weight vector read from shared memory feeds 24 FMAs (the
// kernels are instruction-issue bound, not HBM bound: ncu 54 % issue slots busy at 1.4 TB/s before this layout).
constexpr int kD0Per = 3;
constexpr int kD0Rows = 256 * kD0Per;
constexpr int kD0Pitch = 12;
__host__ __device__ constexpr int d0_positions_per_block() { return kD0Rows - 2; }

struct D0Smem {
  float t[kD0Rows * kD0Pitch];
  float w[24 * 8];
  float ab[16];        // per-channel y = x * a + b of the GroupNorm in front of the conv
  double red[16];
};

__device__ __forceinline__ float silu_fast(float y) { return __fdividef(y, 1.f + __expf(-y)); }

// per-channel GroupNorm coefficients from the fp64 sums (depth 0: 8 groups over 8 channels = instance norm)
__device__ __forceinline__ void d0_coef(D0Smem& sm, const double* stats_b, const float* gamma, const float* beta, int L, float eps) {
  const int tid = threadIdx.x;
  if (tid < 8) {
    const double cnt = (double)L;
    const double mean = stats_b[tid * 2] / cnt;
    double var = stats_b[tid * 2 + 1] / cnt - mean * mean;
    var = var > 0 ? var : 0;
    const float a = (float)(1.0 / sqrt(var + (double)eps)) * gamma[tid];
    sm.ab[tid] = a;
    sm.ab[8 + tid] = beta[tid] - (float)mean * a;
  }
  if (tid < 16) sm.red[tid] = 0.0;
}
__device__ __forceinline__ void d0_act_row(const D0Smem& sm, const float* x, float* row) {
  float t[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) t[j] = silu_fast(x[j] * sm.ab[j] + sm.ab[8 + j]);
  *reinterpret_cast<float4*>(row) = make_float4(t[0], t[1], t[2], t[3]);
  *reinterpret_cast<float4*>(row + 4) = make_float4(t[4], t[5], t[6], t[7]);
}
__device__ __forceinline__ void d0_zero_row(float* row) {
  *reinterpret_cast<float4*>(row) = make_float4(0.f, 0.f, 0.f, 0.f);
  *reinterpret_cast<float4*>(row + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
}
// acc[k][8] += conv3 over the activated rows around tile rows tid + 256 k (k < 3); tile row r is position base - 1 + r.
// Rows are clamped into the tile: the two edge rows (0 and kD0Rows - 1) compute garbage that is never stored.
__device__ __forceinline__ void d0_conv3x3(const D0Smem& sm, int tid, float (*acc)[8]) {
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    float xv[kD0Per][8];
#pragma unroll
    for (int k = 0; k < kD0Per; ++k) {
      int r = tid + 256 * k + t - 1;
      r = r < 0 ? 0 : (r > kD0Rows - 1 ? kD0Rows - 1 : r);
      const float4 x0 = *reinterpret_cast<const float4*>(&sm.t[r * kD0Pitch]);
      const float4 x1 = *reinterpret_cast<const float4*>(&sm.t[r * kD0Pitch + 4]);
      xv[k][0] = x0.x; xv[k][1] = x0.y; xv[k][2] = x0.z; xv[k][3] = x0.w; xv[k][4] = x1.x; xv[k][5] = x1.y; xv[k][6] = x1.z; xv[k][7] = x1.w;
    }
#pragma unroll
    for (int ci = 0; ci < 8; ++ci) {
      const float4 w0 = *reinterpret_cast<const float4*>(&sm.w[(t * 8 + ci) * 8]);
      const float4 w1 = *reinterpret_cast<const float4*>(&sm.w[(t * 8 + ci) * 8 + 4]);
#pragma unroll
      for (int k = 0; k < kD0Per; ++k) {
        acc[k][0] += xv[k][ci] * w0.x; acc[k][1] += xv[k][ci] * w0.y; acc[k][2] += xv[k][ci] * w0.z; acc[k][3] += xv[k][ci] * w0.w;
        acc[k][4] += xv[k][ci] * w1.x; acc[k][5] += xv[k][ci] * w1.y; acc[k][6] += xv[k][ci] * w1.z; acc[k][7] += xv[k][ci] * w1.w;
      }
    }
  }
}
// block-level flush of per-thread (sum, sum of squares) of 8 channels: fp32 over the thread's <= 3 positions, fp64 from
// there on (GroupNorm at depth 0 is a per-channel instance norm; E[x^2] - mean^2 cancels catastrophically in fp32
// whenever a channel is nearly constant, so every cross-position accumulation is done in double).
__device__ __forceinline__ void d0_flush_stats(D0Smem& sm, const float* fa, const float* fq, double* stats_b) {
  double sa[8], sq[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sa[j] = (double)fa[j]; sq[j] = (double)fq[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sa[j] += __shfl_xor_sync(0xffffffffu, sa[j], o); sq[j] += __shfl_xor_sync(0xffffffffu, sq[j], o); }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { atomicAdd(&sm.red[j * 2], sa[j]); atomicAdd(&sm.red[j * 2 + 1], sq[j]); }
  }
  __syncthreads();
  if (threadIdx.x < 16) atomicAdd(&stats_b[threadIdx.x], sm.red[threadIdx.x]);
}

template <typename T> struct RawVec8;
template <> struct RawVec8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) { a = *reinterpret_cast<const float4*>(p); b = *reinterpret_cast<const float4*>(p + 4); }
  __device__ __forceinline__ void get(float* v) const { v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w; }
};
template <> struct RawVec8<__nv_bfloat16> {
  uint4 u;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { u = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void get(float* v) const {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u); }
  }
};

// x [B, L, 8] f32 -> h [B, L, 8] (T) ; w [24][8] f32 (k = tap * 8 + ci, co fastest)
template <typename T>
__global__ void __launch_bounds__(256, 3) d0_gn_conv1_kernel(const float* __restrict__ x, const double* __restrict__ stats_in,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                                             const float* __restrict__ w, const float* __restrict__ bias,
                                                             T* __restrict__ out_t, double* __restrict__ stats_out, int L, float eps) {
  pdl_trigger();
  pdl_wait();
  __shared__ D0Smem sm;
  const int tid = threadIdx.x, b = blockIdx.y;
  const float* xb = x + (size_t)b * L * 8;
  const int base = blockIdx.x * d0_positions_per_block();
  RawVec8<float> raw[kD0Per];
#pragma unroll
  for (int k = 0; k < kD0Per; ++k) {
    const int l = base - 1 + tid + 256 * k;
    if (l >= 0 && l < L) raw[k].load(xb + (size_t)l * 8);
  }
  d0_coef(sm, stats_in + (size_t)b * 16, gamma, beta, L, eps);
  if (tid < 192) sm.w[tid] = w[tid];
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kD0Per; ++k) {
    const int r = tid + 256 * k, l = base - 1 + r;
    if (l >= 0 && l < L) { float xv[8]; raw[k].get(xv); d0_act_row(sm, xv, &sm.t[r * kD0Pitch]); }
    else d0_zero_row(&sm.t[r * kD0Pitch]);
  }
  __syncthreads();
  float acc[kD0Per][8];
#pragma unroll
  for (int k = 0; k < kD0Per; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[k][j] = __ldg(&bias[j]);
  d0_conv3x3(sm, tid, acc);
  float fa[8], fq[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { fa[j] = 0.f; fq[j] = 0.f; }
#pragma unroll
  for (int k = 0; k < kD0Per; ++k) {
    const int r = tid + 256 * k, l = base - 1 + r;
    if (r >= 1 && r <= kD0Rows - 2 && l < L) {
      store_operand8<T>(out_t + ((size_t)b * L + l) * 8, acc[k]);
#pragma unroll
      for (int j = 0; j < 8; ++j) { fa[j] += acc[k][j]; fq[j] = fmaf(acc[k][j], acc[k][j], fq[j]); }
    }
  }
  if (stats_out) d0_flush_stats(sm, fa, fq, stats_out + (size_t)b * 16);
}

// ---------------------------------------------------------------------------------------------------- depth-0 conv on tcgen05
// bf16 mode.  The k = 3, 8 -> 8 channel conv is 192 of the ~500 CUDA-core instructions d0_gn_conv1 spends per position
// (ncu: 56 % issue-slot utilisation at 1.8 TB/s - the kernel is bound by its instruction stream, not by HBM).  Here it
// runs on the tensor core with NO im2col: the activated tile is stored channels-last in shared memory, 16 bytes (8 bf16
// channels) per position, and that dense array IS the K-major, un-swizzled A operand of a K = 16 MMA whose descriptor
// says "the next 16 bytes of K are 16 bytes further on" (LBO = 16 B) - i.e. K chunk 0 of row r is position r, K chunk 1 is
// position r + 1: overlapping rows.  Two MMAs per 128 positions (taps -1 | 0, then +1 | a zero-weight chunk), N = 16
// (8 output channels + 8 zero columns), fp32 accumulators in TMEM, one thread per position in the epilogue.
constexpr int kD0TcTile = 1024;                       // positions per CTA (eight 128-row MMA blocks)
constexpr int kD0TcRows = kD0TcTile + 3;              // rows -1 .. 1025 of the tile: halo + the zero-weight chunk's last row
struct alignas(128) D0TcSmem {
  uint8_t a[(kD0TcRows + 5) * 16];                    // activated bf16 rows, 16 B each (core matrix = 8 consecutive rows)
  uint8_t b[2][512];                                  // per MMA: [k_blk][n_blk] core matrices of 8 (n) x 16 B
  float ab[16];
  double red[16];
  uint64_t bar;
  uint32_t tslot;
};
__device__ __forceinline__ uint32_t d0_pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
// SiLU(2 h) = h + h tanh(h) with ONE MUFU op (tanh.approx.f32: relative error 2^-11, below the bf16 rounding of the result);
// the caller folds the 1/2 into the GroupNorm coefficients.
__device__ __forceinline__ float silu_half(float h) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}
// Sum 16 per-lane doubles over the warp with 16 shuffles instead of 80: at every step a lane keeps half of its values and
// hands the other half to its partner.  Afterwards lane l holds the total of value index
// ((l >> 4) & 1) * 8 + ((l >> 3) & 1) * 4 + ((l >> 2) & 1) * 2 + ((l >> 1) & 1)   (lanes l and l ^ 1 hold the same total).
__device__ __forceinline__ double warp_reduce16(const double (&v)[16], int lane) {
  double w8[8], w4[4], w2[2];
  const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
#pragma unroll
  for (int i = 0; i < 8; ++i) { const double keep = h16 ? v[8 + i] : v[i], send = h16 ? v[i] : v[8 + i]; w8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16); }
#pragma unroll
  for (int i = 0; i < 4; ++i) { const double keep = h8 ? w8[4 + i] : w8[i], send = h8 ? w8[i] : w8[4 + i]; w4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8); }
#pragma unroll
  for (int i = 0; i < 2; ++i) { const double keep = h4 ? w4[2 + i] : w4[i], send = h4 ? w4[i] : w4[2 + i]; w2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4); }
  const double keep = h2 ? w2[1] : w2[0], send = h2 ? w2[0] : w2[1];
  double r = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  r += __shfl_xor_sync(0xffffffffu, r, 1);
  return r;
}
__device__ __forceinline__ unsigned long long f2pack(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2unpack(unsigned long long v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
// y[0..8) += x * w[0..8)   (w: 8 consecutive floats in shared memory, 16-byte aligned; fma.rn.f32x2)
__device__ __forceinline__ void axpy8(unsigned long long (&y)[4], float x, const float* w) {
  const ulonglong2 w0 = *reinterpret_cast<const ulonglong2*>(w), w1 = *reinterpret_cast<const ulonglong2*>(w + 4);
  const unsigned long long xx = f2pack(x, x);
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(y[0]) : "l"(xx), "l"(w0.x));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(y[1]) : "l"(xx), "l"(w0.y));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(y[2]) : "l"(xx), "l"(w1.x));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(y[3]) : "l"(xx), "l"(w1.y));
}
__device__ __forceinline__ void umma_ss_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// x [B, L, 8] f32 -> h [B, L, 8] bf16 ; w [24][8] f32 (k = tap * 8 + ci, co fastest)
__global__ void __launch_bounds__(256, 4) d0_gn_conv1_tc_kernel(const float* __restrict__ x, const double* __restrict__ stats_in,
                                                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                const float* __restrict__ w, const float* __restrict__ bias,
                                                                __nv_bfloat16* __restrict__ out_t, double* __restrict__ stats_out, int L, float eps, int tag) {
  pdl_trigger();
  __shared__ D0TcSmem sm;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, b = blockIdx.y;
  const int base = blockIdx.x * kD0TcTile;
  if (tid == 0) { mbar_init(&sm.bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&sm.tslot, 128); tmem_relinquish(); }
  // weights -> the two B operands: B_m[n][k], n = output channel (8..15: zero), k = k_blk * 8 + ci, tap = 2 m + k_blk (tap 3: zero)
  for (int i = tid; i < 2 * 16 * 16; i += 256) {
    const int m = i >> 8, n = (i >> 4) & 15, k = i & 15;
    const int k_blk = k >> 3, ci = k & 7, tap = 2 * m + k_blk;
    const float v = (n < 8 && tap < 3) ? __ldg(&w[(tap * 8 + ci) * 8 + n]) : 0.f;
    *reinterpret_cast<__nv_bfloat16*>(&sm.b[m][(k_blk * 2 + (n >> 3)) * 128 + (n & 7) * 16 + ci * 2]) = __float2bfloat16_rn(v);
  }
  float bs[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) bs[j] = __ldg(&bias[j]);
  tc_fence_before();
  __syncthreads();
  pdl_wait();
  mark_progress(tag);
  if (tid < 8) {      // per-channel GroupNorm coefficients (depth 0: 8 groups over 8 channels)
    const double cnt = (double)L;
    const double mean = stats_in[(size_t)b * 16 + tid * 2] / cnt;
    double var = stats_in[(size_t)b * 16 + tid * 2 + 1] / cnt - mean * mean;
    var = var > 0 ? var : 0;
    const float a = (float)(1.0 / sqrt(var + (double)eps)) * gamma[tid];
    sm.ab[tid] = 0.5f * a;                                    // h = (GroupNorm(x)) / 2 for silu_half
    sm.ab[8 + tid] = 0.5f * (beta[tid] - (float)mean * a);
  }
  if (tid < 16) sm.red[tid] = 0.0;
  __syncthreads();
  const float* xb = x + (size_t)b * L * 8;
  // phase 1: GroupNorm + SiLU -> bf16 rows (row r = position base - 1 + r); rows outside the clip are the conv's zero padding
  RawVec8<float> raw[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const int r = tid + 256 * k, l = base - 1 + r;
    if (r < kD0TcRows && l >= 0 && l < L) raw[k].load(xb + (size_t)l * 8);
  }
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const int r = tid + 256 * k, l = base - 1 + r;
    if (r < kD0TcRows + 5) {
      uint4 o = make_uint4(0u, 0u, 0u, 0u);
      if (r < kD0TcRows && l >= 0 && l < L) {
        float xv[8], t[8];
        raw[k].get(xv);
#pragma unroll
        for (int j = 0; j < 8; ++j) t[j] = silu_half(fmaf(xv[j], sm.ab[j], sm.ab[8 + j]));
        o = make_uint4(d0_pack_bf16(t[0], t[1]), d0_pack_bf16(t[2], t[3]), d0_pack_bf16(t[4], t[5]), d0_pack_bf16(t[6], t[7]));
      }
      *reinterpret_cast<uint4*>(&sm.a[r * 16]) = o;
    }
  }
  fence_proxy_async();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sm.tslot;
  if (tid == 0) {
    // phase 2: A: K-major, no swizzle, core matrix = 8 rows x 16 B contiguous; next K chunk = next position (+16 B, LBO),
    // next 8 rows = +128 B (SBO).  B: LBO = 256 B between the two K chunks, SBO = 128 B between the two 8-column groups.
    constexpr uint32_t idesc = make_idesc(1 /*bf16*/, 128, 16, 0, 0);
    const uint32_t a0 = smem_u32(sm.a), b0 = smem_u32(sm.b[0]), b1 = smem_u32(sm.b[1]);
#pragma unroll
    for (int mb = 0; mb < kD0TcTile / 128; ++mb) {
      umma_ss_f16(tmem_base + mb * 16, make_smem_desc(a0 + (mb * 128) * 16, 16, 128, 0), make_smem_desc(b0, 256, 128, 0), idesc, 0u);
      umma_ss_f16(tmem_base + mb * 16, make_smem_desc(a0 + (mb * 128 + 2) * 16, 16, 128, 0), make_smem_desc(b1, 256, 128, 0), idesc, 1u);
    }
    umma_commit(&sm.bar);
  }
  mbar_wait_at(&sm.bar, 0, (5u << 16) | __LINE__, (uint32_t)tag);      // bounded, logged wait (ptx.cuh; file id 5 = d0.cuh)
  tc_fence_after();
  // phase 3: one position per thread and MMA block: + bias -> bf16 h, GroupNorm sums of h
  const int q = warp & 3, half = warp >> 2;
  float fa[8], fq[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { fa[j] = 0.f; fq[j] = 0.f; }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int mb = half * 4 + i;
    uint32_t v[16];
    tmem_ld16(tmem_base + (uint32_t(q * 32) << 16) + mb * 16, v);
    tmem_ld_wait();
    const int l = base + mb * 128 + q * 32 + lane;
    if (l < L) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { o[j] = __uint_as_float(v[j]) + bs[j]; fa[j] += o[j]; fq[j] = fmaf(o[j], o[j], fq[j]); }
      Vec8<__nv_bfloat16>::store(out_t + ((size_t)b * L + l) * 8, o);
    }
  }
  if (stats_out) {       // fp32 over the thread's four positions, fp64 from there on (16-shuffle warp reduction, one atomic per value)
    double v16[16];
#pragma unroll
    for (int j = 0; j < 8; ++j) { v16[j] = (double)fa[j]; v16[8 + j] = (double)fq[j]; }
    const double tot = warp_reduce16(v16, lane);
    if ((lane & 1) == 0) {
      const int idx = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
      atomicAdd(&sm.red[(idx & 7) * 2 + (idx >> 3)], tot);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (stats_out && tid < 16) atomicAdd(&stats_out[(size_t)b * 16 + tid], sm.red[tid]);
  if (warp == 0) tmem_dealloc(tmem_base, 128);
}

// Second launch of the depth-0 item on the same scheme: y = Inject(Mod(LN(conv3(SiLU(GN2(h))) + b2 + x))) with the conv on
// tcgen05 (h bf16 -> activated bf16 rows -> two overlapping-row MMAs per 128 positions) and the per-position tail
// (residual, LayerNorm over the 8 channels, Modulation, InjectChannels 10 -> 8) in the one-position-per-thread epilogue.
// The residual / context rows of a thread's four positions are requested BEFORE the accumulator wait.
template <int CTX>
__global__ void __launch_bounds__(256, 3) d0_tail_tc_kernel(const __nv_bfloat16* __restrict__ h, const double* __restrict__ stats_in,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            const float* __restrict__ w, const float* __restrict__ bias,
                                                            const float* x_r, const float* __restrict__ mod, int mod_bstride, int mod_bmod,
                                                            const __nv_bfloat16* __restrict__ ctx, int Bc, const float* __restrict__ wi,
                                                            const float* __restrict__ bi, const float* __restrict__ xbias, int xb_stride,
                                                            float* out_r, __nv_bfloat16* __restrict__ out_t, double* __restrict__ stats_out, int L,
                                                            float eps, int tag) {
  pdl_trigger();
  __shared__ D0TcSmem sm;
  __shared__ __align__(16) float s_wi[(8 + CTX) * 8];
  __shared__ float s_v[4 * 8];     // conv bias | 1 + mod scale | mod shift | inject bias + cross-attention bias
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, b = blockIdx.y;
  const int base = blockIdx.x * kD0TcTile;
  if (tid == 0) { mbar_init(&sm.bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&sm.tslot, 128); tmem_relinquish(); }
  for (int i = tid; i < 2 * 16 * 16; i += 256) {
    const int m = i >> 8, n = (i >> 4) & 15, k = i & 15;
    const int k_blk = k >> 3, ci = k & 7, tap = 2 * m + k_blk;
    const float v = (n < 8 && tap < 3) ? __ldg(&w[(tap * 8 + ci) * 8 + n]) : 0.f;
    *reinterpret_cast<__nv_bfloat16*>(&sm.b[m][(k_blk * 2 + (n >> 3)) * 128 + (n & 7) * 16 + ci * 2]) = __float2bfloat16_rn(v);
  }
  for (int i = tid; i < (8 + CTX) * 8; i += 256) s_wi[i] = wi[i];
  tc_fence_before();
  __syncthreads();
  pdl_wait();
  mark_progress(tag);
  if (tid < 8) {
    const double cnt = (double)L;
    const double mean = stats_in[(size_t)b * 16 + tid * 2] / cnt;
    double var = stats_in[(size_t)b * 16 + tid * 2 + 1] / cnt - mean * mean;
    var = var > 0 ? var : 0;
    const float a = (float)(1.0 / sqrt(var + (double)eps)) * gamma[tid];
    sm.ab[tid] = 0.5f * a;
    sm.ab[8 + tid] = 0.5f * (beta[tid] - (float)mean * a);
    const float* md = mod + (size_t)(b % mod_bmod) * mod_bstride;
    s_v[tid] = bias[tid];
    s_v[8 + tid] = 1.f + md[tid];
    s_v[16 + tid] = md[8 + tid];
    s_v[24 + tid] = bi[tid] + (xbias ? xbias[(size_t)b * xb_stride + tid] : 0.f);
  }
  if (tid < 16) sm.red[tid] = 0.0;
  __syncthreads();
  const __nv_bfloat16* hb = h + (size_t)b * L * 8;
  RawVec8<__nv_bfloat16> raw[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const int r = tid + 256 * k, l = base - 1 + r;
    if (r < kD0TcRows && l >= 0 && l < L) raw[k].load(hb + (size_t)l * 8);
  }
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const int r = tid + 256 * k, l = base - 1 + r;
    if (r < kD0TcRows + 5) {
      uint4 o = make_uint4(0u, 0u, 0u, 0u);
      if (r < kD0TcRows && l >= 0 && l < L) {
        float xv[8], t[8];
        raw[k].get(xv);
#pragma unroll
        for (int j = 0; j < 8; ++j) t[j] = silu_half(fmaf(xv[j], sm.ab[j], sm.ab[8 + j]));
        o = make_uint4(d0_pack_bf16(t[0], t[1]), d0_pack_bf16(t[2], t[3]), d0_pack_bf16(t[4], t[5]), d0_pack_bf16(t[6], t[7]));
      }
      *reinterpret_cast<uint4*>(&sm.a[r * 16]) = o;
    }
  }
  fence_proxy_async();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sm.tslot;
  if (tid == 0) {
    constexpr uint32_t idesc = make_idesc(1 /*bf16*/, 128, 16, 0, 0);
    const uint32_t a0 = smem_u32(sm.a), b0 = smem_u32(sm.b[0]), b1 = smem_u32(sm.b[1]);
#pragma unroll
    for (int mb = 0; mb < kD0TcTile / 128; ++mb) {
      umma_ss_f16(tmem_base + mb * 16, make_smem_desc(a0 + (mb * 128) * 16, 16, 128, 0), make_smem_desc(b0, 256, 128, 0), idesc, 0u);
      umma_ss_f16(tmem_base + mb * 16, make_smem_desc(a0 + (mb * 128 + 2) * 16, 16, 128, 0), make_smem_desc(b1, 256, 128, 0), idesc, 1u);
    }
    umma_commit(&sm.bar);
  }
  // the residual and context rows of this thread's four positions, in flight while the MMAs run
  const int q = warp & 3, half = warp >> 2;
  const float* xr = x_r + (size_t)b * L * 8;
  const __nv_bfloat16* cb = ctx + (size_t)(b % Bc) * L * CTX;
  RawVec8<float> rr[4];
  float rc[4][CTX];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int l = base + (half * 4 + i) * 128 + q * 32 + lane;
    if (l < L) {
      rr[i].load(xr + (size_t)l * 8);
#pragma unroll
      for (int ci = 0; ci < CTX; ++ci) rc[i][ci] = to_f32(cb[(size_t)l * CTX + ci]);
    }
  }
  mbar_wait_at(&sm.bar, 0, (5u << 16) | __LINE__, (uint32_t)tag);      // bounded, logged wait (ptx.cuh; file id 5 = d0.cuh)
  tc_fence_after();
  float fa[8], fq[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { fa[j] = 0.f; fq[j] = 0.f; }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int mb = half * 4 + i;
    uint32_t v[16];
    tmem_ld16(tmem_base + (uint32_t(q * 32) << 16) + mb * 16, v);
    tmem_ld_wait();
    const int l = base + mb * 128 + q * 32 + lane;
    if (l < L) {
      float r8[8], acc[8];
      rr[i].get(r8);
      float mean = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) { acc[j] = __uint_as_float(v[j]) + s_v[j] + r8[j]; mean += acc[j]; }     // conv2 + b2 + x
      mean *= 0.125f;
      float var = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = acc[j] - mean; var += d * d; }
      const float rstd = rsqrtf(var * 0.125f + eps);
      float m[8], y[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { m[j] = (acc[j] - mean) * rstd * s_v[8 + j] + s_v[16 + j]; y[j] = m[j] + s_v[24 + j]; }
      unsigned long long y2[4];          // the 10 -> 8 contraction as packed fp32 pairs (FFMA2: same rounding as two FFMAs)
#pragma unroll
      for (int j = 0; j < 4; ++j) y2[j] = f2pack(y[2 * j], y[2 * j + 1]);
#pragma unroll
      for (int ci = 0; ci < 8; ++ci) axpy8(y2, m[ci], &s_wi[ci * 8]);
#pragma unroll
      for (int ci = 0; ci < CTX; ++ci) axpy8(y2, rc[i][ci], &s_wi[(8 + ci) * 8]);
#pragma unroll
      for (int j = 0; j < 4; ++j) f2unpack(y2[j], y[2 * j], y[2 * j + 1]);
      const size_t g = ((size_t)b * L + l) * 8;
      Vec8<float>::store(out_r + g, y);
      if (out_t) Vec8<__nv_bfloat16>::store(out_t + g, y);
#pragma unroll
      for (int j = 0; j < 8; ++j) { fa[j] += y[j]; fq[j] = fmaf(y[j], y[j], fq[j]); }
    }
  }
  if (stats_out) {
    double v16[16];
#pragma unroll
    for (int j = 0; j < 8; ++j) { v16[j] = (double)fa[j]; v16[8 + j] = (double)fq[j]; }
    const double tot = warp_reduce16(v16, lane);
    if ((lane & 1) == 0) {
      const int idx = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
      atomicAdd(&sm.red[(idx & 7) * 2 + (idx >> 3)], tot);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (stats_out && tid < 16) atomicAdd(&stats_out[(size_t)b * 16 + tid], sm.red[tid]);
  if (warp == 0) tmem_dealloc(tmem_base, 128);
}

// h [B, L, 8] (T), x / out_r [B, L, 8] f32 (in place), ctx [Bc, L, CTX] (T), mod = scale[8] | shift[8] of clip b % mod_bmod,
// wi [8 + CTX][8] f32 (k, co), xbias [B, xb_stride] or null.
template <typename T, int CTX>
__global__ void __launch_bounds__(256, 3) d0_tail_kernel(const T* __restrict__ h, const double* __restrict__ stats_in,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         const float* __restrict__ w, const float* __restrict__ bias,
                                                         const float* x_r, const float* __restrict__ mod, int mod_bstride, int mod_bmod,
                                                         const T* __restrict__ ctx, int Bc, const float* __restrict__ wi,
                                                         const float* __restrict__ bi, const float* __restrict__ xbias, int xb_stride,
                                                         float* out_r, T* __restrict__ out_t, double* __restrict__ stats_out, int L,
                                                         float eps) {
  pdl_trigger();
  pdl_wait();
  __shared__ D0Smem sm;
  __shared__ float s_wi[(8 + CTX) * 8];
  __shared__ float s_v[4 * 8];     // conv bias | 1 + mod scale | mod shift | inject bias + cross-attention bias
  const int tid = threadIdx.x, b = blockIdx.y;
  const T* hb = h + (size_t)b * L * 8;
  const float* xr = x_r + (size_t)b * L * 8;
  const T* cb = ctx + (size_t)(b % Bc) * L * CTX;
  const int base = blockIdx.x * d0_positions_per_block();
  RawVec8<T> rh[kD0Per];
  RawVec8<float> rr[kD0Per];
  float rc[kD0Per][CTX];
#pragma unroll
  for (int k = 0; k < kD0Per; ++k) {
    const int l = base - 1 + tid + 256 * k;
    if (l >= 0 && l < L) {
      rh[k].load(hb + (size_t)l * 8);
      rr[k].load(xr + (size_t)l * 8);
#pragma unroll
      for (int ci = 0; ci < CTX; ++ci) rc[k][ci] = to_f32(cb[(size_t)l * CTX + ci]);
    }
  }
  d0_coef(sm, stats_in + (size_t)b * 16, gamma, beta, L, eps);
  if (tid < 192) sm.w[tid] = w[tid];
  for (int i = tid; i < (8 + CTX) * 8; i += 256) s_wi[i] = wi[i];
  if (tid < 8) {
    const float* md = mod + (size_t)(b % mod_bmod) * mod_bstride;
    s_v[tid] = bias[tid];
    s_v[8 + tid] = 1.f + md[tid];
    s_v[16 + tid] = md[8 + tid];
    s_v[24 + tid] = bi[tid] + (xbias ? xbias[(size_t)b * xb_stride + tid] : 0.f);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kD0Per; ++k) {
    const int r = tid + 256 * k, l = base - 1 + r;
    if (l >= 0 && l < L) { float xv[8]; rh[k].get(xv); d0_act_row(sm, xv, &sm.t[r * kD0Pitch]); }
    else d0_zero_row(&sm.t[r * kD0Pitch]);
  }
  __syncthreads();
  float acc[kD0Per][8];
#pragma unroll
  for (int k = 0; k < kD0Per; ++k) {
    float r8[8];
    rr[k].get(r8);                                     // residual x (garbage for out-of-range rows: never stored)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[k][j] = r8[j] + s_v[j];
  }
  d0_conv3x3(sm, tid, acc);                            // r = conv2(SiLU(GN2(h))) + b2 + x
  float fa[8], fq[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { fa[j] = 0.f; fq[j] = 0.f; }
#pragma unroll
  for (int k = 0; k < kD0Per; ++k) {
    const int r = tid + 256 * k, l = base - 1 + r;
    if (r >= 1 && r <= kD0Rows - 2 && l < L) {
      const size_t g = ((size_t)b * L + l) * 8;
      float mean = 0.f;                                // Modulation: LayerNorm over the 8 channels (two-pass), * (1 + s) + sh
#pragma unroll
      for (int j = 0; j < 8; ++j) mean += acc[k][j];
      mean *= 0.125f;
      float var = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = acc[k][j] - mean; var += d * d; }
      const float rstd = rsqrtf(var * 0.125f + eps);
      float m[8], y[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { m[j] = (acc[k][j] - mean) * rstd * s_v[8 + j] + s_v[16 + j]; y[j] = m[j] + s_v[24 + j]; }
#pragma unroll
      for (int ci = 0; ci < 8; ++ci)
#pragma unroll
        for (int j = 0; j < 8; ++j) y[j] += m[ci] * s_wi[ci * 8 + j];
#pragma unroll
      for (int ci = 0; ci < CTX; ++ci)
#pragma unroll
        for (int j = 0; j < 8; ++j) y[j] += rc[k][ci] * s_wi[(8 + ci) * 8 + j];
      Vec8<float>::store(out_r + g, y);
      if (out_t) store_operand8<T>(out_t + g, y);
#pragma unroll
      for (int j = 0; j < 8; ++j) { fa[j] += y[j]; fq[j] = fmaf(y[j], y[j], fq[j]); }
    }
  }
  if (stats_out) d0_flush_stats(sm, fa, fq, stats_out + (size_t)b * 16);
}


// Depth-0 output projection of the general cross-attention item (a10, M_ctx > 1; C = 8 is below every tensor-core tile):
//   out[b, l, c] = x[b, l, c] + sum_k o[b, l, k] Wo[c, k]      o [B, L, 512] operand precision, Wo [8][512] f32
// + operand copy and GroupNorm sums of the output (group = channel at C = 8).  One position per thread, Wo in smem.
template <typename T>
__global__ void __launch_bounds__(256) xattn_out_c8_kernel(const T* __restrict__ o, const float* __restrict__ wo, const float* resid,
                                                           float* out_r, T* __restrict__ out_t, double* __restrict__ stats, int L) {
  pdl_trigger();
  __shared__ float sw[8 * 512];
  __shared__ Stats8Smem sm;
  for (int i = threadIdx.x; i < 8 * 512; i += 256) sw[i] = wo[i];
  pdl_wait();
  __syncthreads();
  const int b = blockIdx.y;
  const int l = blockIdx.x * 256 + threadIdx.x;
  const bool valid = l < L;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (valid) {
    const T* row = o + ((size_t)b * L + l) * 512;
    for (int k = 0; k < 512; k += 8) {
      float x[8];
      if constexpr (sizeof(T) == 2) {
        const uint4 u = *reinterpret_cast<const uint4*>(row + k);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) { x[2 * j] = __uint_as_float(w[j] << 16); x[2 * j + 1] = __uint_as_float(w[j] & 0xFFFF0000u); }
      } else {
        const float4 a = *reinterpret_cast<const float4*>(row + k), c = *reinterpret_cast<const float4*>(row + k + 4);
        x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = c.x; x[5] = c.y; x[6] = c.z; x[7] = c.w;
      }
#pragma unroll
      for (int c = 0; c < 8; ++c)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[c] = fmaf(x[j], sw[c * 512 + k + j], acc[c]);
    }
    const size_t g = ((size_t)b * L + l) * 8;
    const float4 r0 = *reinterpret_cast<const float4*>(resid + g), r1 = *reinterpret_cast<const float4*>(resid + g + 4);
    acc[0] += r0.x; acc[1] += r0.y; acc[2] += r0.z; acc[3] += r0.w; acc[4] += r1.x; acc[5] += r1.y; acc[6] += r1.z; acc[7] += r1.w;
    *reinterpret_cast<float4*>(out_r + g) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    *reinterpret_cast<float4*>(out_r + g + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
    if (out_t != nullptr) {
#pragma unroll
      for (int c = 0; c < 8; ++c) out_t[g + c] = from_f32<T>(acc[c]);
    }
  }
  if (stats != nullptr) stats8_block_reduce(acc, valid, stats + (size_t)b * 16, sm);
}

}  // namespace sfb
