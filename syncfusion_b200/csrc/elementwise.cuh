// Bandwidth-bound passes of the U-Net: all channels-last, 128-bit accesses.
//   gn_apply_silu : GroupNorm(8 groups) apply + SiLU from statistics accumulated by the producer's epilogue (a6)
//   ln_mod        : per-position LayerNorm over C (no affine) * (1 + scale) + shift (ModulationItem, a7), or plain
//                   normalisation for the attention pre-norm whose affine is folded into W_qkv (a9)
//   sampler_update: CFG combine + v-sampler update (a2, a4)
//   ncl_to_nlc    : onset-pyramid layout change, once per sample() (D.2)
#pragma once
#include "ptx.cuh"

namespace sfb {

template <typename Tin> struct Vec8;
template <> struct Vec8<float> {
  static __device__ __forceinline__ void load(const float* p, float* v) {
    float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float* v) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
};
template <> struct Vec8<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float* v) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float* v) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h2);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

// operand-precision store: bf16 rounds in the pack; fp32-mode operands are pre-rounded to tf32 (see from_f32<float>)
template <typename T>
__device__ __forceinline__ void store_operand8(T* p, const float* v) {
  if constexpr (sizeof(T) == 4) {
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = from_f32<float>(v[j]);
    Vec8<float>::store(p, r);
  } else {
    Vec8<T>::store(p, v);
  }
}

__device__ __forceinline__ float silu_f(float y) { return y / (1.f + __expf(-y)); }

// ---------------------------------------------------------------------------------------------- GroupNorm + SiLU
// in [B, L, C] (Tin), stats [B, 8, 2] f64 (sum, sum of squares over L * gs elements), out [B, L, C] (Tout).
// grid = (blocks_per_clip, B); dynamic smem = 2 * C floats (per-channel a, b with y = x * a + b).
template <typename Tin, typename Tout>
__global__ void __launch_bounds__(256) gn_apply_silu_kernel(const Tin* __restrict__ in, const double* __restrict__ stats,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            Tout* __restrict__ out, int L, int C, int gs, float eps) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float s_ab[];
  __shared__ float s_g[16];   // per group: mean, rstd
  const int b = blockIdx.y;
  if (threadIdx.x < 8) {
    const double cnt = (double)L * gs;
    const double s1 = stats[(size_t)b * 16 + threadIdx.x * 2], s2 = stats[(size_t)b * 16 + threadIdx.x * 2 + 1];
    const double mean = s1 / cnt;
    double var = s2 / cnt - mean * mean;
    var = var > 0 ? var : 0;
    s_g[threadIdx.x * 2] = (float)mean;
    s_g[threadIdx.x * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / gs;
    const float a = s_g[g * 2 + 1] * gamma[c];
    s_ab[c] = a;
    s_ab[C + c] = beta[c] - s_g[g * 2] * a;
  }
  __syncthreads();
  constexpr int U = 4;   // independent 128-bit loads in flight per thread
  const size_t nvec = (size_t)L * C / 8;
  const Tin* src = in + (size_t)b * L * C;
  Tout* dst = out + (size_t)b * L * C;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < nvec; i0 += stride * U) {
    float v[U][8];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t i = i0 + u * stride;
      if (i < nvec) Vec8<Tin>::load(src + i * 8, v[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t i = i0 + u * stride;
      if (i < nvec) {
        const int c0 = (int)((i * 8) % C);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[u][j] = silu_f(v[u][j] * s_ab[c0 + j] + s_ab[C + c0 + j]);
        store_operand8<Tout>(dst + i * 8, v[u]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- LayerNorm (+ modulation)
// in [rows, C] f32; LPR = min(32, C/8) lanes cooperate on a row, each holding NV = C / (8 * LPR) vectors of 8.
// scale/shift: [C] each at (b % bmod) * bstride (null -> plain normalisation).  out_t operand copy, out_r fp32 copy
// (nullable; may alias `in`).
template <typename Tout, int NV>
__global__ void __launch_bounds__(256) ln_mod_kernel(const float* in, const float* __restrict__ scale,
                                                     const float* __restrict__ shift, int bstride, int bmod,
                                                     Tout* __restrict__ out_t, float* out_r, size_t rows, int rows_per_clip,
                                                     int C, float eps) {
  pdl_trigger();
  pdl_wait();
  const int lpr = (C / 8 < 32) ? C / 8 : 32;
  const int rpw = 32 / lpr;
  const int lane = threadIdx.x & 31;
  const size_t warp_global = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const size_t row = warp_global * rpw + lane / lpr;
  const int sub = lane % lpr;
  const bool active = row < rows;
  float v[NV][8];
  float sum = 0.f;
  if (active) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      Vec8<float>::load(in + row * C + (size_t)(k * lpr + sub) * 8, v[k]);
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += v[k][j];
    }
  }
  for (int o = lpr >> 1; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / (float)C;
  float sq = 0.f;
  if (active) {
#pragma unroll
    for (int k = 0; k < NV; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = v[k][j] - mean; sq += d * d; }
  }
  for (int o = lpr >> 1; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  if (!active) return;
  const float rstd = rsqrtf(sq / (float)C + eps);
  const int b = (int)(row / rows_per_clip);
  const float* sc = scale ? scale + (size_t)(b % bmod) * bstride : nullptr;
  const float* sh = shift ? shift + (size_t)(b % bmod) * bstride : nullptr;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c0 = (k * lpr + sub) * 8;
    float y[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float t = (v[k][j] - mean) * rstd;
      if (sc) t = t * (1.f + __ldg(&sc[c0 + j])) + __ldg(&sh[c0 + j]);
      y[j] = t;
    }
    if (out_t) store_operand8<Tout>(out_t + row * C + c0, y);
    if (out_r) Vec8<float>::store(out_r + row * C + c0, y);
  }
}

// ---------------------------------------------------------------------------------------------- sampler update (K10)
// v [Beff, L] (cond rows [0,B), uncond rows [B,2B) when cfg), x [B, L] -> x_next, three-line form of VSampler:
//   v = v_u + (v_c - v_u) * scale ; x_pred = a x - b v ; n_pred = b x + a v ; x' = a' x_pred + b' n_pred
__global__ void __launch_bounds__(256) sampler_update_kernel(const float* __restrict__ x_eval, const float* __restrict__ v,
                                                             float* __restrict__ x_out, float* __restrict__ traj_x,
                                                             float* __restrict__ traj_v, size_t n, int cfg, float scale,
                                                             float a, float b, float a2, float b2) {
  pdl_trigger();
  pdl_wait();
  const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  const float4 xv = *reinterpret_cast<const float4*>(x_eval + i);
  float4 vc = *reinterpret_cast<const float4*>(v + i);
  if (cfg) {
    const float4 vu = *reinterpret_cast<const float4*>(v + n + i);
    vc.x = vu.x + (vc.x - vu.x) * scale;
    vc.y = vu.y + (vc.y - vu.y) * scale;
    vc.z = vu.z + (vc.z - vu.z) * scale;
    vc.w = vu.w + (vc.w - vu.w) * scale;
  }
  float4 o;
  {
    const float xp = a * xv.x - b * vc.x, np = b * xv.x + a * vc.x; o.x = a2 * xp + b2 * np;
  }
  {
    const float xp = a * xv.y - b * vc.y, np = b * xv.y + a * vc.y; o.y = a2 * xp + b2 * np;
  }
  {
    const float xp = a * xv.z - b * vc.z, np = b * xv.z + a * vc.z; o.z = a2 * xp + b2 * np;
  }
  {
    const float xp = a * xv.w - b * vc.w, np = b * xv.w + a * vc.w; o.w = a2 * xp + b2 * np;
  }
  *reinterpret_cast<float4*>(x_out + i) = o;
  if (traj_x) *reinterpret_cast<float4*>(traj_x + i) = o;
  if (traj_v) *reinterpret_cast<float4*>(traj_v + i) = vc;
}

// CFG combine only (free-standing net call): v_out = v_u + (v_c - v_u) * scale
__global__ void __launch_bounds__(256) cfg_combine_kernel(const float* __restrict__ v, float* __restrict__ v_out, size_t n,
                                                          int cfg, float scale) {
  const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  float4 vc = *reinterpret_cast<const float4*>(v + i);
  if (cfg) {
    const float4 vu = *reinterpret_cast<const float4*>(v + n + i);
    vc.x = vu.x + (vc.x - vu.x) * scale;
    vc.y = vu.y + (vc.y - vu.y) * scale;
    vc.z = vu.z + (vc.z - vu.z) * scale;
    vc.w = vu.w + (vc.w - vu.w) * scale;
  }
  *reinterpret_cast<float4*>(v_out + i) = vc;
}

// ---------------------------------------------------------------------------------------------- layout change
// in [B, C, L] f32 (NCL, as main/generation.py:80 passes `channels`) -> out [B, L, C] (operand precision)
template <typename Tout>
__global__ void __launch_bounds__(256) ncl_to_nlc_kernel(const float* __restrict__ in, Tout* __restrict__ out, int C, int L) {
  const int b = blockIdx.y;
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= L) return;
  const float* src = in + (size_t)b * C * L + l;
  Tout* dst = out + ((size_t)b * L + l) * C;
  for (int c = 0; c < C; ++c) dst[c] = from_f32<Tout>(src[(size_t)c * L]);
}

}  // namespace sfb
