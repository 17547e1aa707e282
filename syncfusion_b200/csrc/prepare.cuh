// K9 + the M_ctx = 1 cross-attention fast path: everything that is loop-invariant, run ONCE per sample() call
// (SURVEY.md 0.3, 0.4, D.3).  All fp32, tiny; one warp per output feature.
//   fourier_embed : sigma -> [sigma, sin(2 pi sigma w), cos(2 pi sigma w)]       (TimeConditioningPlugin, a3)
//   linear_act    : out[r, j] = act_out(dot(act_in(in[r, :]), W[j, :]) + bias[j])  (time MLP, Modulation / SkipModulate
//                   Linear(SiLU(features)) tables, cross-attention V and out projections)
//   ln_rows       : affine LayerNorm of the embedding rows (Attention.norm_context, a10)
#pragma once
#include "ptx.cuh"

namespace sfb {

__global__ void fourier_embed_kernel(const float* __restrict__ sigma, const float* __restrict__ w, float* __restrict__ out,
                                     int rows, int half) {
  const int r = blockIdx.x;
  const int i = threadIdx.x;
  if (r >= rows) return;
  const float t = sigma[r];
  float* o = out + (size_t)r * (2 * half + 1);
  if (i == 0) o[0] = t;
  if (i < half) {
    const float f = t * w[i] * 6.283185307179586f;
    o[1 + i] = sinf(f);
    o[1 + half + i] = cosf(f);
  }
}

// LinearSchedule: torch.linspace(1, 0, steps) in fp32 (ATen's two-sided formula)
__global__ void sigma_linspace_kernel(float* __restrict__ out, int steps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= steps) return;
  const float start = 1.f, end = 0.f;
  const float step = (end - start) / (float)(steps - 1);
  out[i] = i < steps / 2 ? start + step * (float)i : end - step * (float)(steps - 1 - i);
}

enum { ACT_NONE = 0, ACT_GELU = 1, ACT_SILU = 2 };
__device__ __forceinline__ float act_apply(float v, int act) {
  if (act == ACT_GELU) return 0.5f * v * (1.f + erff(v * 0.7071067811865476f));
  if (act == ACT_SILU) return v / (1.f + expf(-v));
  return v;
}

// in [rows, K], W [J, K] (row pitch ldw), out [rows, ldo] at column offset; one warp per output feature j.
__global__ void __launch_bounds__(256) linear_act_kernel(const float* __restrict__ in, const float* __restrict__ W,
                                                         const float* __restrict__ bias, float* __restrict__ out, int rows,
                                                         int K, int J, int ldw, int ldo, int act_in, int act_out) {
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (j >= J) return;
  const float* wr = W + (size_t)j * ldw;
  const float bj = bias ? bias[j] : 0.f;
  for (int r = 0; r < rows; ++r) {
    const float* ir = in + (size_t)r * K;
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) acc += act_apply(ir[k], act_in) * wr[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[(size_t)r * ldo + j] = act_apply(acc + bj, act_out);
  }
}

// in [rows, C] -> out [rows, C] = LN(in) * gamma + beta ; one warp per row
__global__ void __launch_bounds__(256) ln_rows_kernel(const float* __restrict__ in, const float* __restrict__ gamma,
                                                      const float* __restrict__ beta, float* __restrict__ out, int rows, int C,
                                                      float eps) {
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float* ir = in + (size_t)r * C;
  float s = 0.f;
  for (int k = lane; k < C; k += 32) s += ir[k];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / C;
  float q = 0.f;
  for (int k = lane; k < C; k += 32) { const float d = ir[k] - mean; q += d * d; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / C + eps);
  for (int k = lane; k < C; k += 32) out[(size_t)r * C + k] = (ir[k] - mean) * rstd * gamma[k] + beta[k];
}

// Effective context rows of the CFG-doubled batch: rows [0, n_cond) = the caller's embedding tokens [B, M, F] as they
// are, rows [n_cond, ...) = FixedEmbedding token (r - n_cond) % M (the mask branch of ClassifierFreeGuidancePlugin, A.4).
__global__ void build_emb_rows_kernel(const float* __restrict__ emb, const float* __restrict__ fixed, float* __restrict__ out,
                                      int n_cond, int M, int F) {
  const int r = blockIdx.x;
  const float* src = r < n_cond ? emb + (size_t)r * F : fixed + (size_t)((r - n_cond) % M) * F;
  for (int k = threadIdx.x; k < F; k += blockDim.x) out[(size_t)r * F + k] = src[k];
}
template <typename T>
__global__ void __launch_bounds__(256) f32_to_operand_kernel(const float* __restrict__ in, T* __restrict__ out, size_t n) {
  const size_t i = blockIdx.x * (size_t)256 + threadIdx.x;
  if (i < n) out[i] = from_f32<T>(in[i]);
}


// ---------------------------------------------------------------------------------------------------- LayerNorm fold
// InjectChannels on a Modulation output (a7 + a8), with the per-position LayerNorm folded OUT of the GEMM operand
// (sk_tc.cuh, SkParams::ln_fold):   W [ (LN(x) (1 + s) + sh) | ctx ] = rstd (W' x - mean ws) + Wc ctx + wsh
//   W'[n, k] = bf16(W[n, k] (1 + s[k]))   ws[n] = sum_k W'[n, k]   wsh[n] = sum_k W[n, k] sh[k]        (k < C)
// (1 + s, sh) change every sampler step (and per clip in unet_forward), so the scaled copy of every streaming-K inject
// weight is rebuilt by ONE launch per U-Net evaluation: one warp per (copy, output row).  ~27 MB per copy.
struct FoldItem {
  const __nv_bfloat16* W;   // [N][K] (K = C + ctx), K-major
  __nv_bfloat16* Wd;        // [copies][N][K]
  float* vec;               // [copies][2 N]: ws | wsh
  int C, K, N, mod_off, row0;
};
__global__ void __launch_bounds__(256) inject_fold_kernel(const FoldItem* __restrict__ items, int n_items, int total_rows,
                                                          const float* __restrict__ frow, int bstride, int copies) {
  pdl_trigger();
  pdl_wait();
  const int gw = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (gw >= total_rows * copies) return;
  const int copy = gw / total_rows, row = gw % total_rows;
  int it = 0;
  while (it + 1 < n_items && items[it + 1].row0 <= row) ++it;
  const FoldItem f = items[it];
  const int n = row - f.row0;
  const float* md = frow + (size_t)copy * bstride + f.mod_off;     // scale [C] | shift [C]
  const __nv_bfloat16* src = f.W + (size_t)n * f.K;
  __nv_bfloat16* dst = f.Wd + ((size_t)copy * f.N + n) * f.K;
  float ws = 0.f, wsh = 0.f;
  for (int k = lane * 8; k < f.K; k += 256) {
    const uint4 u = *reinterpret_cast<const uint4*>(src + k);
    if (k < f.C) {
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float w0 = __uint_as_float(w[j] << 16), w1 = __uint_as_float(w[j] & 0xFFFF0000u);
        const float2 sc = *reinterpret_cast<const float2*>(md + k + 2 * j);
        const float2 sh = *reinterpret_cast<const float2*>(md + f.C + k + 2 * j);
        const __nv_bfloat162 h = __floats2bfloat162_rn(w0 * (1.f + sc.x), w1 * (1.f + sc.y));
        ws += __bfloat162float(h.x) + __bfloat162float(h.y);
        wsh = fmaf(w0, sh.x, fmaf(w1, sh.y, wsh));
        o[j] = *reinterpret_cast<const uint32_t*>(&h);
      }
      *reinterpret_cast<uint4*>(dst + k) = make_uint4(o[0], o[1], o[2], o[3]);
    } else {
      *reinterpret_cast<uint4*>(dst + k) = u;      // context columns are not modulated
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { ws += __shfl_xor_sync(0xffffffffu, ws, o); wsh += __shfl_xor_sync(0xffffffffu, wsh, o); }
  if (lane == 0) {
    f.vec[(size_t)copy * 2 * f.N + n] = ws;
    f.vec[(size_t)copy * 2 * f.N + f.N + n] = wsh;
  }
}

}  // namespace sfb
