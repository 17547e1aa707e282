// K5b: fused self-attention core of AttentionItem (upstream a_unet Attention):
//   out[b, i, h, :] = softmax_j(q[b,i,h,:] . k[b,j,h,:] / sqrt(64)) v[b,j,h,:]     8 heads x 64, non-causal, no mask.
// Input is the fused QKV projection output [B, N, 1536] (q | k | v, heads contiguous 64-wide), output [B, N, 512].
//
// One CTA = one 128-query tile of one (clip, head).  Warp 0 lane 0: TMA producer (Q once, K/V tiles double
// buffered).  Warp 1 lane 0: tcgen05.mma issuer: S = Q K^T (K-major x K-major) into TMEM, then O_j = P V with V
// consumed straight from its TMA tile as an MN-major B operand.  Warps 2-5 (128 threads, one query row each):
// online softmax in fp32 - S is read from TMEM twice (row max, then exp2), P is written to 128B-swizzled smem in
// operand precision, the per-tile P V product is read back from TMEM and folded into a register accumulator with
// the running rescale.  The [B, 8, N, N] score matrix the reference materialises never exists.
#pragma once
#include "ptx.cuh"

namespace sfb {

template <typename T>
struct AttnParams {
  CUtensorMap tmQ;    // qkv viewed [1536, N, B], box [atom, 128, 1]
  CUtensorMap tmKV;   // same tensor, box [atom, BKV, 1]
  CUtensorMap tmV;    // same box; bf16: identical to tmKV, tf32: 128B swizzle with 32-byte atoms (MN-major tf32)
  T* out;             // [B, N, 512]
  int n_tokens;
  float scale_log2;   // log2(e) / sqrt(64)
};

template <typename T> struct AttnCfg;
template <> struct AttnCfg<__nv_bfloat16> { static constexpr int BKV = 128; };
template <> struct AttnCfg<float> { static constexpr int BKV = 64; };

__device__ __forceinline__ float fast_exp2(float x) {   // single MUFU.EX2
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

constexpr int kAttnThreads = 192;
constexpr int kHeadDim = 64;

template <typename T>
__host__ __device__ constexpr int attn_smem_bytes() {
  constexpr int DA = kHeadDim / ElemTraits<T>::kAtomElems;
  constexpr int BKV = AttnCfg<T>::BKV;
  return DA * 128 * 128 /*Q*/ + 2 * 2 * DA * BKV * 128 /*K,V x 2 stages*/ + 2 * 128 * 128 /*P*/ + 256;
}

template <typename T>
__global__ void __launch_bounds__(kAttnThreads, 2) attn_tc_kernel(const __grid_constant__ AttnParams<T> p) {
  pdl_trigger();
  using TR = ElemTraits<T>;
  constexpr int AE = TR::kAtomElems;            // elements per 128-byte row
  constexpr int DA = kHeadDim / AE;             // atoms along head dim (1 bf16, 2 tf32)
  constexpr int BKV = AttnCfg<T>::BKV;
  constexpr int PA = BKV / AE;                  // atoms of P along the key dim (2)
  constexpr int UK = TR::kUmmaK;
  constexpr int kQBytes = DA * 128 * 128;
  constexpr int kKBytes = DA * BKV * 128;
  constexpr int kPBytes = PA * 128 * 128;
  constexpr uint32_t kTmemCols = 256;           // S: [0, BKV), O_j: [BKV, BKV + 64)

  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + kQBytes;                  // stage s: K at sKV + s*2*kKBytes, V right after K
  uint8_t* sP = sKV + 4 * kKBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + kPBytes);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;    // [2]
  uint64_t* kv_empty = bars + 3;   // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_ready = bars + 6;
  uint64_t* o_full = bars + 7;
  uint64_t* o_free = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int nkv = (p.n_tokens + BKV - 1) / BKV;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmKV);
    tma_prefetch_desc(&p.tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    mbar_init(s_full, 1);
    mbar_init(p_ready, 128);
    mbar_init(o_full, 1);
    mbar_init(o_free, 128);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, kTmemCols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  pdl_wait();            // everything above is independent of the previous kernel's output
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + BKV;

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------- TMA producer
    mbar_expect_tx(q_full, kQBytes);
    for (int a = 0; a < DA; ++a) tma_load_3d(sQ + a * 128 * 128, &p.tmQ, q_full, h * kHeadDim + a * AE, q0, b);
    for (int j = 0; j < nkv; ++j) {
      const int s = j & 1;
      mbar_wait(&kv_empty[s], ((j >> 1) & 1) ^ 1);
      mbar_expect_tx(&kv_full[s], 2 * kKBytes);
      uint8_t* sk = sKV + s * 2 * kKBytes;
      for (int a = 0; a < DA; ++a) {
        tma_load_3d(sk + a * BKV * 128, &p.tmKV, &kv_full[s], 512 + h * kHeadDim + a * AE, j * BKV, b);
        tma_load_3d(sk + kKBytes + a * BKV * 128, &p.tmV, &kv_full[s], 1024 + h * kHeadDim + a * AE, j * BKV, b);
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ------------------------------------------------------------- MMA issuer
    constexpr uint32_t idesc_s = make_idesc(TR::kFmt, 128, BKV, 0, 0);        // Q (K-major) x K (K-major)
    constexpr uint32_t idesc_o = make_idesc(TR::kFmt, 128, kHeadDim, 0, 1);   // P (K-major) x V (MN-major)
    auto issue_s = [&](int s) {
      const uint32_t aq = smem_u32(sQ), ak = smem_u32(sKV + s * 2 * kKBytes);
#pragma unroll
      for (int k = 0; k < kHeadDim / UK; ++k) {
        const int atom = (k * UK) / AE, within = k % (AE / UK);
        const uint64_t da = make_smem_desc_sw128(aq + atom * 128 * 128 + within * 32, 16, 1024);
        const uint64_t db = make_smem_desc_sw128(ak + atom * BKV * 128 + within * 32, 16, 1024);
        umma_ss<TR::kTF32>(tmem_S, da, db, idesc_s, k != 0);
      }
    };
    auto issue_o = [&](int s) {
      const uint32_t ap = smem_u32(sP), av = smem_u32(sKV + s * 2 * kKBytes + kKBytes);
#pragma unroll
      for (int k = 0; k < BKV / UK; ++k) {
        const int atom = (k * UK) / AE, within = k % (AE / UK);
        const uint64_t da = make_smem_desc_sw128(ap + atom * 128 * 128 + within * 32, 16, 1024);
        // MN-major B: rows of the tile are keys (the MMA K dim) at 128-byte pitch; 8-key groups 1024 bytes apart
        // (SBO); head-dim atoms BKV*128 bytes apart (LBO, only used by the two-atom tf32 layout).
        // bf16: SWIZZLE_128B, 8-key groups; tf32: SWIZZLE_128B_BASE32B, 4-key groups 512 bytes apart.
        const uint64_t db = (sizeof(T) == 2) ? make_smem_desc(av + k * UK * 128, BKV * 128, 1024, 2)
                                             : make_smem_desc(av + k * UK * 128, BKV * 128, 512, 1);
        umma_ss<TR::kTF32>(tmem_O, da, db, idesc_o, k != 0);
      }
    };
    mbar_wait(q_full, 0);
    mbar_wait(&kv_full[0], 0);
    tc_fence_after();
    issue_s(0);
    umma_commit(s_full);
    for (int j = 0; j < nkv; ++j) {
      const int s = j & 1;
      mbar_wait(p_ready, j & 1);
      if (j > 0) mbar_wait(o_free, (j - 1) & 1);
      tc_fence_after();
      issue_o(s);
      umma_commit(o_full);
      umma_commit(&kv_empty[s]);
      if (j + 1 < nkv) {
        const int s2 = (j + 1) & 1;
        mbar_wait(&kv_full[s2], ((j + 1) >> 1) & 1);
        tc_fence_after();
        issue_s(s2);
        umma_commit(s_full);
      }
    }
  } else if (warp >= 2) {
    // ------------------------------------------------------------- softmax + output (one query row per thread)
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    float o_acc[kHeadDim];
#pragma unroll
    for (int i = 0; i < kHeadDim; ++i) o_acc[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    const float sc = p.scale_log2;
    for (int j = 0; j < nkv; ++j) {
      const int kv_valid = min(BKV, p.n_tokens - j * BKV);
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      float mx = m_run;
      const bool full_tile = kv_valid == BKV;   // tile-uniform: only the last tile of a ragged sequence is masked
#pragma unroll 1
      for (int c = 0; c < BKV; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_S + lane_off + c, v);
        tmem_ld_wait();
        if (full_tile) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c + i < kv_valid) mx = fmaxf(mx, __uint_as_float(v[i]));
        }
      }
      const float alpha = (m_run == -INFINITY) ? 0.f : exp2f((m_run - mx) * sc);
      const float moff = mx * sc;
      float rs0 = 0.f, rs1 = 0.f;
#pragma unroll 1
      for (int c = 0; c < BKV; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_S + lane_off + c, v);
        tmem_ld_wait();
        float pv[32];
        if (full_tile) {
#pragma unroll
          for (int i = 0; i < 32; ++i) pv[i] = fast_exp2(fmaf(__uint_as_float(v[i]), sc, -moff));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) pv[i] = (c + i < kv_valid) ? fast_exp2(fmaf(__uint_as_float(v[i]), sc, -moff)) : 0.f;
        }
        // row sum in fp32 on the un-rounded probabilities (two independent chains); the operand-precision rounding of
        // P is zero-mean, so the normaliser differs from sum(round(P)) by ~1e-4 relative at most.
#pragma unroll
        for (int i = 0; i < 32; i += 2) { rs0 += pv[i]; rs1 += pv[i + 1]; }
        // write this row's 32 probabilities into the swizzled K-major P tile (operand precision)
        if constexpr (sizeof(T) == 2) {
          const int atom = c / AE;               // 64 keys per atom
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {       // 4 chunks of 8 keys (16 bytes)
            const int cc = ((c % AE) / 8) + ch;  // chunk index inside the 128-byte row
            uint32_t w[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              __nv_bfloat162 h2 = __floats2bfloat162_rn(pv[ch * 8 + 2 * t], pv[ch * 8 + 2 * t + 1]);
              w[t] = *reinterpret_cast<uint32_t*>(&h2);
            }
            uint4* dst = reinterpret_cast<uint4*>(sP + atom * 128 * 128 + r * 128 + ((cc ^ (r & 7)) * 16));
            *dst = make_uint4(w[0], w[1], w[2], w[3]);
          }
        } else {
          const int atom = c / AE;               // 32 keys per atom: this 32-column pass fills one atom row
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) {       // 8 chunks of 4 keys
            uint4* dst = reinterpret_cast<uint4*>(sP + atom * 128 * 128 + r * 128 + ((ch ^ (r & 7)) * 16));
            *dst = make_uint4(__float_as_uint(from_f32<float>(pv[ch * 4])), __float_as_uint(from_f32<float>(pv[ch * 4 + 1])),
                              __float_as_uint(from_f32<float>(pv[ch * 4 + 2])), __float_as_uint(from_f32<float>(pv[ch * 4 + 3])));
          }
        }
      }
      const float rs = rs0 + rs1;
      tc_fence_before();
      fence_proxy_async();
      mbar_arrive(p_ready);
      l_run = l_run * alpha + rs;
      m_run = mx;
      mbar_wait(o_full, j & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < kHeadDim; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_O + lane_off + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o_acc[c + i] = o_acc[c + i] * alpha + __uint_as_float(v[i]);
      }
      tc_fence_before();
      mbar_arrive(o_free);
    }
    if (q0 + r < p.n_tokens) {
      const float inv = 1.f / l_run;
      T* dst = p.out + ((size_t)b * p.n_tokens + q0 + r) * 512 + h * kHeadDim;
      if constexpr (sizeof(T) == 2) {
#pragma unroll
        for (int c = 0; c < kHeadDim; c += 8) {
          uint32_t w[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            __nv_bfloat162 h2 = __floats2bfloat162_rn(o_acc[c + 2 * t] * inv, o_acc[c + 2 * t + 1] * inv);
            w[t] = *reinterpret_cast<uint32_t*>(&h2);
          }
          *reinterpret_cast<uint4*>(dst + c) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      } else {
#pragma unroll
        for (int c = 0; c < kHeadDim; c += 4)
          *reinterpret_cast<float4*>(dst + c) =
              make_float4(o_acc[c] * inv, o_acc[c + 1] * inv, o_acc[c + 2] * inv, o_acc[c + 3] * inv);
      }
    }
  }
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace sfb
