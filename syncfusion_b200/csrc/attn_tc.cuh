// K5b: fused self-attention core of AttentionItem (upstream a_unet Attention):
//   out[b, i, h, :] = softmax_j(q[b,i,h,:] . k[b,j,h,:] / sqrt(64)) v[b,j,h,:]     8 heads x 64, non-causal, no mask.
// Input is the fused QKV projection output [B, N, 1536] (q | k | v, heads contiguous 64-wide), output [B, N, 512].
//
// One CTA = one 128-query tile of one (clip, head), two CTAs per SM.  Warp 0 lane 0: TMA producer (Q once, K/V tiles
// double buffered).  Warp 1 lane 0: tcgen05.mma issuer: S = Q K^T (K-major x K-major) into TMEM, O += P V ACCUMULATED
// IN TMEM over all key tiles (V consumed straight from its TMA tile as an MN-major B operand).  Warps 2-5 (128
// threads, one query row each): S is read from TMEM ONCE into registers (which frees the S columns at once: the MMA
// warp issues S(j+1) while softmax(j) is still computing), row max, exp2, P written to 128B-swizzled smem in operand
// precision.  The running max is LAZY: the O accumulator and the row sum are rescaled only when a row's max grows by
// more than 2^8 (then the warp reads O from TMEM, scales, writes it back); otherwise probabilities simply stay
// relative to the older max (<= 2^8, exact in fp32 / harmless in bf16) - no per-tile O read-back, no per-tile
// multiply of the accumulator.  The [B, 8, N, N] score matrix the reference materialises never exists.
#pragma once
#include "ptx.cuh"
#undef SFB_FILE_ID
#define SFB_FILE_ID 2   // attn_tc.cuh

namespace sfb {

template <typename T>
struct AttnParams {
  CUtensorMap tmQ;    // qkv viewed [1536, N, B], box [atom, 128, 1]
  CUtensorMap tmKV;   // same tensor, box [atom, BKV, 1]
  CUtensorMap tmV;    // same box; bf16: identical to tmKV, tf32: 128B swizzle with 32-byte atoms (MN-major tf32)
  T* out;             // [B, N, 512]
  int n_tokens;       // query tokens N
  int kv_tokens;      // key / value tokens (= N for self-attention; M_ctx for cross-attention, a9 / a10)
  int q_col0, k_col0, v_col0;   // first column of q / k / v in their tensors (self: 0 / 512 / 1024 of one [.., 1536] tensor;
                                // cross: q in a [.., 512] tensor, k | v in a [.., 1024] tensor)
  float scale_log2;   // log2(e) / sqrt(64)
  int tag;            // plan op index (wait log)
};

template <typename T> struct AttnCfg;
template <> struct AttnCfg<__nv_bfloat16> { static constexpr int BKV = 128; };
template <> struct AttnCfg<float> { static constexpr int BKV = 64; };

__device__ __forceinline__ float fast_exp2(float x) {   // single MUFU.EX2
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

constexpr int kAttnThreads = 320;      // TMA warp, MMA warp, 8 softmax warps (two threads per query row)
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
  constexpr uint32_t kTmemCols = 256;           // S: [0, BKV), O: [BKV, BKV + 64), then 6 exchange columns

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
  uint64_t* s_free = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int nkv = (p.kv_tokens + BKV - 1) / BKV;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmKV);
    tma_prefetch_desc(&p.tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    mbar_init(s_full, 1);
    mbar_init(p_ready, 256);
    mbar_init(o_full, 1);
    mbar_init(s_free, 256);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, kTmemCols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  pdl_wait();            // everything above is independent of the previous kernel's output
  mark_progress(p.tag);
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + BKV;

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------- TMA producer
    mbar_expect_tx(q_full, kQBytes);
    for (int a = 0; a < DA; ++a) tma_load_3d(sQ + a * 128 * 128, &p.tmQ, q_full, p.q_col0 + h * kHeadDim + a * AE, q0, b);
    for (int j = 0; j < nkv; ++j) {
      const int s = j & 1;
      mbar_wait(&kv_empty[s], ((j >> 1) & 1) ^ 1);
      mbar_expect_tx(&kv_full[s], 2 * kKBytes);
      uint8_t* sk = sKV + s * 2 * kKBytes;
      for (int a = 0; a < DA; ++a) {
        tma_load_3d(sk + a * BKV * 128, &p.tmKV, &kv_full[s], p.k_col0 + h * kHeadDim + a * AE, j * BKV, b);
        tma_load_3d(sk + kKBytes + a * BKV * 128, &p.tmV, &kv_full[s], p.v_col0 + h * kHeadDim + a * AE, j * BKV, b);
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
    auto issue_o = [&](int s, bool accumulate) {
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
        umma_ss<TR::kTF32>(tmem_O, da, db, idesc_o, (accumulate || k != 0) ? 1u : 0u);
      }
    };
    mbar_wait(q_full, 0);
    mbar_wait(&kv_full[0], 0);
    tc_fence_after();
    issue_s(0);
    umma_commit(s_full);
    for (int j = 0; j < nkv; ++j) {
      const int s = j & 1;
      if (j + 1 < nkv) {            // S(j+1) as soon as the softmax warps hold S(j) in registers
        const int s2 = (j + 1) & 1;
        mbar_wait(s_free, j & 1);
        mbar_wait(&kv_full[s2], ((j + 1) >> 1) & 1);
        tc_fence_after();
        issue_s(s2);
        umma_commit(s_full);
      }
      mbar_wait(p_ready, j & 1);
      tc_fence_after();
      issue_o(s, j > 0);
      umma_commit(o_full);
      umma_commit(&kv_empty[s]);
    }
  } else if (warp >= 2) {
    // ------------------------------------------------------------- softmax: TWO threads per query row
    // Warps 2-5 take key columns [0, BKV/2) of their TMEM lane quarter, warps 6-9 columns [BKV/2, BKV) of the same
    // quarter (a warp may only touch lanes 32 (warp % 4) ...).  Four softmax warps per scheduler (two CTAs per SM)
    // instead of two: the MUFU.EX2 stream of one warp overlaps the TMEM loads, row-max chains and P stores of the
    // others (r1: 44 % MUFU utilisation with one thread per row), and 64 instead of 128 score registers per thread end
    // the spills.  The partners agree on the row max through two spare TMEM columns per tile parity (shared memory is
    // full at two CTAs per SM) and on the row sum once at the end.
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    constexpr int HC = BKV / 2;                    // score columns per thread
    constexpr int OC = kHeadDim / 2;               // output columns per thread
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    const uint32_t tmem_X = tmem_base + BKV + kHeadDim + lane_off;     // exchange columns: [2 tile parities][2 halves] max, then [2] sums
    auto pair_sync = [&]() {                       // named barrier of this warp pair (64 threads), id = 1 + q as an immediate
      if (q == 0) asm volatile("bar.sync 1, 64;" ::: "memory");
      else if (q == 1) asm volatile("bar.sync 2, 64;" ::: "memory");
      else if (q == 2) asm volatile("bar.sync 3, 64;" ::: "memory");
      else asm volatile("bar.sync 4, 64;" ::: "memory");
    };
    float m_used = -INFINITY, l_run = 0.f;
    const float sc = p.scale_log2;
    for (int j = 0; j < nkv; ++j) {
      const int kv_valid = min(BKV, p.kv_tokens - j * BKV);
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      uint32_t v[HC];
#pragma unroll
      for (int c = 0; c < HC; c += 32) tmem_ld32(tmem_S + lane_off + half * HC + c, v + c);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(s_free);                       // the S columns may be overwritten by S(j+1)
      if (kv_valid != BKV) {                     // tile-uniform: only the last tile of a ragged sequence is masked
#pragma unroll
        for (int i = 0; i < HC; ++i)
          if (half * HC + i >= kv_valid) v[i] = 0xFF800000u;   // -inf
      }
      float mx0 = __uint_as_float(v[0]), mx1 = __uint_as_float(v[1]);
#pragma unroll
      for (int i = 2; i < HC; i += 2) { mx0 = fmaxf(mx0, __uint_as_float(v[i])); mx1 = fmaxf(mx1, __uint_as_float(v[i + 1])); }
      float mx = fmaxf(mx0, mx1);
      {   // row max over both halves
        const uint32_t xc = tmem_X + 2 * (j & 1);
        tmem_st1(xc + half, __float_as_uint(mx));
        tmem_st_wait();
        tc_fence_before();
        pair_sync();
        tc_fence_after();
        const uint32_t other = tmem_ld1(xc + (half ^ 1));
        tmem_ld_wait();
        mx = fmaxf(mx, __uint_as_float(other));
      }
      // lazy running max: rescale only when this row's max grew by more than 2^8 relative to the max in use
      // (both partners see the same mx and m_used, so they take the same branch)
      const bool need = (mx - m_used) * sc > 8.f;             // true on the first tile (m_used = -inf)
      bool o_waited = false;
      if (__any_sync(0xffffffffu, need)) {
        const float alpha = need ? ((m_used == -INFINITY) ? 0.f : exp2f((m_used - mx) * sc)) : 1.f;
        if (need) m_used = mx;
        l_run *= alpha;
        if (j > 0) {               // O(j-1) is complete: read - scale - write back this thread's half of the row
          mbar_wait(o_full, (j - 1) & 1);
          tc_fence_after();
          o_waited = true;
          uint32_t o[OC];
          tmem_ld32(tmem_O + lane_off + half * OC, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < OC; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st32(tmem_O + lane_off + half * OC, o);
          tmem_st_wait();
        }
      }
      const float moff = m_used * sc;
      float rs0 = 0.f, rs1 = 0.f;
      // probabilities (fp32), row sum on the un-rounded values (two independent chains); the operand-precision rounding
      // of P is zero-mean, so the normaliser differs from sum(round(P)) by ~1e-4 relative at most.  The packed values
      // go to their OWN register array: packing in place (v[i / 2] = pack(v[i], v[i + 1])) demoted v[] to local memory.
      constexpr int PW = sizeof(T) == 2 ? HC / 2 : HC;        // 32-bit words of this thread's P half row
      uint32_t pk[PW];
#pragma unroll
      for (int i = 0; i < HC; i += 2) {
        const float p0 = fast_exp2(fmaf(__uint_as_float(v[i]), sc, -moff)), p1 = fast_exp2(fmaf(__uint_as_float(v[i + 1]), sc, -moff));
        rs0 += p0; rs1 += p1;
        if constexpr (sizeof(T) == 2) {
          __nv_bfloat162 h2 = __floats2bfloat162_rn(p0, p1);
          pk[i >> 1] = *reinterpret_cast<uint32_t*>(&h2);
        } else {
          pk[i] = __float_as_uint(from_f32<float>(p0)); pk[i + 1] = __float_as_uint(from_f32<float>(p1));
        }
      }
      l_run += rs0 + rs1;
      if (j > 0 && !o_waited) mbar_wait(o_full, (j - 1) & 1);   // P V(j-1) no longer reads the P tile
      // this thread's half row of the swizzled K-major P tile (operand precision): exactly one 128-byte atom row
      {
        uint8_t* prow = sP + half * 128 * 128 + r * 128;
#pragma unroll
        for (int cc = 0; cc < 8; ++cc)
          *reinterpret_cast<uint4*>(prow + ((cc ^ (r & 7)) * 16)) = make_uint4(pk[cc * 4], pk[cc * 4 + 1], pk[cc * 4 + 2], pk[cc * 4 + 3]);
      }
      tc_fence_before();
      fence_proxy_async();
      mbar_arrive(p_ready);
    }
    // O is complete in TMEM: combine the two half-row sums, normalise and store this thread's half of the head
    mbar_wait(o_full, (nkv - 1) & 1);
    tc_fence_after();
    {
      const uint32_t xc = tmem_X + 4;
      tmem_st1(xc + half, __float_as_uint(l_run));
      tmem_st_wait();
      tc_fence_before();
      pair_sync();
      tc_fence_after();
      const uint32_t other = tmem_ld1(xc + (half ^ 1));
      tmem_ld_wait();
      l_run += __uint_as_float(other);
    }
    const float inv = 1.f / l_run;
    const bool valid = q0 + r < p.n_tokens;
    T* dst = p.out + ((size_t)b * p.n_tokens + q0 + r) * 512 + h * kHeadDim + half * OC;
    {
      uint32_t o[OC];
      tmem_ld32(tmem_O + lane_off + half * OC, o);
      tmem_ld_wait();
      if (valid) {
        if constexpr (sizeof(T) == 2) {
#pragma unroll
          for (int c8 = 0; c8 < OC; c8 += 8) {
            uint32_t w[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(o[c8 + 2 * t]) * inv, __uint_as_float(o[c8 + 2 * t + 1]) * inv);
              w[t] = *reinterpret_cast<uint32_t*>(&h2);
            }
            *reinterpret_cast<uint4*>(dst + c8) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        } else {
#pragma unroll
          for (int c4 = 0; c4 < OC; c4 += 4)
            *reinterpret_cast<float4*>(dst + c4) =
                make_float4(__uint_as_float(o[c4]) * inv, __uint_as_float(o[c4 + 1]) * inv, __uint_as_float(o[c4 + 2]) * inv,
                            __uint_as_float(o[c4 + 3]) * inv);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace sfb
