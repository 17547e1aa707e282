// K5b: fused self-attention core of AttentionItem (upstream a_unet Attention):
//   out[b, i, h, :] = softmax_j(q[b,i,h,:] . k[b,j,h,:] / sqrt(64)) v[b,j,h,:]     8 heads x 64, non-causal, no mask.
// Input is the fused QKV projection output [B, N, 1536] (q | k | v, heads contiguous 64-wide), output [B, N, 512].
//
// Two forms: attn_tc_kernel<T> (this one; 128-key tiles in bf16 / 64 in TF32, the whole score row in registers, two CTAs
// per SM - used by fp32 mode) and attn2_tc_kernel (bf16, further down: 64-key tiles, chunked two-pass softmax, four CTAs
// per SM - the bf16 path).
//
// One CTA = one 128-query tile of one (clip, head), two CTAs per SM.  Warp 0 lane 0: TMA producer (Q once, K/V tiles
// in a three-stage ring).  Warp 1 lane 0: tcgen05.mma issuer: S = Q K^T (K-major x K-major) into TMEM, O += P V
// ACCUMULATED IN TMEM over all key tiles (P is the A operand read FROM TENSOR MEMORY, V is consumed straight from its
// TMA tile as an MN-major B operand).  Warps 2-5 (128 threads, one query row each): S is read from TMEM ONCE into
// registers (which frees the S columns at once: the MMA warp issues S(j+1) while softmax(j) is still computing), row
// max (FMNMX3), exp2, P written back to its own TMEM columns in operand precision with tcgen05.st - no shared-memory P
// tile, no generic -> async proxy fence on the softmax -> MMA hand-off.  The running max is LAZY: the O accumulator and
// the row sum are rescaled only when a row's max grows by
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
  long long* dbg;     // optional timeline (CTA 0 only): [role][tile][8] clock64 stamps (tools/attn_timeline.py)
};
#ifdef SFB_ATTN_TL      // build with -DSFB_ATTN_TL (python -m syncfusion_b200.build with SFB_ATTN_TL=1) for the per-warp timeline
#define ATTN_STAMP(role, j, k) do { if (p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (j) < 32) p.dbg[((role) * 32 + (j)) * 8 + (k)] = clock64(); } while (0)
#else
#define ATTN_STAMP(role, j, k) do { } while (0)
#endif

template <typename T> struct AttnCfg;
template <> struct AttnCfg<__nv_bfloat16> { static constexpr int BKV = 64; };      // attn2_tc_kernel below
template <> struct AttnCfg<float> { static constexpr int BKV = 64; };

__device__ __forceinline__ float fmax3(float a, float b, float c) {   // FMNMX3
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float fast_exp2(float x) {   // single MUFU.EX2
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

constexpr int kAttnThreads = 192;
constexpr int kAttnStages = 3;      // K/V ring depth (P lives in TMEM, so shared memory holds only Q and the ring)
constexpr int kHeadDim = 64;

template <typename T>
__host__ __device__ constexpr int attn_smem_bytes() {
  constexpr int DA = kHeadDim / ElemTraits<T>::kAtomElems;
  constexpr int BKV = AttnCfg<T>::BKV;
  return DA * 128 * 128 /*Q*/ + kAttnStages * 2 * DA * BKV * 128 /*K,V stages*/ + 256;
}

template <typename T>
__global__ void __launch_bounds__(kAttnThreads, 2) attn_tc_kernel(const __grid_constant__ AttnParams<T> p) {
  pdl_trigger();
  using TR = ElemTraits<T>;
  constexpr int AE = TR::kAtomElems;            // elements per 128-byte row
  constexpr int DA = kHeadDim / AE;             // atoms along head dim (1 bf16, 2 tf32)
  constexpr int BKV = AttnCfg<T>::BKV;
  constexpr int UK = TR::kUmmaK;
  constexpr int kQBytes = DA * 128 * 128;
  constexpr int kKBytes = DA * BKV * 128;
  constexpr int NS = kAttnStages;
  constexpr int PC = BKV * (int)sizeof(T) / 4;  // 32-bit TMEM columns of a P row in operand precision (64)
  constexpr uint32_t kTmemCols = 256;           // S: [0, BKV), O: [BKV, BKV + 64), P: [BKV + 64, BKV + 64 + PC)
  static_assert(BKV + kHeadDim + PC <= 256, "TMEM budget");

  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + kQBytes;                  // stage s: K at sKV + s*2*kKBytes, V right after K
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + NS * 2 * kKBytes);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;    // [NS <= 4]
  uint64_t* kv_empty = bars + 5;   // [NS]
  uint64_t* s_full = bars + 9;
  uint64_t* p_ready = bars + 10;
  uint64_t* o_full = bars + 11;
  uint64_t* s_free = bars + 12;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int nkv = (p.kv_tokens + BKV - 1) / BKV;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmKV);
    tma_prefetch_desc(&p.tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < NS; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    mbar_init(s_full, 1);
    mbar_init(p_ready, 128);
    mbar_init(o_full, 1);
    mbar_init(s_free, 128);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, kTmemCols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  pdl_wait();            // everything above is independent of the previous kernel's output
  mark_progress(p.tag);
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + BKV, tmem_P = tmem_base + BKV + kHeadDim;

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------- TMA producer
    mbar_expect_tx(q_full, kQBytes);
    for (int a = 0; a < DA; ++a) tma_load_3d(sQ + a * 128 * 128, &p.tmQ, q_full, p.q_col0 + h * kHeadDim + a * AE, q0, b);
    int s = 0;
    uint32_t ph = 0;
    for (int j = 0; j < nkv; ++j) {
      mbar_wait(&kv_empty[s], ph ^ 1);
      mbar_expect_tx(&kv_full[s], 2 * kKBytes);
      uint8_t* sk = sKV + s * 2 * kKBytes;
      for (int a = 0; a < DA; ++a) {
        tma_load_3d(sk + a * BKV * 128, &p.tmKV, &kv_full[s], p.k_col0 + h * kHeadDim + a * AE, j * BKV, b);
        tma_load_3d(sk + kKBytes + a * BKV * 128, &p.tmV, &kv_full[s], p.v_col0 + h * kHeadDim + a * AE, j * BKV, b);
      }
      if (++s == NS) { s = 0; ph ^= 1; }
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
    auto issue_o = [&](int s, bool accumulate) {      // O (+)= P V: P from TENSOR MEMORY (A operand), V from its TMA tile
      const uint32_t av = smem_u32(sKV + s * 2 * kKBytes + kKBytes);
#pragma unroll
      for (int k = 0; k < BKV / UK; ++k) {
        // MN-major B: rows of the tile are keys (the MMA K dim) at 128-byte pitch; 8-key groups 1024 bytes apart
        // (SBO); head-dim atoms BKV*128 bytes apart (LBO, only used by the two-atom tf32 layout).
        // bf16: SWIZZLE_128B, 8-key groups; tf32: SWIZZLE_128B_BASE32B, 4-key groups 512 bytes apart.
        const uint64_t db = (sizeof(T) == 2) ? make_smem_desc(av + k * UK * 128, BKV * 128, 1024, 2)
                                             : make_smem_desc(av + k * UK * 128, BKV * 128, 512, 1);
        umma_ts<TR::kTF32>(tmem_O, tmem_P + k * 8, db, idesc_o, (accumulate || k != 0) ? 1u : 0u);   // one k step = 32 B of a P row = 8 columns
      }
    };
    mbar_wait(q_full, 0);
    mbar_wait(&kv_full[0], 0);
    tc_fence_after();
    issue_s(0);
    umma_commit(s_full);
    int s = 0, s2 = 0;
    uint32_t ph2 = 0;
    for (int j = 0; j < nkv; ++j) {
      if (j + 1 < nkv) {            // S(j+1) as soon as the softmax warps hold S(j) in registers
        if (++s2 == NS) { s2 = 0; ph2 ^= 1; }
        mbar_wait(s_free, j & 1);
        mbar_wait(&kv_full[s2], ph2);
        tc_fence_after();
        issue_s(s2);
        umma_commit(s_full);
      }
      ATTN_STAMP(4, j, 0);
      mbar_wait(p_ready, j & 1);
      ATTN_STAMP(4, j, 1);
      tc_fence_after();
      issue_o(s, j > 0);
      umma_commit(o_full);
      umma_commit(&kv_empty[s]);
      if (++s == NS) s = 0;
    }
  } else if (warp >= 2) {
    // ------------------------------------------------------------- softmax (one query row per thread)
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    float m_used = -INFINITY, l_run = 0.f;
    const float sc = p.scale_log2;
    for (int j = 0; j < nkv; ++j) {
      const int kv_valid = min(BKV, p.kv_tokens - j * BKV);
      if (lane == 0) ATTN_STAMP(warp - 2, j, 0);
      mbar_wait(s_full, j & 1);
      if (lane == 0) ATTN_STAMP(warp - 2, j, 1);
      tc_fence_after();
      uint32_t v[BKV];
#pragma unroll
      for (int c = 0; c < BKV; c += 32) tmem_ld32(tmem_S + lane_off + c, v + c);
      tmem_ld_wait();
      if (lane == 0) ATTN_STAMP(warp - 2, j, 2);
      tc_fence_before();
      mbar_arrive(s_free);                       // the S columns may be overwritten by S(j+1)
      if (kv_valid != BKV) {                     // tile-uniform: only the last tile of a ragged sequence is masked
#pragma unroll
        for (int i = 0; i < BKV; ++i)
          if (i >= kv_valid) v[i] = 0xFF800000u;   // -inf
      }
      float mx0 = __uint_as_float(v[0]), mx1 = __uint_as_float(v[1]), mx2 = __uint_as_float(v[2]), mx3 = __uint_as_float(v[3]);
#pragma unroll
      for (int i = 4; i < BKV; i += 8) {        // three-input max (FMNMX3), four independent chains
        mx0 = fmax3(mx0, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
        mx1 = fmax3(mx1, __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
        if (i + 4 < BKV) {
          mx2 = fmax3(mx2, __uint_as_float(v[i + 4]), __uint_as_float(v[i + 5]));
          mx3 = fmax3(mx3, __uint_as_float(v[i + 6]), __uint_as_float(v[i + 7]));
        }
      }
      const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      // lazy running max: rescale only when this row's max grew by more than 2^8 relative to the max in use
      const bool need = (mx - m_used) * sc > 8.f;             // true on the first tile (m_used = -inf)
      bool o_waited = false;
      if (__any_sync(0xffffffffu, need)) {
        const float alpha = need ? ((m_used == -INFINITY) ? 0.f : exp2f((m_used - mx) * sc)) : 1.f;
        if (need) m_used = mx;
        l_run *= alpha;
        if (j > 0) {               // O(j-1) is complete: read - scale - write back (whole warp, per-lane factor)
          mbar_wait(o_full, (j - 1) & 1);
          tc_fence_after();
          o_waited = true;
#pragma unroll
          for (int c = 0; c < kHeadDim; c += 32) {
            uint32_t o[32];
            tmem_ld32(tmem_O + lane_off + c, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st32(tmem_O + lane_off + c, o);
          }
          tmem_st_wait();
        }
      }
      if (lane == 0) ATTN_STAMP(warp - 2, j, 3);
      const float moff = m_used * sc;
      float rs0 = 0.f, rs1 = 0.f;
      // probabilities in place (fp32), row sum on the un-rounded values (two independent chains); the operand-precision
      // rounding of P is zero-mean, so the normaliser differs from sum(round(P)) by ~1e-4 relative at most.
#pragma unroll
      for (int i = 0; i < BKV; i += 2) {
        const float p0 = fast_exp2(fmaf(__uint_as_float(v[i]), sc, -moff)), p1 = fast_exp2(fmaf(__uint_as_float(v[i + 1]), sc, -moff));
        rs0 += p0; rs1 += p1;
        if constexpr (sizeof(T) == 2) {
          __nv_bfloat162 h2 = __floats2bfloat162_rn(p0, p1);
          v[i >> 1] = *reinterpret_cast<uint32_t*>(&h2);       // packed pair i/2 (slots below i are already consumed)
        } else {
          v[i] = __float_as_uint(from_f32<float>(p0)); v[i + 1] = __float_as_uint(from_f32<float>(p1));
        }
      }
      l_run += rs0 + rs1;
      if (lane == 0) ATTN_STAMP(warp - 2, j, 4);
      if (j > 0 && !o_waited) mbar_wait(o_full, (j - 1) & 1);
      if (lane == 0) ATTN_STAMP(warp - 2, j, 5);   // P V(j-1) no longer reads the P columns
      // this row's probabilities (operand precision) into the P columns of tensor memory: no shared-memory tile, no
      // generic -> async proxy fence; the P V MMA takes its A operand straight from TMEM
#pragma unroll
      for (int c = 0; c < PC; c += 32) tmem_st32(tmem_P + lane_off + c, v + c);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_ready);
      if (lane == 0) ATTN_STAMP(warp - 2, j, 6);
    }
    // O is complete in TMEM: normalise and store
    mbar_wait(o_full, (nkv - 1) & 1);
    tc_fence_after();
    const float inv = 1.f / l_run;
    const bool valid = q0 + r < p.n_tokens;
    T* dst = p.out + ((size_t)b * p.n_tokens + q0 + r) * 512 + h * kHeadDim;
#pragma unroll
    for (int c = 0; c < kHeadDim; c += 32) {
      uint32_t o[32];
      tmem_ld32(tmem_O + lane_off + c, o);
      tmem_ld_wait();
      if (valid) {
        if constexpr (sizeof(T) == 2) {
#pragma unroll
          for (int c8 = 0; c8 < 32; c8 += 8) {
            uint32_t w[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(o[c8 + 2 * t]) * inv, __uint_as_float(o[c8 + 2 * t + 1]) * inv);
              w[t] = *reinterpret_cast<uint32_t*>(&h2);
            }
            *reinterpret_cast<uint4*>(dst + c + c8) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        } else {
#pragma unroll
          for (int c4 = 0; c4 < 32; c4 += 4)
            *reinterpret_cast<float4*>(dst + c + c4) =
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


// ------------------------------------------------------------------------------------------------ attention, second form
// bf16 only (fp32 / TF32 mode keeps attn_tc_kernel above).  Same math as attn_tc_kernel, re-shaped for MORE RESIDENT WARPS per scheduler: the first form keeps a whole
// 128-key score row in registers (168 registers, two CTAs per SM, two softmax warps per scheduler running a serial
// TMEM-load -> max -> exp2 -> store chain at 61 % of the MUFU floor).  Here a key tile is 64 wide, the row is taken from
// tensor memory in 32-column chunks TWICE (pass 1: row max; pass 2: exp2, row sum, P) so a thread never holds more than
// one chunk, and P overwrites the S columns it came from (chunk c of P = 16 packed columns lands in S columns the same
// thread has already consumed).  TMEM per CTA: S | P 64 + O 64 = 128 columns; shared memory 48 KB; ~80 registers:
// FOUR CTAs per SM.  A CTA is strictly serial (S(j) -> softmax(j) -> P V(j) -> S(j+1)); the overlap comes from the
// other three CTAs on the SM.  With four softmax warps per scheduler the kernel is issue bound, not MUFU bound: computing
// a quarter / third / half of the exponentials with a cubic on the FMA pipe made it 5 / 10 / 16 % SLOWER (measured).
constexpr int kAttn2Threads = 192;
constexpr int kAttn2Stages = 2;       // two K/V stages: 48 KB per CTA (three stages at three CTAs per SM measured 13 % slower)
constexpr int kAttn2BKV = 64;
__host__ __device__ constexpr int attn2_smem_bytes() { return 128 * 128 /*Q*/ + kAttn2Stages * 2 * kAttn2BKV * 128 /*K, V*/ + 256; }

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

__global__ void __launch_bounds__(kAttn2Threads, 4) attn2_tc_kernel(const __grid_constant__ AttnParams<__nv_bfloat16> p) {
  pdl_trigger();
  constexpr int BKV = kAttn2BKV, NS = kAttn2Stages;
  constexpr int kQBytes = 128 * 128, kKBytes = BKV * 128;
  constexpr uint32_t kTmemCols = 128;           // S | P: [0, 64), O: [64, 128)
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + kQBytes;                  // stage s: K at sKV + s*2*kKBytes, V right after K
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + NS * 2 * kKBytes);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;    // [NS <= 4]
  uint64_t* kv_empty = bars + 5;   // [NS]
  uint64_t* s_full = bars + 9;
  uint64_t* p_ready = bars + 10;
  uint64_t* o_full = bars + 11;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int nkv = (p.kv_tokens + BKV - 1) / BKV;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmKV);
    mbar_init(q_full, 1);
    for (int s = 0; s < NS; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    mbar_init(s_full, 1);
    mbar_init(p_ready, 128);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, kTmemCols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  pdl_wait();
  mark_progress(p.tag);
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + BKV;

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------- TMA producer
    mbar_expect_tx(q_full, kQBytes);
    tma_load_3d(sQ, &p.tmQ, q_full, p.q_col0 + h * kHeadDim, q0, b);
    int s = 0;
    uint32_t ph = 0;
    for (int j = 0; j < nkv; ++j) {
      mbar_wait(&kv_empty[s], ph ^ 1);
      mbar_expect_tx(&kv_full[s], 2 * kKBytes);
      uint8_t* sk = sKV + s * 2 * kKBytes;
      tma_load_3d(sk, &p.tmKV, &kv_full[s], p.k_col0 + h * kHeadDim, j * BKV, b);
      tma_load_3d(sk + kKBytes, &p.tmV, &kv_full[s], p.v_col0 + h * kHeadDim, j * BKV, b);
      if (++s == NS) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1 && lane == 0) {
    // ------------------------------------------------------------- MMA issuer (the tensor pipe runs a CTA's MMAs in issue order)
    constexpr uint32_t idesc_s = make_idesc(1, 128, BKV, 0, 0);        // Q (K-major) x K (K-major)
    constexpr uint32_t idesc_o = make_idesc(1, 128, kHeadDim, 0, 1);   // P (TMEM) x V (MN-major)
    mbar_wait(q_full, 0);
    int s = 0;
    uint32_t ph = 0;
    for (int j = 0; j < nkv; ++j) {
      mbar_wait(&kv_full[s], ph);
      tc_fence_after();
      const uint32_t aq = smem_u32(sQ), ak = smem_u32(sKV + s * 2 * kKBytes), av = ak + kKBytes;
#pragma unroll
      for (int k = 0; k < kHeadDim / 16; ++k)       // S(j) = Q K(j)^T: overwrites P(j-1), which P V(j-1) (issued before) has consumed
        umma_ss<false>(tmem_S, make_smem_desc_sw128(aq + k * 32, 16, 1024), make_smem_desc_sw128(ak + k * 32, 16, 1024), idesc_s, k != 0);
      umma_commit(s_full);
      mbar_wait(p_ready, j & 1);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < BKV / 16; ++k)            // O (+)= P V: P from tensor memory (8 columns per k step), V MN-major from its TMA tile
        umma_ts<false>(tmem_O, tmem_S + k * 8, make_smem_desc(av + k * 16 * 128, BKV * 128, 1024, 2), idesc_o, (j > 0 || k != 0) ? 1u : 0u);
      umma_commit(&kv_empty[s]);
      if (j + 1 == nkv) umma_commit(o_full);
      if (++s == NS) { s = 0; ph ^= 1; }
    }
  } else if (warp >= 2) {
    // ------------------------------------------------------------- softmax (one query row per thread, 32-column chunks)
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    float m_used = -INFINITY, l_run = 0.f;
    const float sc = p.scale_log2;
    for (int j = 0; j < nkv; ++j) {
      const int kv_valid = min(BKV, p.kv_tokens - j * BKV);
      mbar_wait(s_full, j & 1);                  // S(j) complete - and with it P V(j-1): O may be rescaled, P may be overwritten
      tc_fence_after();
      // pass 1: row max
      float mx;
      {
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int c = 0; c < BKV; c += 32) {
          uint32_t v[32];
          tmem_ld32(tmem_S + lane_off + c, v);
          tmem_ld_wait();
          if (kv_valid != BKV) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c + i >= kv_valid) v[i] = 0xFF800000u;   // -inf
          }
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            mx0 = fmax3(mx0, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
            mx1 = fmax3(mx1, __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
            mx2 = fmax3(mx2, __uint_as_float(v[i + 4]), __uint_as_float(v[i + 5]));
            mx3 = fmax3(mx3, __uint_as_float(v[i + 6]), __uint_as_float(v[i + 7]));
          }
        }
        mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      }
      // lazy running max: rescale only when this row's max grew by more than 2^8 relative to the max in use
      const bool need = (mx - m_used) * sc > 8.f;             // true on the first tile (m_used = -inf)
      if (__any_sync(0xffffffffu, need)) {
        const float alpha = need ? ((m_used == -INFINITY) ? 0.f : exp2f((m_used - mx) * sc)) : 1.f;
        if (need) m_used = mx;
        l_run *= alpha;
        if (j > 0) {               // O(j-1) is complete (s_full): read - scale - write back (whole warp, per-lane factor)
#pragma unroll
          for (int c = 0; c < kHeadDim; c += 32) {
            uint32_t o[32];
            tmem_ld32(tmem_O + lane_off + c, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st32(tmem_O + lane_off + c, o);
          }
          tmem_st_wait();
        }
      }
      // pass 2: probabilities, row sum (un-rounded values), P in operand precision over the consumed S columns
      const float moff = m_used * sc;
      float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
      for (int c = 0; c < BKV; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_S + lane_off + c, v);
        tmem_ld_wait();
        if (kv_valid != BKV) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c + i >= kv_valid) v[i] = 0xFF800000u;
        }
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float p0 = fast_exp2(fmaf(__uint_as_float(v[i]), sc, -moff)), p1 = fast_exp2(fmaf(__uint_as_float(v[i + 1]), sc, -moff));
          rs0 += p0; rs1 += p1;
          __nv_bfloat162 h2 = __floats2bfloat162_rn(p0, p1);
          pk[i >> 1] = *reinterpret_cast<uint32_t*>(&h2);
        }
        tmem_st16(tmem_S + lane_off + (c >> 1), pk);          // P columns [c / 2, c / 2 + 16): S columns this thread has already read
      }
      l_run += rs0 + rs1;
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_ready);
    }
    // O is complete in TMEM: normalise and store
    mbar_wait(o_full, 0);
    tc_fence_after();
    const float inv = 1.f / l_run;
    const bool valid = q0 + r < p.n_tokens;
    __nv_bfloat16* dst = p.out + ((size_t)b * p.n_tokens + q0 + r) * 512 + h * kHeadDim;
#pragma unroll
    for (int c = 0; c < kHeadDim; c += 32) {
      uint32_t o[32];
      tmem_ld32(tmem_O + lane_off + c, o);
      tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int c8 = 0; c8 < 32; c8 += 8) {
          uint32_t w[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(o[c8 + 2 * t]) * inv, __uint_as_float(o[c8 + 2 * t + 1]) * inv);
            w[t] = *reinterpret_cast<uint32_t*>(&h2);
          }
          *reinterpret_cast<uint4*>(dst + c + c8) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// launch of the attention core of one plan op / debug call (bf16: the second form when it is compiled in)
template <typename T>
inline cudaError_t attn_set_attrs() {
  if constexpr (sizeof(T) == 2) return cudaFuncSetAttribute(attn2_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, attn2_smem_bytes());
  else return cudaFuncSetAttribute(attn_tc_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, attn_smem_bytes<T>());
}
template <typename T>
inline void attn_launch(dim3 grid, cudaStream_t st, const AttnParams<T>& p) {
  if constexpr (sizeof(T) == 2) launch_pdl(attn2_tc_kernel, grid, kAttn2Threads, attn2_smem_bytes(), st, p);
  else launch_pdl(attn_tc_kernel<T>, grid, kAttnThreads, attn_smem_bytes<T>(), st, p);
}

}  // namespace sfb
