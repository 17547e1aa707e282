// Resident-weight fused tcgen05 kernels for the HBM-bound depths (C <= 128; SURVEY.md 0.5: 48-192 flop/B).
//
// One persistent CTA per SM keeps ALL weights of the op resident in shared memory and streams 128-position tiles of
// one clip through a four-role pipeline, so every activation tensor crosses HBM exactly once per op and the
// normalisation / activation / modulation passes of the reference (a6, a7, a8) never exist as separate kernels:
//
//   warp 0      TMA producer   raw A tile [136 rows x K1] (rows l0-1 .. l0+134; the conv halo and the clip edges are
//                              TMA zero fill), residual tile [128 x N] fp32, onset-context tile [128 x ctx]
//   warps 2-5   A transform    GroupNorm(8) apply + SiLU on the raw tile -> bf16 UMMA operand (128B-swizzled K-major);
//                              the statistics come from the PRODUCER kernel's epilogue (fp64 sums)
//   warp 1      MMA issuer     k=3 conv = three ROW-SHIFTED descriptors (+128 B per tap) of the SAME smem tile, so the
//                              A tile is loaded and transformed once; fp32 accumulators double-buffered in TMEM; the
//                              chained second MMA (InjectChannels 1x1 over [m | ctx]) runs one tile behind
//   warps 6-9   epilogue       one accumulator row per thread: +bias +residual -> per-position LayerNorm over C with
//                              the step's Modulation (1+scale, shift) -> bf16 operand of the chained MMA -> +bias +m
//                              +cross-attention bias -> fp32 tile written IN PLACE over the residual tile and stored
//                              with one TMA store; GroupNorm statistics of the output are accumulated per thread in
//                              fp32 over the CTA's tiles of a clip and flushed as fp64 atomics once per clip.
//
// Instantiations (bf16 mode): ResNet conv1 (R1), ResNet conv2 + Modulation + Inject chain (R2), patchify Down, Up.
#pragma once
#include "ptx.cuh"
#undef SFB_FILE_ID
#define SFB_FILE_ID 3   // rk_tc.cuh

namespace sfb {

struct RkParams {
  CUtensorMap tmA;   // raw A: bf16 [K1, L, B] box [64, ROWS_A, 1]  |  fp32 [K1, L, B] box [32, ROWS_A, 1]
  CUtensorMap tmW;   // packed weight tiles bf16 [64, n_tiles * N] box [64, N]
  CUtensorMap tmR;   // residual in / fp32 out [N, L, B] box [32, 128, 1]
  CUtensorMap tmT;   // bf16 out [N, L, B] box [64, 128, 1]
  CUtensorMap tmC;   // onset context bf16 [ctx, L, Bc] box [ctx, 128, 1], no swizzle (dense rows)
  const double* stats_in;   // [B, 8, 2] GroupNorm sums of the raw A tensor
  double* stats_out;        // [B, 8, 2] sums of the output (nullable)
  const float* gamma;       // [K1] GroupNorm affine of the A transform
  const float* beta;
  const float* bias;        // [BMOD]
  const float* colscale;    // [BMOD] (nullable) SkipModulate scale of this step, row (b % cs_bmod) * cs_bstride
  const float* rowvec;      // [B, rowvec_stride] (nullable) cross-attention bias (added to the final output)
  const float* mod;         // [2N] Modulation scale | shift of this step, row (b % mod_bmod) * mod_bstride
  const float* bias2;       // [N] inject bias (chain)
  int cs_bstride, cs_bmod, rowvec_stride, mod_bstride, mod_bmod;
  int L, tiles_per_clip, total_tiles, ctx_bmod, ctx_ch;
  int has_resid, has_out_r, has_out_t;
  float eps;
  int tag;                  // plan op index (wait log)
};

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16_rn(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// SiLU of two values on the FMA pipe + HALF a MUFU op each.  MUFU is the scarce unit here (measured on B200: a
// [136 x 64] tile of tanh.approx / ex2+rcp costs 2200-3200 cycles per SM, more than the tile's 1536 MMA cycles):
//   silu(z) = z / (1 + 2^(-z log2 e));  2^t for two elements with ONE packed ex2.approx.ftz.bf16x2 (sigma = 1 / (1 + e) is
//   well conditioned in e, so the bf16 precision of e costs <= 2^-9 relative, the precision of the bf16 result), and the
//   reciprocal by a bit-trick seed + two Newton steps (relative error ~2e-4) instead of MUFU.RCP.
__device__ __forceinline__ void silu2(float z0, float z1, float& y0, float& y1) {
  const float t0 = fminf(z0 * -1.4426950408889634f, 64.f), t1 = fminf(z1 * -1.4426950408889634f, 64.f);
  uint32_t pk, e;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk) : "f"(t1), "f"(t0));
  asm("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(e) : "r"(pk));
  const float d0 = 1.f + __uint_as_float(e << 16), d1 = 1.f + __uint_as_float(e & 0xFFFF0000u);
  float r0 = __uint_as_float(0x7EF311C7u - __float_as_uint(d0)), r1 = __uint_as_float(0x7EF311C7u - __float_as_uint(d1));
  r0 = r0 * fmaf(-d0, r0, 2.f); r1 = r1 * fmaf(-d1, r1, 2.f);
  r0 = r0 * fmaf(-d0, r0, 2.f); r1 = r1 * fmaf(-d1, r1, 2.f);
  y0 = z0 * r0; y1 = z1 * r1;
}
// Packed bf16x2 SiLU(GroupNorm(x)) for UMMA operand tiles: h = x * a' + b' (a' = a / 2, b' = b / 2 split hi + lo so the
// per-channel offset keeps fp32-like precision), y = h * tanh(h) + h.  Three FMA-pipe ops and one MUFU op per PAIR of
// elements, no unpack / pack: the transform warps are instruction-issue bound (measured: ~4 cycles per instruction per
// warp), and the fp32 form costs ~11 instructions per element against ~2 here.
__device__ __forceinline__ uint32_t silu_gn_bf16x2(uint32_t x, uint32_t pa, uint32_t pbh, uint32_t pbl) {
  uint32_t h, t, y;
  asm("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(h) : "r"(x), "r"(pa), "r"(pbh));
  asm("add.rn.bf16x2 %0, %1, %2;" : "=r"(h) : "r"(h), "r"(pbl));
  asm("tanh.approx.bf16x2 %0, %1;" : "=r"(t) : "r"(h));
  asm("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(y) : "r"(h), "r"(t), "r"(h));
  return y;
}
// (a, b) fp32 -> packed halves for channel pair (c, c + 1): a' = a / 2, b' = b / 2 = hi + lo
__device__ __forceinline__ void gn_pack_coef(float a0, float b0, float a1, float b1, uint32_t& pa, uint32_t& pbh, uint32_t& pbl) {
  pa = pack_bf16_rn(0.5f * a0, 0.5f * a1);
  const __nv_bfloat16 h0 = __float2bfloat16_rn(0.5f * b0), h1 = __float2bfloat16_rn(0.5f * b1);
  pbh = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
  pbl = pack_bf16_rn(0.5f * b0 - __bfloat162float(h0), 0.5f * b1 - __bfloat162float(h1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// XF: 0 = A is operand-ready bf16, 1 = bf16 raw + GN/SiLU in place, 2 = fp32 raw + GN/SiLU -> bf16 operand buffer
// EPI: 0 = plain (bias, colscale, rowvec, resid), 1 = + LayerNorm/Modulation, 2 = LN/Mod + chained inject MMA
// USE_R: the op has a residual input and/or an fp32 output (R ring allocated)
template <int K1, int N, int TAPS, int XF, int EPI, int BMOD, int USE_R>
struct RkCfg {
  static_assert(K1 % 32 == 0 && N % 32 == 0 && N <= 256, "shape");
  static_assert(EPI != 2 || (K1 == N && N <= 64 && USE_R), "chain needs K2 = N = K1 <= 64");
  static_assert(EPI == 0 || USE_R, "LayerNorm epilogue works in place on the residual tile");
  static constexpr int KA = (K1 + 63) / 64;
  static constexpr int ROWS_A = TAPS == 3 ? 136 : 128;
  static constexpr int ATOM_A = ROWS_A * 128;
  static constexpr int OP_BYTES = KA * ATOM_A;
  static constexpr int RAW_ATOMS = K1 / 32;
  static constexpr int RAW_BYTES = XF == 2 ? RAW_ATOMS * ATOM_A : 0;
  static constexpr int W1_TILES = TAPS * KA;
  static constexpr int KA2 = EPI == 2 ? (N + 64) / 64 : 0;   // chained K = [m (N) | ctx]: ctx <= 32 if N == 32, <= 64 if N == 64
  static constexpr int W_BYTES = (W1_TILES + KA2) * N * 128;
  static constexpr int RA = N / 32;
  static constexpr int R_BYTES = USE_R ? RA * 128 * 128 : 0;
  static constexpr int CTX_BYTES = EPI == 2 ? 4096 : 0;   // dense [128 rows][ctx * 2 B]: ctx <= 16
  static constexpr int TA = (N + 63) / 64;
  static constexpr int T_BYTES = EPI == 2 ? 0 : TA * 128 * 128;   // chain: the bf16 output copy aliases the A2 buffer
  static constexpr int A2_BYTES = KA2 * 128 * 128;
  static constexpr int VEC_FLOATS = 4 * N;     // ep_mul | ep_add | ep_g | ep_sh
  static constexpr int GS = BMOD / 8;
  static constexpr int kThreads = XF ? 320 : 192;
  static constexpr int EPI_WARP0 = XF ? 6 : 2;
  static constexpr int TMEM_NEED = EPI == 2 ? 3 * N : 2 * N;
  static constexpr int TMEM_COLS = TMEM_NEED <= 32 ? 32 : TMEM_NEED <= 64 ? 64 : TMEM_NEED <= 128 ? 128 : TMEM_NEED <= 256 ? 256 : 512;
  // CTAs per SM.  The depth-1 items (C = 32) are bound by the serial chain of their single epilogue warpgroup (wait
  // accumulator -> TMEM loads -> LayerNorm -> chained MMA round trip -> stage -> fence -> TMA store -> wait for the
  // store to drain the tile: ~5000 cycles per [128 x 32] tile at 43 % of HBM bandwidth): two independent CTAs per SM
  // double the tiles in flight with no change to the pipeline protocol (each gets half the ring budget).
#ifdef SFB_RK_NO_OCC2
  static constexpr int OCC = 1;
#else
  static constexpr int OCC = (K1 == 32 && N == 32 && TAPS == 3 && XF != 0) ? 2 : 1;
#endif
  // ring depths from the shared-memory budget: these kernels are HBM-latency bound (ncu: every role waits on TMA data at
  // 28 % DRAM utilisation with two A stages), so whatever the resident weights leave goes into bytes in flight.
  static constexpr int kSmemCap = OCC == 2 ? 115712 : 232448;          // 2 x (113 KB + 1 KB reserved) <= 228 KB per SM
  static constexpr int NT = OCC == 2 ? 1 : 2;    // bf16 staging tiles (one: the store is drained before the next tile is staged)
  static constexpr int kBudget = kSmemCap - (W_BYTES + NT * T_BYTES + A2_BYTES + VEC_FLOATS * 4 + 256 + 1024);
  static constexpr int PER_A = OP_BYTES + RAW_BYTES;
  static constexpr int PER_R = R_BYTES + CTX_BYTES;
  static constexpr int NSR_FIT = PER_R > 0 ? (kBudget - 2 * PER_A) / PER_R : 3;
  // two residual stages are enough only where a tile's slot goes back to the producer as soon as its own store has
  // drained it (the chained epilogue, see the end of the tile loop); the other shapes release one tile late
  static constexpr int NSR_MIN = EPI == 2 ? 2 : 3;
  static constexpr int NSR = NSR_FIT >= 5 ? 5 : (NSR_FIT >= 4 ? 4 : (NSR_FIT >= 3 ? 3 : NSR_MIN));
  static constexpr int NSA_FIT = (kBudget - NSR * PER_R) / PER_A;
  static constexpr int NSA = NSA_FIT >= 4 ? 4 : (NSA_FIT >= 3 ? 3 : 2);
  // smem carve-up (every tile region is a multiple of 1024 B)
  static constexpr int OFF_W = 0;
  static constexpr int OFF_OP = OFF_W + W_BYTES;
  static constexpr int OFF_RAW = OFF_OP + NSA * OP_BYTES;
  static constexpr int OFF_R = OFF_RAW + NSA * RAW_BYTES;
  static constexpr int OFF_CTX = OFF_R + NSR * R_BYTES;
  static constexpr int OFF_T = OFF_CTX + NSR * CTX_BYTES;
  static constexpr int OFF_A2 = OFF_T + NT * T_BYTES;
  static constexpr int OFF_VEC = OFF_A2 + A2_BYTES;
  static constexpr int OFF_BAR = OFF_VEC + VEC_FLOATS * 4;
  static constexpr int SMEM = OFF_BAR + 256 + 1024 /*alignment slack*/;
  static_assert(W_BYTES % 1024 == 0 && OP_BYTES % 1024 == 0 && RAW_BYTES % 1024 == 0, "alignment");
  static_assert(SMEM <= kSmemCap, "shared memory budget");
};

template <int K1, int N, int TAPS, int XF, int EPI, int BMOD, int USE_R>
__global__ void __launch_bounds__(RkCfg<K1, N, TAPS, XF, EPI, BMOD, USE_R>::kThreads, RkCfg<K1, N, TAPS, XF, EPI, BMOD, USE_R>::OCC) rk_kernel(const __grid_constant__ RkParams p) {
  pdl_trigger();
  using C = RkCfg<K1, N, TAPS, XF, EPI, BMOD, USE_R>;
  constexpr int NSR = C::NSR;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW = smem + C::OFF_W;
  uint8_t* sOP = smem + C::OFF_OP;
  uint8_t* sRAW = smem + C::OFF_RAW;
  uint8_t* sR = smem + C::OFF_R;
  uint8_t* sCTX = smem + C::OFF_CTX;
  uint8_t* sT = smem + C::OFF_T;
  uint8_t* sA2 = smem + C::OFF_A2;
  float* sVEC = reinterpret_cast<float*>(smem + C::OFF_VEC);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  constexpr int NSA = C::NSA;
  uint64_t* w_full = bars + 0;
  uint64_t* a_full = bars + 1;       // [NSA <= 4]   TMA -> transform (XF) / MMA
  uint64_t* a_empty = bars + 5;      // [NSA]        MMA commit -> TMA
  uint64_t* op_full = bars + 9;      // [NSA]        transform -> MMA
  uint64_t* acc1_full = bars + 13;   // [2]   MMA commit -> epilogue
  uint64_t* acc1_empty = bars + 15;  // [2]   epilogue -> MMA
  uint64_t* a2_full = bars + 17;     //       epilogue -> MMA (chained operand written)
  uint64_t* acc2_full = bars + 18;   //       MMA commit -> epilogue
  uint64_t* r_full = bars + 19;      // [NSR <= 5]   TMA -> epilogue (residual + context tile)
  uint64_t* r_empty = bars + 24;     // [NSR]        epilogue (store has read the tile) -> TMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 29);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // contiguous tile range per CTA: a CTA stays inside one clip as long as possible (few statistic flushes)
  const int t_begin = (int)(((long long)p.total_tiles * blockIdx.x) / gridDim.x);
  const int t_end = (int)(((long long)p.total_tiles * (blockIdx.x + 1)) / gridDim.x);
  const bool use_r = p.has_resid || EPI == 2;   // the R ring carries TMA loads (else it is plain output staging)

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmW);
    mbar_init(w_full, 1);
    for (int s = 0; s < NSA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); mbar_init(&op_full[s], 128); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc1_full[s], 1); mbar_init(&acc1_empty[s], 128); }
    for (int s = 0; s < NSR; ++s) { mbar_init(&r_full[s], 1); mbar_init(&r_empty[s], 1); }
    mbar_init(a2_full, 128);
    mbar_init(acc2_full, 1);
    fence_barrier_init();
  }
  // zero the K padding that no TMA load / epilogue write ever touches
  if (XF == 2 && K1 < 64)
    for (int i = threadIdx.x; i < C::NSA * C::OP_BYTES / 16; i += C::kThreads) reinterpret_cast<uint4*>(sOP)[i] = make_uint4(0, 0, 0, 0);
  if (EPI == 2)
    for (int i = threadIdx.x; i < C::A2_BYTES / 16; i += C::kThreads) reinterpret_cast<uint4*>(sA2)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  if (warp == 1) { tmem_alloc(tmem_slot, C::TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  if (warp == 0 && lane == 0) {      // the resident weights are static: loaded under the previous kernel's tail
    mbar_expect_tx(w_full, C::W_BYTES);
    for (int j = 0; j < C::W1_TILES + C::KA2; ++j) tma_load_2d(sW + j * N * 128, &p.tmW, w_full, 0, j * N);
  }
  pdl_wait();            // everything above is independent of the previous kernel's output
  mark_progress(p.tag);
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------- TMA producer
      uint32_t i = 0;
      for (int t = t_begin; t < t_end; ++t, ++i) {
        const int b = t / p.tiles_per_clip;
        const int l0 = (t % p.tiles_per_clip) * 128;
        const int s = i % NSA;
        mbar_wait(&a_empty[s], ((i / NSA) & 1) ^ 1);
        if (XF == 2) {
          mbar_expect_tx(&a_full[s], C::RAW_BYTES);
          for (int a = 0; a < C::RAW_ATOMS; ++a)
            tma_load_3d(sRAW + s * C::RAW_BYTES + a * C::ATOM_A, &p.tmA, &a_full[s], a * 32, l0 - (TAPS == 3 ? 1 : 0), b);
        } else {
          mbar_expect_tx(&a_full[s], C::OP_BYTES);
          for (int a = 0; a < C::KA; ++a)
            tma_load_3d(sOP + s * C::OP_BYTES + a * C::ATOM_A, &p.tmA, &a_full[s], a * 64, l0 - (TAPS == 3 ? 1 : 0), b);
        }
        if (use_r) {
          const int rs = i % NSR;
          mbar_wait(&r_empty[rs], ((i / NSR) & 1) ^ 1);
          mbar_expect_tx(&r_full[rs], (p.has_resid ? C::R_BYTES : 0) + (EPI == 2 ? 128 * p.ctx_ch * 2 : 0));
          if (p.has_resid)
            for (int a = 0; a < C::RA; ++a) tma_load_3d(sR + rs * C::R_BYTES + a * 128 * 128, &p.tmR, &r_full[rs], a * 32, l0, b);
          if (EPI == 2) tma_load_3d(sCTX + rs * C::CTX_BYTES, &p.tmC, &r_full[rs], 0, l0, b % p.ctx_bmod);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------------------------------------- MMA issuer
      constexpr uint32_t idesc = make_idesc(1 /*bf16*/, 128, N, 0, 0);
      mbar_wait(w_full, 0);
      const int k2steps = (N + p.ctx_ch + 15) / 16;
      auto issue_mma2 = [&](uint32_t j) {   // chained inject MMA of this CTA's tile j: acc2 = [m | ctx] W_inj^T
        mbar_wait(a2_full, j & 1);
        tc_fence_after();
        const uint32_t tacc = tmem_base + 2 * N;
        const uint32_t a2 = smem_u32(sA2), w2 = smem_u32(sW + C::W1_TILES * N * 128);
        for (int k = 0; k < k2steps; ++k)
          umma_ss<false>(tacc, make_smem_desc_sw128(a2 + (k >> 2) * 128 * 128 + (k & 3) * 32, 16, 1024),
                         make_smem_desc_sw128(w2 + (k >> 2) * N * 128 + (k & 3) * 32, 16, 1024), idesc, k ? 1u : 0u);
        umma_commit(acc2_full);
      };
      uint32_t i = 0;
      for (int t = t_begin; t < t_end; ++t, ++i) {
        const int s = i & 1, sa = i % NSA;
        mbar_wait(XF ? &op_full[sa] : &a_full[sa], (i / NSA) & 1);
        mbar_wait(&acc1_empty[s], ((i >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tacc = tmem_base + s * N;
        const uint32_t opb = smem_u32(sOP + sa * C::OP_BYTES), wb = smem_u32(sW);
        bool first = true;
#pragma unroll
        for (int tap = 0; tap < TAPS; ++tap) {
#pragma unroll
          for (int a = 0; a < C::KA; ++a) {
            const int ks = (K1 - a * 64) >= 64 ? 4 : (K1 - a * 64) / 16;
            for (int k = 0; k < ks; ++k) {
              // tap = +128 B (one row) on the start address of the same swizzled tile; k-step = +32 B inside the row
              umma_ss<false>(tacc, make_smem_desc_sw128(opb + a * C::ATOM_A + tap * 128 + k * 32, 16, 1024),
                             make_smem_desc_sw128(wb + (tap * C::KA + a) * N * 128 + k * 32, 16, 1024), idesc, first ? 0u : 1u);
              first = false;
            }
          }
        }
        umma_commit(&a_empty[sa]);
        umma_commit(&acc1_full[s]);
        if (EPI == 2 && i > 0) issue_mma2(i - 1);
      }
      if (EPI == 2 && i > 0) issue_mma2(i - 1);
    }
  } else if (XF && warp < 6) {
    // ------------------------------------------------------------- A transform: GroupNorm apply + SiLU -> bf16 operand
    const int tid = threadIdx.x - 64;           // 0..127
    constexpr int G = K1 / 8;                    // 8-channel chunks per row
    constexpr int RPI = 128 / G;                 // rows per iteration
    const int g = tid % G, rsub = tid / G;
    constexpr int GSA = K1 / 8;                  // GroupNorm group size of the A tensor (8 groups)
    float ca[8], cb[8];
    uint32_t pa[4], pbh[4], pbl[4];
    int cur_b = -1;
    uint32_t i = 0;
    for (int t = t_begin; t < t_end; ++t, ++i) {
      const int b = t / p.tiles_per_clip;
      const int l0 = (t % p.tiles_per_clip) * 128;
      const int s = i % NSA;
      if (b != cur_b) {
        cur_b = b;
        const double inv_cnt = 1.0 / ((double)p.L * GSA);   // fp64 only where E[x^2] - mean^2 cancels; no fp64 div / sqrt per channel
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int ch = g * 8 + j, grp = ch / GSA;
          const double s1 = p.stats_in[(size_t)b * 16 + grp * 2], s2 = p.stats_in[(size_t)b * 16 + grp * 2 + 1];
          const double mean = s1 * inv_cnt;
          const double var = fma(-mean, mean, s2 * inv_cnt);
          const float a = rsqrtf(fmaxf((float)var, 0.f) + p.eps) * __ldg(&p.gamma[ch]);
          ca[j] = a;
          cb[j] = __ldg(&p.beta[ch]) - (float)mean * a;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) gn_pack_coef(ca[2 * j], cb[2 * j], ca[2 * j + 1], cb[2 * j + 1], pa[j], pbh[j], pbl[j]);
      }
      mbar_wait(&a_full[s], (i / NSA) & 1);
      uint8_t* op = sOP + s * C::OP_BYTES + (g / 8) * C::ATOM_A;
      const int co = g % 8;
      constexpr int NIT = (C::ROWS_A + RPI - 1) / RPI;
      // load every row first, then compute, then store: independent chains overlap the LDS / MUFU latencies
      uint4 u0[NIT], u1[NIT];
#pragma unroll
      for (int it = 0; it < NIT; ++it) {
        const int r = it * RPI + rsub;
        if (r < C::ROWS_A) {
          if (XF == 2) {
            const uint8_t* raw = sRAW + s * C::RAW_BYTES + (g / 4) * C::ATOM_A + r * 128;
            const int c0 = (g % 4) * 2;
            u0[it] = *reinterpret_cast<const uint4*>(raw + ((c0 ^ (r & 7)) << 4));
            u1[it] = *reinterpret_cast<const uint4*>(raw + (((c0 + 1) ^ (r & 7)) << 4));
          } else {
            u0[it] = *reinterpret_cast<const uint4*>(op + r * 128 + ((co ^ (r & 7)) << 4));
          }
        }
      }
#pragma unroll
      for (int it = 0; it < NIT; ++it) {
        const int r = it * RPI + rsub;
        const int l = l0 - (TAPS == 3 ? 1 : 0) + r;
        const bool ok = l >= 0 && l < p.L;       // conv zero padding applies AFTER the activation: out-of-clip rows stay zero
        if (XF == 1) {      // bf16 raw tile: packed bf16x2 path, in place
          const uint32_t w[4] = {u0[it].x, u0[it].y, u0[it].z, u0[it].w};
          uint32_t o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) o[j] = ok ? silu_gn_bf16x2(w[j], pa[j], pbh[j], pbl[j]) : 0u;
          u0[it] = make_uint4(o[0], o[1], o[2], o[3]);
          continue;
        }
        float v[8];
        if (XF == 2) {
          v[0] = __uint_as_float(u0[it].x); v[1] = __uint_as_float(u0[it].y); v[2] = __uint_as_float(u0[it].z); v[3] = __uint_as_float(u0[it].w);
          v[4] = __uint_as_float(u1[it].x); v[5] = __uint_as_float(u1[it].y); v[6] = __uint_as_float(u1[it].z); v[7] = __uint_as_float(u1[it].w);
        } else {
          const uint32_t w[4] = {u0[it].x, u0[it].y, u0[it].z, u0[it].w};
#pragma unroll
          for (int j = 0; j < 4; ++j) { v[2 * j] = __uint_as_float(w[j] << 16); v[2 * j + 1] = __uint_as_float(w[j] & 0xFFFF0000u); }
        }
        float y[8];
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          silu2(fmaf(v[j], ca[j], cb[j]), fmaf(v[j + 1], ca[j + 1], cb[j + 1]), y[j], y[j + 1]);
          y[j] = ok ? y[j] : 0.f;
          y[j + 1] = ok ? y[j + 1] : 0.f;
        }
        u0[it] = make_uint4(pack_bf16(y[0], y[1]), pack_bf16(y[2], y[3]), pack_bf16(y[4], y[5]), pack_bf16(y[6], y[7]));
      }
#pragma unroll
      for (int it = 0; it < NIT; ++it) {
        const int r = it * RPI + rsub;
        if (r < C::ROWS_A) *reinterpret_cast<uint4*>(op + r * 128 + ((co ^ (r & 7)) << 4)) = u0[it];
      }
      fence_proxy_async();
      mbar_arrive(&op_full[s]);
    }
  } else if (warp >= C::EPI_WARP0) {
    // ------------------------------------------------------------- epilogue: one accumulator row per thread
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int et = threadIdx.x - C::EPI_WARP0 * 32;     // 0..127
    const bool elected = et == 0;
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    const int sw = row & 7;
    float* ep_mul = sVEC;            // [N]  colscale (plain) | inject bias + cross-attention bias (chain, final add)
    float* ep_add = sVEC + N;        // [N]  bias * colscale + rowvec (plain) | conv bias (LN)
    float* ep_g = sVEC + 2 * N;      // [N]  1 + modulation scale
    float* ep_sh = sVEC + 3 * N;     // [N]  modulation shift
    float s1[8], s2[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { s1[k] = 0.f; s2[k] = 0.f; }
    int cur_b = -1;
    auto flush_stats = [&](int b) {
      if (p.stats_out == nullptr || b < 0) return;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float a = s1[k], c = s2[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); c += __shfl_xor_sync(0xffffffffu, c, o); }
        if (lane == 0) {
          atomicAdd(&p.stats_out[(size_t)b * 16 + k * 2], (double)a);
          atomicAdd(&p.stats_out[(size_t)b * 16 + k * 2 + 1], (double)c);
        }
        s1[k] = 0.f; s2[k] = 0.f;
      }
    };
    auto add_stats = [&](const float (&y)[32], int c0) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int grp = ((c0 + j) % BMOD) / C::GS;
        s1[grp] += y[j];
        s2[grp] = fmaf(y[j], y[j], s2[grp]);
      }
    };
    auto store_bf16_chunk = [&](uint8_t* tile, const float (&y)[32], int c0) {   // 32 columns -> 4 swizzled 16-byte chunks
#pragma unroll
      for (int j8 = 0; j8 < 4; ++j8) {
        const int ch8 = (c0 % 64) / 8 + j8;
        *reinterpret_cast<uint4*>(tile + (c0 / 64) * 128 * 128 + row * 128 + ((ch8 ^ sw) << 4)) =
            make_uint4(pack_bf16(y[j8 * 8], y[j8 * 8 + 1]), pack_bf16(y[j8 * 8 + 2], y[j8 * 8 + 3]),
                       pack_bf16(y[j8 * 8 + 4], y[j8 * 8 + 5]), pack_bf16(y[j8 * 8 + 6], y[j8 * 8 + 7]));
      }
    };
    // Software-pipelined chain (EPI == 2 with >= 4 residual stages): the second epilogue of tile i-1 (acc2 of the chained
    // inject MMA) runs AFTER the first epilogue of tile i, so the MMA round trip (a2_full -> issue -> commit -> acc2_full)
    // is covered by LayerNorm work instead of idling the only epilogue warpgroup (ncu: 72 % no-eligible cycles).
#ifdef SFB_RK_PIPE
    constexpr bool PIPE = EPI == 2 && C::NSR >= 4;      // depth 1 (five residual stages); depth 2 has only three
#else
    // Parity-green and 2.6 % faster on the depth-1 inject at N = 1, but the only 2-GPU run of a build that had it on
    // hung (cause not established, GPU budget exhausted): off by default until a multi-GPU run clears it.
    constexpr bool PIPE = false;
#endif
    bool pend = false;
    int p_b = 0, p_l0 = 0, p_rs = 0;
    bool p_valid = false;
    uint32_t p_i = 0;
    auto finish_tile = [&](int fb, int fl0, int frs, bool fvalid, uint32_t fi) {
      uint8_t* rt = sR + frs * C::R_BYTES;
      uint8_t* tt = sA2;
      mbar_wait(acc2_full, fi & 1);
      tc_fence_after();
      const uint32_t tacc2 = tmem_base + 2 * N + lane_off;
#pragma unroll
      for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tacc2 + c0, v);
        tmem_ld_wait();
        float y[32];
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 ad = *reinterpret_cast<const float4*>(&ep_mul[c0 + j4 * 4]);
          uint8_t* slot = rt + (c0 / 32) * 128 * 128 + row * 128 + ((j4 ^ sw) << 4);
          const float4 mm = *reinterpret_cast<const float4*>(slot);
          y[j4 * 4 + 0] = __uint_as_float(v[j4 * 4 + 0]) + ad.x + mm.x;
          y[j4 * 4 + 1] = __uint_as_float(v[j4 * 4 + 1]) + ad.y + mm.y;
          y[j4 * 4 + 2] = __uint_as_float(v[j4 * 4 + 2]) + ad.z + mm.z;
          y[j4 * 4 + 3] = __uint_as_float(v[j4 * 4 + 3]) + ad.w + mm.w;
          *reinterpret_cast<float4*>(slot) = make_float4(y[j4 * 4], y[j4 * 4 + 1], y[j4 * 4 + 2], y[j4 * 4 + 3]);
        }
        if (p.has_out_t) store_bf16_chunk(tt, y, c0);    // aliases A2: the chained MMA has completed (acc2_full)
        if (p.stats_out != nullptr && fvalid) add_stats(y, c0);
      }
      tc_fence_before();
      fence_proxy_async();
      named_bar(1, 128);
      if (elected) {
        if (p.has_out_r)
          for (int a = 0; a < C::RA; ++a) tma_store_3d(&p.tmR, rt + a * 128 * 128, a * 32, fl0, fb);
        if (p.has_out_t)
          for (int a = 0; a < C::TA; ++a) tma_store_3d(&p.tmT, tt + a * 128 * 128, a * 64, fl0, fb);
        bulk_commit();
        if (p.has_out_t) bulk_wait_read<0>();   // A2 (aliased output copy) is rewritten right after
        else bulk_wait_read<1>();
        if (use_r && fi > 0) mbar_arrive(&r_empty[(fi - 1) % NSR]);
      }
      if (p.has_out_t) named_bar(1, 128);       // everybody waits for the elected thread's wait_read before A2 is rewritten
    };
    uint32_t i = 0;
    for (int t = t_begin; t < t_end; ++t, ++i) {
      const int b = t / p.tiles_per_clip;
      const int l0 = (t % p.tiles_per_clip) * 128;
      const int s = i & 1;
      const int rs = i % NSR;
      const bool row_valid = l0 + row < p.L;
      if (PIPE && pend && b != cur_b) {      // the pending tile belongs to the previous clip: finish it before its vectors go
        finish_tile(p_b, p_l0, p_rs, p_valid, p_i);
        pend = false;
      }
      // (1) top-of-tile barrier: the elected thread has confirmed (bulk_wait_read) that earlier TMA stores no longer
      //     read the buffers this tile rewrites; per-clip epilogue vectors are rebuilt when the clip changes.
      if (b != cur_b) {
        flush_stats(cur_b);
        named_bar(2, 128);           // nobody still reads the old vectors
        cur_b = b;
        for (int n = et; n < N; n += 128) {
          const int nm = n % BMOD;
          if (EPI == 0) {
            const float cs = p.colscale ? p.colscale[(size_t)(b % p.cs_bmod) * p.cs_bstride + nm] : 1.f;
            const float rv = p.rowvec ? p.rowvec[(size_t)b * p.rowvec_stride + nm] : 0.f;
            ep_mul[n] = cs;
            ep_add[n] = (p.bias ? p.bias[nm] : 0.f) * cs + rv;
          } else {
            const float* md = p.mod + (size_t)(b % p.mod_bmod) * p.mod_bstride;
            ep_add[n] = p.bias[nm];
            ep_g[n] = 1.f + md[n];
            ep_sh[n] = md[N + n];
            const float rv = p.rowvec ? p.rowvec[(size_t)b * p.rowvec_stride + n] : 0.f;
            ep_mul[n] = (EPI == 2 ? p.bias2[n] : 0.f) + rv;
          }
        }
      }
      named_bar(2, 128);
      uint8_t* rt = sR + rs * C::R_BYTES;      // fp32 [N/32 atoms][128 rows][128 B], 16-byte chunks XOR-swizzled by row & 7
      uint8_t* tt = EPI == 2 ? sA2 : sT + (s % C::NT) * C::T_BYTES;   // bf16 [N/64 atoms][128 rows][128 B]
      if (use_r) mbar_wait(&r_full[rs], (i / NSR) & 1);
      mbar_wait(&acc1_full[s], (i >> 1) & 1);
      tc_fence_after();
      const uint32_t tacc = tmem_base + s * N + lane_off;

      if (EPI == 0) {
#pragma unroll
        for (int c0 = 0; c0 < N; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(tacc + c0, v);
          tmem_ld_wait();
          float y[32];
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 mu = *reinterpret_cast<const float4*>(&ep_mul[c0 + j4 * 4]);
            const float4 ad = *reinterpret_cast<const float4*>(&ep_add[c0 + j4 * 4]);
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
            uint8_t* slot = rt + (c0 / 32) * 128 * 128 + row * 128 + ((j4 ^ sw) << 4);
            if (p.has_resid) x = *reinterpret_cast<const float4*>(slot);
            y[j4 * 4 + 0] = fmaf(__uint_as_float(v[j4 * 4 + 0]), mu.x, ad.x) + x.x;
            y[j4 * 4 + 1] = fmaf(__uint_as_float(v[j4 * 4 + 1]), mu.y, ad.y) + x.y;
            y[j4 * 4 + 2] = fmaf(__uint_as_float(v[j4 * 4 + 2]), mu.z, ad.z) + x.z;
            y[j4 * 4 + 3] = fmaf(__uint_as_float(v[j4 * 4 + 3]), mu.w, ad.w) + x.w;
            if (p.has_out_r) *reinterpret_cast<float4*>(slot) = make_float4(y[j4 * 4], y[j4 * 4 + 1], y[j4 * 4 + 2], y[j4 * 4 + 3]);
          }
          if (p.has_out_t) store_bf16_chunk(tt, y, c0);
          if (p.stats_out != nullptr && row_valid) add_stats(y, c0);
        }
        tc_fence_before();
        mbar_arrive(&acc1_empty[s]);
      } else {
        // ---- r = conv + bias + x ; LayerNorm over the N channels of this position (shifted single pass, pivot = r[0])
        float pivot = 0.f, sa = 0.f, sq = 0.f;
#pragma unroll
        for (int c0 = 0; c0 < N; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(tacc + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 ad = *reinterpret_cast<const float4*>(&ep_add[c0 + j4 * 4]);
            const float4 x = *reinterpret_cast<const float4*>(rt + (c0 / 32) * 128 * 128 + row * 128 + ((j4 ^ sw) << 4));
            const float r0 = __uint_as_float(v[j4 * 4 + 0]) + ad.x + x.x, r1 = __uint_as_float(v[j4 * 4 + 1]) + ad.y + x.y;
            const float r2 = __uint_as_float(v[j4 * 4 + 2]) + ad.z + x.z, r3 = __uint_as_float(v[j4 * 4 + 3]) + ad.w + x.w;
            if (c0 == 0 && j4 == 0) pivot = r0;
            const float d0 = r0 - pivot, d1 = r1 - pivot, d2 = r2 - pivot, d3 = r3 - pivot;
            sa += (d0 + d1) + (d2 + d3);
            sq = fmaf(d0, d0, sq); sq = fmaf(d1, d1, sq); sq = fmaf(d2, d2, sq); sq = fmaf(d3, d3, sq);
          }
        }
        const float dm = sa * (1.f / N);
        const float mean = pivot + dm;
        const float var = fmaxf(sq * (1.f / N) - dm * dm, 0.f);
        const float rstd = rsqrtf(var + 1e-5f);
        // ---- m = LN(r) * (1 + scale) + shift  -> fp32 in place (inject residual / output), bf16 operand or output copy
#pragma unroll
        for (int c0 = 0; c0 < N; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(tacc + c0, v);
          tmem_ld_wait();
          float m[32];
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 ad = *reinterpret_cast<const float4*>(&ep_add[c0 + j4 * 4]);
            const float4 gg = *reinterpret_cast<const float4*>(&ep_g[c0 + j4 * 4]);
            const float4 sh = *reinterpret_cast<const float4*>(&ep_sh[c0 + j4 * 4]);
            uint8_t* slot = rt + (c0 / 32) * 128 * 128 + row * 128 + ((j4 ^ sw) << 4);
            const float4 x = *reinterpret_cast<const float4*>(slot);
            m[j4 * 4 + 0] = fmaf((__uint_as_float(v[j4 * 4 + 0]) + ad.x + x.x - mean) * rstd, gg.x, sh.x);
            m[j4 * 4 + 1] = fmaf((__uint_as_float(v[j4 * 4 + 1]) + ad.y + x.y - mean) * rstd, gg.y, sh.y);
            m[j4 * 4 + 2] = fmaf((__uint_as_float(v[j4 * 4 + 2]) + ad.z + x.z - mean) * rstd, gg.z, sh.z);
            m[j4 * 4 + 3] = fmaf((__uint_as_float(v[j4 * 4 + 3]) + ad.w + x.w - mean) * rstd, gg.w, sh.w);
            *reinterpret_cast<float4*>(slot) = make_float4(m[j4 * 4], m[j4 * 4 + 1], m[j4 * 4 + 2], m[j4 * 4 + 3]);
          }
          if ((EPI == 2 && !PIPE) || (EPI != 2 && p.has_out_t)) store_bf16_chunk(tt, m, c0);
          if (EPI == 1 && p.stats_out != nullptr && row_valid) add_stats(m, c0);
        }
        tc_fence_before();
        mbar_arrive(&acc1_empty[s]);
        if (PIPE) {
          if (pend) finish_tile(p_b, p_l0, p_rs, p_valid, p_i);      // its chained MMA ran while this tile was normalised
          // chained operand of THIS tile: bf16(m) from the fp32 slot + the onset context -> A2
#pragma unroll
          for (int c0 = 0; c0 < N; c0 += 32) {
            float m[32];
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 mm = *reinterpret_cast<const float4*>(rt + (c0 / 32) * 128 * 128 + row * 128 + ((j4 ^ sw) << 4));
              m[j4 * 4 + 0] = mm.x; m[j4 * 4 + 1] = mm.y; m[j4 * 4 + 2] = mm.z; m[j4 * 4 + 3] = mm.w;
            }
            store_bf16_chunk(sA2, m, c0);
          }
          {
            const uint8_t* cx = sCTX + rs * C::CTX_BYTES + row * (p.ctx_ch * 2);
            for (int j = 0; j < p.ctx_ch / 8; ++j) {
              const int ch8 = (N % 64) / 8 + j;
              *reinterpret_cast<uint4*>(sA2 + (N / 64) * 128 * 128 + row * 128 + ((ch8 ^ sw) << 4)) =
                  *reinterpret_cast<const uint4*>(cx + j * 16);
            }
          }
          fence_proxy_async();
          mbar_arrive(a2_full);
          pend = true; p_b = b; p_l0 = l0; p_rs = rs; p_valid = row_valid; p_i = i;
          continue;                       // stores of this tile happen in finish_tile
        }
        if (EPI == 2) {
          // onset context of this position -> K columns [N, N + ctx) of the chained operand
          {
            const uint8_t* cx = sCTX + rs * C::CTX_BYTES + row * (p.ctx_ch * 2);
            for (int j = 0; j < p.ctx_ch / 8; ++j) {
              const int ch8 = (N % 64) / 8 + j;       // N = 32: chunks 4.. of atom 0 ; N = 64: chunks 0.. of atom 1
              *reinterpret_cast<uint4*>(sA2 + (N / 64) * 128 * 128 + row * 128 + ((ch8 ^ sw) << 4)) =
                  *reinterpret_cast<const uint4*>(cx + j * 16);
            }
          }
          fence_proxy_async();            // A2 was written through the generic proxy, the MMA reads it through the async proxy
          mbar_arrive(a2_full);
          mbar_wait(acc2_full, i & 1);
          tc_fence_after();
          const uint32_t tacc2 = tmem_base + 2 * N + lane_off;
#pragma unroll
          for (int c0 = 0; c0 < N; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tacc2 + c0, v);
            tmem_ld_wait();
            float y[32];
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 ad = *reinterpret_cast<const float4*>(&ep_mul[c0 + j4 * 4]);
              uint8_t* slot = rt + (c0 / 32) * 128 * 128 + row * 128 + ((j4 ^ sw) << 4);
              const float4 mm = *reinterpret_cast<const float4*>(slot);
              y[j4 * 4 + 0] = __uint_as_float(v[j4 * 4 + 0]) + ad.x + mm.x;
              y[j4 * 4 + 1] = __uint_as_float(v[j4 * 4 + 1]) + ad.y + mm.y;
              y[j4 * 4 + 2] = __uint_as_float(v[j4 * 4 + 2]) + ad.z + mm.z;
              y[j4 * 4 + 3] = __uint_as_float(v[j4 * 4 + 3]) + ad.w + mm.w;
              *reinterpret_cast<float4*>(slot) = make_float4(y[j4 * 4], y[j4 * 4 + 1], y[j4 * 4 + 2], y[j4 * 4 + 3]);
            }
            if (p.has_out_t) store_bf16_chunk(tt, y, c0);    // aliases A2: the chained MMA has completed (acc2_full)
            if (p.stats_out != nullptr && row_valid) add_stats(y, c0);
          }
          tc_fence_before();
        }
      }
      // (2) hand the finished tile(s) to the TMA store engine
      fence_proxy_async();
      named_bar(1, 128);
      if (elected) {
        if (p.has_out_r)
          for (int a = 0; a < C::RA; ++a) tma_store_3d(&p.tmR, rt + a * 128 * 128, a * 32, l0, b);
        if (p.has_out_t)
          for (int a = 0; a < C::TA; ++a) tma_store_3d(&p.tmT, tt + a * 128 * 128, a * 64, l0, b);
        bulk_commit();
        if (EPI == 2 && p.has_out_t) {
          bulk_wait_read<0>();                               // A2 (aliased output copy) is rewritten by the very next tile
          mbar_arrive(&r_empty[rs]);                         // ... and this tile's own residual slot has been drained too
        } else if (C::NT == 1) {
          bulk_wait_read<0>();                               // the single staging tile is rewritten by the very next tile
          if (use_r) mbar_arrive(&r_empty[rs]);
        } else {
          bulk_wait_read<1>();                               // tile i-1's stores no longer read their buffers
          if (use_r && i > 0) mbar_arrive(&r_empty[(i - 1) % NSR]);
        }
      }
    }
    if (PIPE && pend) finish_tile(p_b, p_l0, p_rs, p_valid, p_i);
    flush_stats(cur_b);
    if (elected) bulk_wait<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, C::TMEM_COLS);
}


// ------------------------------------------------------------------------------------------------ instantiation table
//        K1   N  TAPS XF EPI BMOD USE_R
#define SFB_RK_LIST(X)                                                                          \
  X(32, 32, 3, 2, 0, 32, 0)   /* R1  d1: GN+SiLU(x fp32) -> conv3 -> h bf16 (+stats)          */ \
  X(64, 64, 3, 2, 0, 64, 0)   /* R1  d2                                                        */ \
  X(32, 32, 3, 1, 2, 32, 1)   /* R2  d1: GN+SiLU(h) -> conv3 +x -> LN/Mod -> inject -> y       */ \
  X(64, 64, 3, 1, 2, 64, 1)   /* R2  d2                                                        */ \
  X(32, 32, 1, 0, 0, 32, 1)   /* Down into d1 (K = 4 x 8)                                      */ \
  X(128, 64, 1, 0, 0, 64, 1)  /* Down into d2 (K = 4 x 32)                                     */ \
  X(32, 32, 3, 0, 0, 8, 1)    /* Up from d1, nearest (folded conv3, N = 4 x 8) + SkipModulate  */ \
  X(32, 32, 1, 0, 0, 8, 1)    /* Up from d1, transpose                                         */

struct RkKey { int K1, N, TAPS, XF, EPI, BMOD, USE_R; };

inline int rk_find(int K1, int N, int TAPS, int XF, int EPI, int BMOD, int USE_R) {
  const RkKey keys[] = {
#define X(a, b, c, d, e, f, g) {a, b, c, d, e, f, g},
      SFB_RK_LIST(X)
#undef X
  };
  for (int i = 0; i < (int)(sizeof(keys) / sizeof(keys[0])); ++i) {
    const RkKey& k = keys[i];
    if (k.K1 == K1 && k.N == N && k.TAPS == TAPS && k.XF == XF && k.EPI == EPI && k.BMOD == BMOD && k.USE_R == USE_R) return i;
  }
  return -1;
}
inline RkKey rk_key(int id) {
  const RkKey keys[] = {
#define X(a, b, c, d, e, f, g) {a, b, c, d, e, f, g},
      SFB_RK_LIST(X)
#undef X
  };
  return keys[id];
}
inline int rk_ctx_capacity(int N) { (void)N; return 16; }   // context channels the chained operand / staging tile can take

inline cudaError_t rk_set_attrs() {
  cudaError_t e = cudaSuccess;
#define X(a, b, c, d, e_, f, g)                                                                                          \
  if (e == cudaSuccess)                                                                                                  \
    e = cudaFuncSetAttribute(rk_kernel<a, b, c, d, e_, f, g>, cudaFuncAttributeMaxDynamicSharedMemorySize,               \
                             RkCfg<a, b, c, d, e_, f, g>::SMEM);
  SFB_RK_LIST(X)
#undef X
  return e;
}
inline void rk_launch(int id, const RkParams& p, int num_sms, cudaStream_t st) {
  int i = 0;
#define X(a, b, c, d, e_, f, g)                                                                                          \
  if (id == i++) {                                                                                                       \
    const int cap = num_sms * RkCfg<a, b, c, d, e_, f, g>::OCC;                                                          \
    const int grid = p.total_tiles < cap ? p.total_tiles : cap;                                                          \
    launch_pdl(rk_kernel<a, b, c, d, e_, f, g>, grid, RkCfg<a, b, c, d, e_, f, g>::kThreads, RkCfg<a, b, c, d, e_, f, g>::SMEM, st, p); \
    return;                                                                                                              \
  }
  SFB_RK_LIST(X)
#undef X
}

}  // namespace sfb
