// Streaming-K fused tcgen05 implicit-GEMM for the tensor-bound depths (C >= 128), bf16 operands.
//
//   D[b, l, n] = sum_{tap, k} f(A1)[b, l + tap - pad, k] W[tap N + n, k]  +  sum_k A2[b % m, l, k] W[n, K1 + k]
//   out        = colscale[n'] (D + bias[n']) + rowvec[b, n'] + g(resid)                    n' = n % bias_mod
//
// f = none | GroupNorm(8)+SiLU (ResNet convs, a6) | per-position LayerNorm (x Modulation) (inject a7/a8, QKV pre-norm a9)
// g = none | identity | LayerNorm/Modulation recomputed in fp32 (InjectChannels adds the MODULATED tensor, which is
// never stored).  So the GroupNorm / Modulation / pre-norm passes of the reference do not exist as kernels: the
// statistics they need (fp64 group sums per clip, fp32 row sums per position) are emitted by the PRODUCER's epilogue.
//
// Persistent, warp-specialised, 128 x BN output tiles (n fastest), fp32 accumulators double-buffered in TMEM:
//   warp 0      TMA producer (A): ring of [136 x 64] bf16 tiles - ONE load per K chunk serves all three conv taps
//               (row-shifted UMMA descriptors, +128 B per tap)
//   warp 3      TMA producer (B): ring of [BN x 64] weight tiles, its own thread so that a full weight ring never holds
//               back the A loads the transform warps are waiting for; the first ring fill of static weights is issued
//               before griddepcontrol.wait (it overlaps the previous kernel's tail)
//   warp 1      tcgen05.mma issuer
//   warp 2      TMA producer (epilogue): residual chunks [128 x 32] fp32 into the R ring
//   warps 4-7   A transform in place on the landed tile (fence.proxy.async before the MMA sees it)
//   warps 8-11  epilogue, one accumulator row per thread, 32-column chunks: TMEM -> regs -> math -> fp32 chunk written
//               IN PLACE over the residual chunk + bf16 chunk -> TMA stores; output statistics on the fly.
#pragma once
#include "rk_tc.cuh"
#undef SFB_FILE_ID
#define SFB_FILE_ID 4   // sk_tc.cuh

namespace sfb {

struct SkParams {
  CUtensorMap tmA1;   // bf16 [K1, L, B]    box [64, 136 | 128, 1]
  CUtensorMap tmA2;   // bf16 [K2, L, B2]   box [64, 128, 1]
  CUtensorMap tmW;    // bf16 [K1 + K2, taps * N, copies] box [64, BN, 1]; copy = b % w_bmod (per-step / per-clip scaled weights)
  CUtensorMap tmR;    // fp32 [N, L, B]     box [32, 128, 1], 128-byte swizzle  residual in (one load per 32-column chunk)
  CUtensorMap tmRs;   // fp32 [N, L, B]     box [32, 32, 1],  128-byte swizzle  fp32 out (one store per epilogue warp and chunk)
  CUtensorMap tmT;    // bf16 [N, L, B]     box [32, 32, 1],  64-byte swizzle   bf16 out (one store per epilogue warp and chunk)
  // A transform
  int xf;                       // 0 none, 1 GroupNorm + SiLU, 2 LayerNorm (x Modulation when mod != null),
                                // 3 A2 rows / rstd (the context half of a LayerNorm-folded inject, see ln_fold)
  const double* stats_in;       // [B, 8, 2] group sums of A1 (xf == 1)
  const float *gamma, *beta;    // [K1]
  const float* rowstats_in;     // [B * L, rs_parts, 2] partial (sum, sum of squares) of every A1 / residual row
  int rs_parts;
  const float* mod;             // [2 * K1] Modulation scale | shift of this step, row (b % mod_bmod) * mod_bstride
  int mod_bstride, mod_bmod;
  // epilogue
  const float* bias;            // [bias_mod] or null
  const float* colscale;        // [bias_mod] or null, row (b % cs_bmod) * cs_bstride
  const float* rowvec;          // [B, rowvec_stride] or null
  int bias_mod, cs_bstride, cs_bmod, rowvec_stride;
  int resid_mode;               // 0 none, 1 + resid, 2 + LayerNorm/Modulation(resid)
  // LayerNorm folded out of the A operand: W LN(x) = rstd (W x - mean W 1), so the MMA runs on the RAW bf16 rows and
  // the accumulator is fixed per row in the epilogue: D' = rstd_l (D - mean_l ws[n]).  A Modulation scale lives in
  // the weights (W diag(1 + s), rebuilt per step by inject_fold_kernel), its shift in addvec = W sh.
  int ln_fold;
  int w_bmod, ws_bstride;       // weight copy / ws / addvec row = (b % w_bmod) (* ws_bstride)
  const float* ws;              // [copies][ws_bstride] sum_k W[n, k < K1]  (of the bf16 weights the MMA reads)
  const float* addvec;          // [copies][ws_bstride] extra additive column vector or null
  int has_out_r, has_out_t;
  double* stats_out;            // [B, 8, 2] group sums of the output (group = (n / GS) % 8) or null
  float* rowstats_out;          // [B * L, n_tiles, 2] or null
  int L, tiles_per_clip, N, n_tiles, total_tiles, taps, k1_chunks, k2_chunks, K1, a2_bmod;
  float eps;
  // shared-memory ring depths of this op (host: sk_pick_rings) and the epilogue width
  int na, nb, nr;               // A / weight / residual stages
  int epi12;                    // xf == 0 ops: the four idle A-transform warps join the epilogue (12 warps, three chunk groups)
  int w_static;                 // the weight tensor is never written inside the launch chain: its first tiles are loaded BEFORE griddepcontrol.wait
  int tag;                      // plan op index (wait log)
  long long* dbg;               // optional timeline buffer (CTA 0 only): [role][256] clock64 stamps (tools/sk_timeline.py)
};

#define SK_STAMP(role, idx) do { if (p.dbg != nullptr && blockIdx.x == 0 && (idx) < 256) p.dbg[(role) * 256 + (idx)] = clock64(); } while (0)

template <int BN> struct SkCfg {
  static constexpr int A_BYTES = 136 * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int R_BYTES = 128 * 128;    // [128 rows][32 fp32], 128B swizzle
  static constexpr int T_WARP = 2048;          // per epilogue warp: [32 rows][32 bf16], 64B swizzle
  static constexpr int KMAX = 1024;            // largest transformed K1
  static constexpr int MAX_NA = 4, MAX_NB = 6, MAX_NR = 6;
  // fixed part behind the rings: bf16 staging (8 or 12 warps) | coefficient tables | row table | epilogue vectors | barriers
  static constexpr int TAB_BYTES = 2 * KMAX * 4, ROWTAB_BYTES = 2 * 136 * 4 + 64, VEC_BYTES = 5 * BN * 4, BAR_BYTES = 512;
  __host__ __device__ static constexpr int fixed_bytes(int epi12) { return (epi12 ? 12 : 8) * T_WARP + TAB_BYTES + ROWTAB_BYTES + VEC_BYTES + BAR_BYTES; }
  __host__ __device__ static constexpr int smem_bytes(int na, int nb, int nr, int epi12) {
    return na * A_BYTES + nb * B_BYTES + nr * R_BYTES + fixed_bytes(epi12);
  }
  static constexpr int kThreads = 512;
  static constexpr int kMaxSmem = 232448;
};
// Ring depths per op class under the 227 KB budget (host side).  The mainloop of the k = 3 convs wants A stages (load ->
// in-place transform -> MMA each hold one), the HBM-bound 1 x 1 ops want residual chunks in flight.
template <int BN> inline bool fits_static(int a, int b2, int r, int epi12) { return SkCfg<BN>::smem_bytes(a, b2, r, epi12) <= SkCfg<BN>::kMaxSmem; }
template <int BN>
inline void sk_pick_rings(int taps, int xf, bool uses_r, bool has_resid, int epi12, int& na, int& nb, int& nr) {
  using C = SkCfg<BN>;
  const int npar = epi12 ? 3 : 2;           // epilogue chunk groups
  if (taps == 3) { na = uses_r ? 3 : 4; nr = uses_r ? 3 : 0; }
  else if (epi12) { na = uses_r ? 2 : 3; nr = uses_r ? 5 : 0; }
  else { na = 2; nr = uses_r ? 4 : 0; }
  if (uses_r && !has_resid) nr = npar;      // plain output staging: exactly one private slot per chunk group
  if (const char* e = getenv(taps == 3 ? "SFB_SK_RINGS3" : (has_resid ? "SFB_SK_RINGS1R" : "SFB_SK_RINGS1"))) {   // tuning aid: "na,nb,nr"
    int a = 0, b2 = 0, r = 0;
    if (sscanf(e, "%d,%d,%d", &a, &b2, &r) == 3 && a >= 2 && b2 >= 2 && a <= C::MAX_NA && b2 <= C::MAX_NB && r <= C::MAX_NR &&
        r >= (uses_r ? (has_resid ? 2 : npar) : 0) && fits_static<BN>(a, b2, uses_r ? r : 0, epi12)) { na = a; nb = b2; nr = uses_r ? r : 0; return; }
  }
  const int nr_min = uses_r ? (has_resid ? 2 : npar) : 0;
  auto fits = [&](int a, int b2, int r) { return C::smem_bytes(a, b2, r, epi12) <= C::kMaxSmem; };
  nb = taps == 3 ? C::MAX_NB : na + 1;      // one weight tile per tap: a 1 x 1 op never runs further ahead on B than on A
  while (nb > 2 && !fits(na, nb, nr)) --nb;
  while (!fits(na, nb, nr) && nr > nr_min) --nr;
  while (!fits(na, nb, nr) && na > 2) --na;
}

// EPI: compile-time epilogue shape (the per-column math sits in the hottest loop of the small-K ops, where uniform
// run-time branches cost a third of the issue slots and pin every shared-memory load behind a branch):
//   0 plain   1 + resid   2 LayerNorm fold   3 LayerNorm fold + Modulation(resid)   4 colscale + resid   5 run-time flags
__host__ __device__ constexpr int sk_epi_of(int ln_fold, int resid_mode, int has_colscale) {
  return (!ln_fold && resid_mode == 0 && !has_colscale) ? 0 : (!ln_fold && resid_mode == 1 && !has_colscale) ? 1
       : (ln_fold && resid_mode == 0 && !has_colscale) ? 2 : (ln_fold && resid_mode == 2 && !has_colscale) ? 3
       : (!ln_fold && resid_mode == 1 && has_colscale) ? 4 : 5;
}
template <int BN, int GS, int EPI>
__global__ void __launch_bounds__(512, 1) sk_kernel(const __grid_constant__ SkParams p) {
  pdl_trigger();
  using C = SkCfg<BN>;
  const int NA = p.na, NB = p.nb, NR = p.nr;          // ring depths of this op (uniform)
  const bool epi12 = p.epi12 != 0;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = sA + NA * C::A_BYTES;
  uint8_t* sR = sB + NB * C::B_BYTES;
  uint8_t* sT = sR + NR * C::R_BYTES;
  float* tab_a = reinterpret_cast<float*>(sT + (epi12 ? 12 : 8) * C::T_WARP);
  float* tab_b = tab_a + C::KMAX;
  float* rowtab = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(tab_a) + C::TAB_BYTES);
  float* sVEC = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(rowtab) + C::ROWTAB_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sVEC) + C::VEC_BYTES);
  uint64_t* a_full = bars;                       // [NA]
  uint64_t* a_empty = a_full + C::MAX_NA;        // [NA]
  uint64_t* op_full = a_empty + C::MAX_NA;       // [NA]
  uint64_t* b_full = op_full + C::MAX_NA;        // [NB]
  uint64_t* b_empty = b_full + C::MAX_NB;        // [NB]
  uint64_t* acc_full = b_empty + C::MAX_NB;      // [2]
  uint64_t* acc_empty = acc_full + 2;            // [2]
  uint64_t* rc_full = acc_empty + 2;             // [NR]
  uint64_t* rc_empty = rc_full + C::MAX_NR;      // [NR]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rc_empty + C::MAX_NR);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t_begin = (int)(((long long)p.total_tiles * blockIdx.x) / gridDim.x);
  const int t_end = (int)(((long long)p.total_tiles * (blockIdx.x + 1)) / gridDim.x);
  const int pad = p.taps == 3 ? 1 : 0;
  const int rows_a = p.taps == 3 ? 136 : 128;
  constexpr int NCH = BN / 32;             // epilogue chunks per tile

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmA1);
    tma_prefetch_desc(&p.tmW);
    for (int s = 0; s < NA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); mbar_init(&op_full[s], 128); }
    for (int s = 0; s < NB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], epi12 ? 384 : 256); }
    for (int s = 0; s < NR; ++s) { mbar_init(&rc_full[s], 1); mbar_init(&rc_empty[s], epi12 ? 12 : 8); }   // every epilogue warp, once per fill
    fence_barrier_init();
  }
  if (warp == 3) { tmem_alloc(tmem_slot, 2 * BN); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  // Static weights do not depend on the previous kernel: the first fill of the weight ring (tile t_begin, first k1
  // chunks) and the GroupNorm affine vectors are requested now, under the previous kernel's tail.
  int n_pre = 0;
  if (p.w_static && t_begin < t_end) {
    const int first_b = p.k1_chunks * p.taps;            // weight tiles of the K1 part of one output tile
    n_pre = NB < first_b ? NB : first_b;
    if (warp == 3 && lane == 0) {
      const int n0 = (t_begin % p.n_tiles) * BN;
      for (int j = 0; j < n_pre; ++j) {
        mbar_expect_tx(&b_full[j], C::B_BYTES);
        tma_load_3d(sB + j * C::B_BYTES, &p.tmW, &b_full[j], (j / p.taps) * 64, (j % p.taps) * p.N + n0, 0);
      }
    }
    if (p.xf == 1 && warp >= 4 && warp < 8)
      for (int c = (threadIdx.x - 128) * 32; c < p.K1; c += 128 * 32) { prefetch_l1(p.gamma + c); prefetch_l1(p.beta + c); }
  }
  pdl_wait();            // everything above is independent of the previous kernel's output
  mark_progress(p.tag);
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) SK_STAMP(7, 0);

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------- mainloop TMA producer: A tiles
      uint32_t ia = 0;
      int sa = 0;                         // ring slots and phases advance incrementally (the depths are run-time values:
      uint32_t pha = 0;                   // a division per k chunk on this single thread would cost more than the TMA issue)
      for (int t = t_begin; t < t_end; ++t) {
        const int mi = t / p.n_tiles;
        const int b = mi / p.tiles_per_clip;
        const int l0 = (mi % p.tiles_per_clip) * 128;
        for (int kc = 0; kc < p.k1_chunks + p.k2_chunks; ++kc, ++ia) {
          const bool second = kc >= p.k1_chunks;
          mbar_wait(&a_empty[sa], pha ^ 1);
          if (!second) {
            mbar_expect_tx(&a_full[sa], rows_a * 128);
            tma_load_3d(sA + sa * C::A_BYTES, &p.tmA1, &a_full[sa], kc * 64, l0 - pad, b);
          } else {
            mbar_expect_tx(&a_full[sa], 128 * 128);
            tma_load_3d(sA + sa * C::A_BYTES, &p.tmA2, &a_full[sa], (kc - p.k1_chunks) * 64, l0, b % p.a2_bmod);
          }
          SK_STAMP(0, ia);
          if (++sa == NA) { sa = 0; pha ^= 1; }
        }
      }
    }
  } else if (warp == 3) {
    if (lane == 0) {
      // ------------------------------------------------------------- mainloop TMA producer: weight tiles
      uint32_t ib = 0;
      int sb = 0;
      uint32_t phb = 0;
      for (int t = t_begin; t < t_end; ++t) {
        const int n0 = (t % p.n_tiles) * BN;
        const int b = (t / p.n_tiles) / p.tiles_per_clip;
        const int wcopy = b % p.w_bmod;
        for (int kc = 0; kc < p.k1_chunks + p.k2_chunks; ++kc) {
          const bool second = kc >= p.k1_chunks;
          const int ntap = second ? 1 : p.taps;
          for (int tap = 0; tap < ntap; ++tap, ++ib) {
            if (ib >= (uint32_t)n_pre) {             // the first n_pre tiles were issued before griddepcontrol.wait
              mbar_wait(&b_empty[sb], phb ^ 1);
              mbar_expect_tx(&b_full[sb], C::B_BYTES);
              tma_load_3d(sB + sb * C::B_BYTES, &p.tmW, &b_full[sb], second ? p.K1 + (kc - p.k1_chunks) * 64 : kc * 64, tap * p.N + n0, wcopy);
              SK_STAMP(1, ib);
            }
            if (++sb == NB) { sb = 0; phb ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------------------------------------- MMA issuer
      constexpr uint32_t idesc = make_idesc(1 /*bf16*/, 128, BN, 0, 0);
      uint32_t ia = 0, ib = 0, i = 0;
      int sa = 0, sb = 0;
      uint32_t pha = 0, phb = 0;
      for (int t = t_begin; t < t_end; ++t, ++i) {
        const uint32_t acc = i & 1;
        mbar_wait(&acc_empty[acc], ((i >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tacc = tmem_base + acc * BN;
        bool first = true;
        for (int kc = 0; kc < p.k1_chunks + p.k2_chunks; ++kc, ++ia) {
          const bool second = kc >= p.k1_chunks;
          mbar_wait(&a_full[sa], pha);
          if (p.xf) mbar_wait(&op_full[sa], pha);
          tc_fence_after();
          const uint32_t abase = smem_u32(sA + sa * C::A_BYTES);
          SK_STAMP(3, ia);
          const int ntap = second ? 1 : p.taps;
          for (int tap = 0; tap < ntap; ++tap, ++ib) {
            mbar_wait(&b_full[sb], phb);
            tc_fence_after();
            const uint32_t bbase = smem_u32(sB + sb * C::B_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_ss<false>(tacc, make_smem_desc_sw128(abase + tap * 128 + k * 32, 16, 1024),
                             make_smem_desc_sw128(bbase + k * 32, 16, 1024), idesc, first ? 0u : 1u);
              first = false;
            }
            umma_commit(&b_empty[sb]);
            if (++sb == NB) { sb = 0; phb ^= 1; }
          }
          umma_commit(&a_empty[sa]);
          if (++sa == NA) { sa = 0; pha ^= 1; }
        }
        umma_commit(&acc_full[acc]);
      }
    }
  } else if (warp == 2) {
    if (lane == 0 && p.resid_mode != 0) {
      // ------------------------------------------------------------- epilogue TMA producer: residual chunks
      int rs = 0;
      uint32_t phr = 0;
      for (int t = t_begin; t < t_end; ++t) {
        const int n0 = (t % p.n_tiles) * BN;
        const int mi = t / p.n_tiles;
        const int b = mi / p.tiles_per_clip;
        const int l0 = (mi % p.tiles_per_clip) * 128;
        for (int c = 0; c < NCH; ++c) {
          mbar_wait(&rc_empty[rs], phr ^ 1);
          mbar_expect_tx(&rc_full[rs], C::R_BYTES);
          tma_load_3d(sR + rs * C::R_BYTES, &p.tmR, &rc_full[rs], n0 + c * 32, l0, b);
          if (++rs == NR) { rs = 0; phr ^= 1; }
        }
      }
    }
  } else if (warp >= 4 && warp < 8 && !epi12) {
    if (p.xf) {
      // ------------------------------------------------------------- A transform (in place on the landed bf16 tile)
      const int tid = threadIdx.x - 128;        // 0..127
      const int g = tid & 7;                     // 16-byte chunk (8 channels) of the 128-byte row
      const int rsub = tid >> 3;                 // 0..15
      int cur_b = -1, cur_mi = -1;
      uint32_t ia = 0;
      int sa = 0;
      uint32_t pha = 0;
      for (int t = t_begin; t < t_end; ++t) {
        const int mi = t / p.n_tiles;
        const int b = mi / p.tiles_per_clip;
        const int l0 = (mi % p.tiles_per_clip) * 128;
        if (b != cur_b) {         // per-clip channel coefficients: y = x * a[c] + b[c]
          cur_b = b;
          named_bar(3, 128);
          if (p.xf == 1) {
            // eight threads turn the fp64 group sums into (mean, rstd) - ONE global-load latency - then every thread
            // writes the packed bf16x2 coefficients of its channel pairs straight into the tables
            const int gsa = p.K1 / 8;
            if (tid < 8) {
              const double inv_cnt = 1.0 / ((double)p.L * gsa);
              const double s1 = p.stats_in[(size_t)b * 16 + tid * 2], s2 = p.stats_in[(size_t)b * 16 + tid * 2 + 1];
              const double mean = s1 * inv_cnt;
              const double var = fma(-mean, mean, s2 * inv_cnt);
              rowtab[2 * tid] = (float)mean;
              rowtab[2 * tid + 1] = rsqrtf(fmaxf((float)var, 0.f) + p.eps);
            }
            named_bar(3, 128);
            uint32_t* tabp = reinterpret_cast<uint32_t*>(tab_a);
#pragma unroll 4
            for (int c2 = tid; c2 < p.K1 / 2; c2 += 128) {
              const int grp = (2 * c2) / gsa;                  // gsa is even: both channels of a pair share the group
              const float mean = rowtab[2 * grp], rstd = rowtab[2 * grp + 1];
              const float2 gm = __ldg(reinterpret_cast<const float2*>(p.gamma) + c2);
              const float2 bt = __ldg(reinterpret_cast<const float2*>(p.beta) + c2);
              const float a0 = rstd * gm.x, a1 = rstd * gm.y;
              uint32_t k0, k1, k2;
              gn_pack_coef(a0, bt.x - mean * a0, a1, bt.y - mean * a1, k0, k1, k2);
              tabp[c2] = k0; tabp[C::KMAX / 2 + c2] = k1; tabp[C::KMAX + c2] = k2;
            }
          } else if (p.xf == 2) {
            const float* md = p.mod ? p.mod + (size_t)(b % p.mod_bmod) * p.mod_bstride : nullptr;
            for (int c = tid; c < p.K1; c += 128) {
              tab_a[c] = md ? 1.f + md[c] : 1.f;
              tab_b[c] = md ? md[p.K1 + c] : 0.f;
            }
          }
          named_bar(3, 128);
        }
        if (p.xf >= 2 && mi != cur_mi) {   // per-position LayerNorm statistics of this M tile's rows
          cur_mi = mi;
          named_bar(3, 128);
          {
            const int l = l0 + tid;
            float mean = 0.f, rstd = 0.f;
            if (l < p.L) {
              const float* rsrc = p.rowstats_in + ((size_t)b * p.L + l) * p.rs_parts * 2;
              float s1 = 0.f, s2 = 0.f;
              for (int j = 0; j < p.rs_parts; ++j) { s1 += rsrc[2 * j]; s2 += rsrc[2 * j + 1]; }
              mean = s1 / (float)p.K1;
              const float var = fmaxf(s2 / (float)p.K1 - mean * mean, 0.f) + p.eps;
              rstd = p.xf == 3 ? sqrtf(var) : rsqrtf(var);       // xf 3 scales the context rows by 1 / rstd
            }
            rowtab[2 * tid] = mean;
            rowtab[2 * tid + 1] = rstd;
          }
          named_bar(3, 128);
        }
        for (int kc = 0; kc < p.k1_chunks + p.k2_chunks; ++kc, ++ia) {
          mbar_wait(&a_full[sa], pha);
          if (tid == 0) SK_STAMP(2, 2 * ia);
          if (p.xf == 3 ? kc >= p.k1_chunks : kc < p.k1_chunks) {
            uint8_t* tile = sA + sa * C::A_BYTES;
            float ca[8], cb[8];
            uint32_t pa[4], pbh[4], pbl[4];
            if (p.xf == 1) {
              const uint32_t* tabp = reinterpret_cast<const uint32_t*>(tab_a);
              const uint4 q0 = *reinterpret_cast<const uint4*>(&tabp[kc * 32 + g * 4]);
              const uint4 q1 = *reinterpret_cast<const uint4*>(&tabp[C::KMAX / 2 + kc * 32 + g * 4]);
              const uint4 q2 = *reinterpret_cast<const uint4*>(&tabp[C::KMAX + kc * 32 + g * 4]);
              pa[0] = q0.x; pa[1] = q0.y; pa[2] = q0.z; pa[3] = q0.w;
              pbh[0] = q1.x; pbh[1] = q1.y; pbh[2] = q1.z; pbh[3] = q1.w;
              pbl[0] = q2.x; pbl[1] = q2.y; pbl[2] = q2.z; pbl[3] = q2.w;
            } else if (p.xf == 2) {
              const float4 a0 = *reinterpret_cast<const float4*>(&tab_a[kc * 64 + g * 8]), a1 = *reinterpret_cast<const float4*>(&tab_a[kc * 64 + g * 8 + 4]);
              const float4 b0 = *reinterpret_cast<const float4*>(&tab_b[kc * 64 + g * 8]), b1 = *reinterpret_cast<const float4*>(&tab_b[kc * 64 + g * 8 + 4]);
              ca[0] = a0.x; ca[1] = a0.y; ca[2] = a0.z; ca[3] = a0.w; ca[4] = a1.x; ca[5] = a1.y; ca[6] = a1.z; ca[7] = a1.w;
              cb[0] = b0.x; cb[1] = b0.y; cb[2] = b0.z; cb[3] = b0.w; cb[4] = b1.x; cb[5] = b1.y; cb[6] = b1.z; cb[7] = b1.w;
            }
            uint4 u[9];
#pragma unroll
            for (int it = 0; it < 9; ++it) {
              const int r = rsub + 16 * it;
              if (r < rows_a) u[it] = *reinterpret_cast<const uint4*>(tile + r * 128 + ((g ^ (r & 7)) << 4));
            }
            if (p.xf == 1) {
              // GroupNorm apply + SiLU, packed bf16x2 (see silu_gn_bf16x2), in three PHASES over all nine rows: the warp
              // issues in order, so a row-by-row chain stalls on every MUFU.TANH -> PRMT -> HFMA2 dependency (measured:
              // ~2100 cycles per tile, 4x the MUFU floor); phased, the 72 MUFU ops of a thread issue back to back.
              // (volatile asm keeps the compiler from re-fusing the phases.)
              uint32_t hh[9][4], tt[9][4];
#pragma unroll
              for (int it = 0; it < 9; ++it) {
                const uint32_t w[4] = {u[it].x, u[it].y, u[it].z, u[it].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  uint32_t h1;
                  asm volatile("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(h1) : "r"(w[j]), "r"(pa[j]), "r"(pbh[j]));
                  asm volatile("add.rn.bf16x2 %0, %1, %2;" : "=r"(hh[it][j]) : "r"(h1), "r"(pbl[j]));
                }
              }
#pragma unroll
              for (int it = 0; it < 9; ++it)
#pragma unroll
                for (int j = 0; j < 4; ++j) asm volatile("tanh.approx.bf16x2 %0, %1;" : "=r"(tt[it][j]) : "r"(hh[it][j]));
#pragma unroll
              for (int it = 0; it < 9; ++it) {
                const int r = rsub + 16 * it;
                const int l = l0 - pad + r;
                const bool ok = l >= 0 && l < p.L;      // conv zero padding applies AFTER the activation
                uint32_t o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  asm volatile("fma.rn.bf16x2 %0, %1, %2, %1;" : "=r"(o[j]) : "r"(hh[it][j]), "r"(tt[it][j]));
                  o[j] = ok ? o[j] : 0u;
                }
                u[it] = make_uint4(o[0], o[1], o[2], o[3]);
              }
            } else {
#pragma unroll
            for (int it = 0; it < 9; ++it) {
              const int r = rsub + 16 * it;
              const int l = l0 - pad + r;
              const bool ok = l >= 0 && l < p.L;      // conv zero padding applies AFTER the activation
              const uint32_t w[4] = {u[it].x, u[it].y, u[it].z, u[it].w};
              float y[8];
              float x[8];
#pragma unroll
              for (int j = 0; j < 4; ++j) { x[2 * j] = __uint_as_float(w[j] << 16); x[2 * j + 1] = __uint_as_float(w[j] & 0xFFFF0000u); }
              if (p.xf == 2) {      // LayerNorm (x Modulation)
                const float mean = rowtab[2 * (r & 127)], rstd = rowtab[2 * (r & 127) + 1];
#pragma unroll
                for (int j = 0; j < 8; ++j) y[j] = fmaf((x[j] - mean) * rstd, ca[j], cb[j]);
              } else {              // context rows / rstd (ln_fold multiplies the whole accumulator row by rstd)
                const float isd = rowtab[2 * (r & 127) + 1];
#pragma unroll
                for (int j = 0; j < 8; ++j) y[j] = x[j] * isd;
              }
#pragma unroll
              for (int j = 0; j < 8; ++j) y[j] = ok ? y[j] : 0.f;
              u[it] = make_uint4(pack_bf16(y[0], y[1]), pack_bf16(y[2], y[3]), pack_bf16(y[4], y[5]), pack_bf16(y[6], y[7]));
            }
            }
#pragma unroll
            for (int it = 0; it < 9; ++it) {
              const int r = rsub + 16 * it;
              if (r < rows_a) *reinterpret_cast<uint4*>(tile + r * 128 + ((g ^ (r & 7)) << 4)) = u[it];
            }
            fence_proxy_async();
          }
          mbar_arrive(&op_full[sa]);
          if (tid == 0) SK_STAMP(2, 2 * ia + 1);
          if (++sa == NA) { sa = 0; pha ^= 1; }
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------- epilogue: 8 warps (12 with the idle transform warps of
    // an xf == 0 op), one accumulator row per thread.
    // The two warps of a TMEM lane quarter (same SM sub-partition) take ALTERNATE 32-column chunks and own them end to
    // end: TMEM -> registers -> math -> private slice of the staging buffers -> own TMA stores (4 KB fp32 + 2 KB bf16
    // per store).  No CTA-level barrier in the chunk loop; the sibling warp hides TMEM / shared-memory / bulk-wait
    // latencies.  32-column chunks halve the per-column cost of the fixed per-chunk work (bulk wait, proxy fence,
    // store issue, barrier traffic) against 16-column ones and give every thread 32 independent values to work on.
    const int q4 = warp & 3;
    const int par = warp >= 8 ? (warp - 8) >> 2 : 2;        // chunk group: warps 8-11, 12-15, (4-7)
    const int npar = epi12 ? 3 : 2;
    const int nthr_epi = epi12 ? 384 : 256;
    const int row = q4 * 32 + lane;
    const int et = warp >= 8 ? threadIdx.x - 256 : threadIdx.x + 128;     // 0..255, helpers 256..383
    const uint32_t lane_off = uint32_t(q4 * 32) << 16;
    const int resid_mode = EPI == 5 ? p.resid_mode : (EPI == 1 || EPI == 4) ? 1 : EPI == 3 ? 2 : 0;
    const bool ln_fold = EPI == 5 ? p.ln_fold != 0 : (EPI == 2 || EPI == 3);
    const bool use_mul = EPI == 5 || EPI == 4;
    const bool has_out_r = p.has_out_r != 0, has_out_t = p.has_out_t != 0;
    float* ep_mul = sVEC;             // [BN] colscale
    float* ep_add = sVEC + BN;        // [BN] bias * colscale + rowvec
    float* ep_g = sVEC + 2 * BN;      // [BN] 1 + modulation scale   (resid_mode 2)
    float* ep_sh = sVEC + 3 * BN;     // [BN] modulation shift
    float* ep_ws = sVEC + 4 * BN;     // [BN] weight column sums (ln_fold)
    uint8_t* tt = sT + (par * 4 + q4) * C::T_WARP; // this warp's bf16 staging: [32 rows][64 B], 64B swizzle
    float s1[8], s2[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { s1[k] = 0.f; s2[k] = 0.f; }
    int cur_b = -1, cur_n = -1, cur_goff = 0;
    auto flush_stats = [&](int b, int goff) {
      if (p.stats_out == nullptr || b < 0) return;
      constexpr int NG = (BN / GS) < 8 ? (BN / GS) : 8;     // distinct local groups inside a tile (local index (c / GS) & 7)
#pragma unroll
      for (int k = 0; k < NG; ++k) {
        float a = s1[k], c = s2[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); c += __shfl_xor_sync(0xffffffffu, c, o); }
        if (lane == 0) {
          const int grp = (k + goff) & 7;
          atomicAdd(&p.stats_out[(size_t)b * 16 + grp * 2], (double)a);
          atomicAdd(&p.stats_out[(size_t)b * 16 + grp * 2 + 1], (double)c);
        }
        s1[k] = 0.f; s2[k] = 0.f;
      }
    };
    uint32_t i = 0;
    int prev_slot = -1;                 // R slot of this warp's previous chunk (handed back once its store was read)
    int rs = 0;                         // R slot / phase of chunk q_cur
    uint32_t phr = 0;
    uint32_t q_cur = 0;                 // next global chunk index this warp has not yet passed (its own or another group's)
    // Residual ring protocol.  The chunk groups take ALTERNATE chunks of one shared ring, so a warp works on only every
    // npar-th fill of a slot - but an mbarrier wait knows a phase by its PARITY only: a warp that skipped a fill cannot tell
    // "my fill landed" from "the fill two phases earlier landed" and would run ahead of the producer on stale data, after
    // which the arrival accounting of rc_empty is off and the ring deadlocks (the r1 hang: seen as all epilogue warps
    // parked on rc_full while the producer waits for a release that was counted one phase early).  Hence every warp
    // PASSES EVERY chunk in order: for a chunk of another group lane 0 waits for its fill and hands the slot straight
    // back (it never touches the data), so each warp observes every phase of every slot and rc_empty counts every
    // epilogue warp once per fill (init count = number of epilogue warps).
    auto pass_foreign_chunks = [&](uint32_t upto) {
      for (; q_cur < upto; ++q_cur) {
        if (resid_mode != 0 && lane == 0) {
          mbar_wait(&rc_full[rs], phr);
          mbar_arrive(&rc_empty[rs]);
        }
        if (++rs == NR) { rs = 0; phr ^= 1; }
      }
    };
    for (int t = t_begin; t < t_end; ++t, ++i) {
      const int n_idx = t % p.n_tiles;
      const int n0 = n_idx * BN;
      const int mi = t / p.n_tiles;
      const int b = mi / p.tiles_per_clip;
      const int l0 = (mi % p.tiles_per_clip) * 128;
      const bool row_valid = l0 + row < p.L;
      const uint32_t acc = i & 1;
      if (t + 1 < t_end) {       // warm L1 with the NEXT tile's per-column vectors and per-row statistics (global latency off the path)
        const int n_idx2 = (t + 1) % p.n_tiles, mi2 = (t + 1) / p.n_tiles;
        const int b2 = mi2 / p.tiles_per_clip, l02 = (mi2 % p.tiles_per_clip) * 128, n02 = n_idx2 * BN;
        if ((resid_mode == 2 || ln_fold) && l02 + row < p.L) prefetch_l1(p.rowstats_in + ((size_t)b2 * p.L + l02 + row) * p.rs_parts * 2);
        if (et < BN && (b2 != b || n_idx2 != n_idx)) {
          const int nm2 = (n02 + et) % p.bias_mod;
          const size_t wrow2 = (size_t)(b2 % p.w_bmod) * p.ws_bstride + n02 + et;
          if (p.bias) prefetch_l1(&p.bias[nm2]);
          if (use_mul && p.colscale) prefetch_l1(&p.colscale[(size_t)(b2 % p.cs_bmod) * p.cs_bstride + nm2]);
          if (p.rowvec) prefetch_l1(&p.rowvec[(size_t)b2 * p.rowvec_stride + nm2]);
          if (p.addvec) prefetch_l1(&p.addvec[wrow2]);
          if (ln_fold) prefetch_l1(&p.ws[wrow2]);
          if (resid_mode == 2) {
            const float* md2 = p.mod + (size_t)(b2 % p.mod_bmod) * p.mod_bstride;
            prefetch_l1(&md2[n02 + et]);
            prefetch_l1(&md2[p.N + n02 + et]);
          }
        }
      }
      if (b != cur_b || n_idx != cur_n) {
        const int goff = (n0 / GS) & 7;
        if (b != cur_b || goff != cur_goff) flush_stats(cur_b, cur_goff);
        cur_goff = goff;
        named_bar(2, nthr_epi);
        cur_b = b; cur_n = n_idx;
        for (int n = et; n < BN; n += 256) {
          const int nm = (n0 + n) % p.bias_mod;
          const float cs = p.colscale ? p.colscale[(size_t)(b % p.cs_bmod) * p.cs_bstride + nm] : 1.f;
          const float rv = p.rowvec ? p.rowvec[(size_t)b * p.rowvec_stride + nm] : 0.f;
          const size_t wrow = (size_t)(b % p.w_bmod) * p.ws_bstride + n0 + n;
          ep_mul[n] = cs;
          ep_add[n] = ((p.bias ? p.bias[nm] : 0.f) + (p.addvec ? p.addvec[wrow] : 0.f)) * cs + rv;
          if (ln_fold) ep_ws[n] = p.ws[wrow];
          if (resid_mode == 2) {
            const float* md = p.mod + (size_t)(b % p.mod_bmod) * p.mod_bstride;
            ep_g[n] = 1.f + md[n0 + n];
            ep_sh[n] = md[p.N + n0 + n];
          }
        }
        named_bar(2, nthr_epi);
      }
      float r_mean = 0.f, r_rstd = 0.f;
      if ((resid_mode == 2 || ln_fold) && row_valid) {     // LayerNorm statistics of this thread's A1 / residual row
        const float* rsrc = p.rowstats_in + ((size_t)b * p.L + l0 + row) * p.rs_parts * 2;
        float a = 0.f, c = 0.f;
        for (int j = 0; j < p.rs_parts; ++j) { a += rsrc[2 * j]; c += rsrc[2 * j + 1]; }
        r_mean = a / (float)p.K1;
        r_rstd = rsqrtf(fmaxf(c / (float)p.K1 - r_mean * r_mean, 0.f) + p.eps);
      }
      const float f_mul = ln_fold ? r_rstd : 1.f, f_sub = ln_fold ? -r_rstd * r_mean : 0.f;
      if (et == 0) SK_STAMP(6, 2 * i);
      mbar_wait(&acc_full[acc], (i >> 1) & 1);
      if (et == 0) SK_STAMP(6, 2 * i + 1);
      tc_fence_after();
      const uint32_t tacc = tmem_base + acc * BN + lane_off;
      float rs0 = 0.f, rs1 = 0.f, rq0 = 0.f, rq1 = 0.f;
      const int sw7 = row & 7, sw3 = (lane >> 1) & 3;
      // The chunk loop is NOT unrolled: a fully unrolled epilogue is ~64 KB of SASS and misses the instruction cache.
#pragma unroll 1
      for (int c = par; c < NCH; c += npar) {
        const uint32_t qg = i * NCH + c;           // chunk counter shared with the residual producer
        uint32_t v[32];
        tmem_ld32(tacc + c * 32, v);
        if (et == 0) SK_STAMP(4, 4 * (qg >> 1));
        // this warp's previous stores no longer read shared memory: its T slice is free again and the previous chunk's
        // residual buffer (overwritten in place by the fp32 output) goes back to the TMA producer
        if (lane == 0) {
          bulk_wait_read<0>();
          if (resid_mode != 0 && prev_slot >= 0) mbar_arrive(&rc_empty[prev_slot]);
          prev_slot = -1;
        }
        pass_foreign_chunks(qg);                   // lane 0 observes + releases the other groups' chunks before qg
        __syncwarp();                              // ... so every lane's parity wait below is phase-exact
        // Residual ops: the R slot of this chunk (filled by the TMA producer, handed back through rc_empty).  Ops WITHOUT a
        // residual use the R ring as plain output staging: each chunk group owns slot `par` outright - a slot shared between
        // warps of different groups would be rewritten by one warp while the other warp's TMA store still reads it (only
        // the issuing thread's bulk_wait_read orders a store against later writes).  sk_pick_rings keeps nr >= npar.
        uint8_t* rt = sR + (resid_mode != 0 ? rs : par) * C::R_BYTES;         // [128 rows][128 B], 128B swizzle
        const int my_slot = rs;
        if (resid_mode != 0) mbar_wait(&rc_full[rs], phr);
        tmem_ld_wait();
        if (c + npar >= NCH) {        // this thread's last TMEM read of the tile: hand the accumulator back
          tc_fence_before();
          mbar_arrive(&acc_empty[acc]);
        }
        if (et == 0) SK_STAMP(4, 4 * (qg >> 1) + 1);
        const int c0 = c * 32;
        float y[32];
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          float4 mu = make_float4(1.f, 1.f, 1.f, 1.f);
          if (use_mul) mu = *reinterpret_cast<const float4*>(&ep_mul[c0 + j4 * 4]);
          const float4 ad = *reinterpret_cast<const float4*>(&ep_add[c0 + j4 * 4]);
          uint8_t* slot = rt + row * 128 + ((j4 ^ sw7) << 4);
          float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
          if (resid_mode != 0) x = *reinterpret_cast<const float4*>(slot);
          if (resid_mode == 2) {
            const float4 gg = *reinterpret_cast<const float4*>(&ep_g[c0 + j4 * 4]);
            const float4 sh = *reinterpret_cast<const float4*>(&ep_sh[c0 + j4 * 4]);
            x.x = fmaf((x.x - r_mean) * r_rstd, gg.x, sh.x); x.y = fmaf((x.y - r_mean) * r_rstd, gg.y, sh.y);
            x.z = fmaf((x.z - r_mean) * r_rstd, gg.z, sh.z); x.w = fmaf((x.w - r_mean) * r_rstd, gg.w, sh.w);
          }
          float d0 = __uint_as_float(v[j4 * 4 + 0]), d1 = __uint_as_float(v[j4 * 4 + 1]);
          float d2 = __uint_as_float(v[j4 * 4 + 2]), d3 = __uint_as_float(v[j4 * 4 + 3]);
          if (ln_fold && !use_mul) {   // D' + add = rstd D + (-rstd mean ws[n] + add[n])   (two FMAs per value)
            const float4 wsv = *reinterpret_cast<const float4*>(&ep_ws[c0 + j4 * 4]);
            y[j4 * 4 + 0] = fmaf(f_mul, d0, fmaf(f_sub, wsv.x, ad.x)) + x.x;
            y[j4 * 4 + 1] = fmaf(f_mul, d1, fmaf(f_sub, wsv.y, ad.y)) + x.y;
            y[j4 * 4 + 2] = fmaf(f_mul, d2, fmaf(f_sub, wsv.z, ad.z)) + x.z;
            y[j4 * 4 + 3] = fmaf(f_mul, d3, fmaf(f_sub, wsv.w, ad.w)) + x.w;
          } else {
            if (ln_fold) {            // D' = rstd (D - mean ws[n])
              const float4 wsv = *reinterpret_cast<const float4*>(&ep_ws[c0 + j4 * 4]);
              d0 = fmaf(f_mul, d0, f_sub * wsv.x); d1 = fmaf(f_mul, d1, f_sub * wsv.y);
              d2 = fmaf(f_mul, d2, f_sub * wsv.z); d3 = fmaf(f_mul, d3, f_sub * wsv.w);
            }
            y[j4 * 4 + 0] = fmaf(d0, mu.x, ad.x) + x.x;
            y[j4 * 4 + 1] = fmaf(d1, mu.y, ad.y) + x.y;
            y[j4 * 4 + 2] = fmaf(d2, mu.z, ad.z) + x.z;
            y[j4 * 4 + 3] = fmaf(d3, mu.w, ad.w) + x.w;
          }
          if (has_out_r) *reinterpret_cast<float4*>(slot) = make_float4(y[j4 * 4], y[j4 * 4 + 1], y[j4 * 4 + 2], y[j4 * 4 + 3]);
        }
        if (has_out_t) {
#pragma unroll
          for (int j8 = 0; j8 < 4; ++j8)
            *reinterpret_cast<uint4*>(tt + lane * 64 + ((j8 ^ sw3) << 4)) =
                make_uint4(pack_bf16(y[j8 * 8], y[j8 * 8 + 1]), pack_bf16(y[j8 * 8 + 2], y[j8 * 8 + 3]),
                           pack_bf16(y[j8 * 8 + 4], y[j8 * 8 + 5]), pack_bf16(y[j8 * 8 + 6], y[j8 * 8 + 7]));
        }
        if (row_valid) {
          if (p.stats_out != nullptr) {
            constexpr int NSB = GS >= 32 ? 1 : 32 / GS;      // sub-blocks of one group inside the 32 columns
            constexpr int SBW = 32 / NSB;
            const int off = (c0 / GS) & 7;                   // local group of sub-block sb = (off + sb) & 7
#pragma unroll
            for (int sb = 0; sb < NSB; ++sb) {
              float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;  // two independent chains per sum
#pragma unroll
              for (int j = 0; j < SBW; j += 2) {
                a0 += y[sb * SBW + j]; a1 += y[sb * SBW + j + 1];
                b0 = fmaf(y[sb * SBW + j], y[sb * SBW + j], b0); b1 = fmaf(y[sb * SBW + j + 1], y[sb * SBW + j + 1], b1);
              }
              a0 += a1; b0 += b1;
#pragma unroll
              for (int k = 0; k < 8; ++k)
                if (((off + sb) & 7) == k) { s1[k] += a0; s2[k] += b0; }
            }
          }
          if (p.rowstats_out != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              rs0 += y[j]; rs1 += y[j + 1];
              rq0 = fmaf(y[j], y[j], rq0); rq1 = fmaf(y[j + 1], y[j + 1], rq1);
            }
          }
        }
        if (et == 0) SK_STAMP(4, 4 * (qg >> 1) + 2);
        fence_proxy_async();          // generic-proxy writes of this warp's slices -> visible to the TMA (async proxy)
        __syncwarp();
        if (lane == 0) {
          if (has_out_r) tma_store_3d(&p.tmRs, rt + q4 * 4096, n0 + c0, l0 + q4 * 32, b);
          if (has_out_t) tma_store_3d(&p.tmT, tt, n0 + c0, l0 + q4 * 32, b);
          bulk_commit();
          if (et == 0) SK_STAMP(4, 4 * (qg >> 1) + 3);
          if (epi12 && resid_mode != 0) {      // three warps per sub-partition hide this wait; the slot returns a chunk earlier
            bulk_wait_read<0>();
            mbar_arrive(&rc_empty[my_slot]);
          }
        }
        prev_slot = epi12 ? -1 : my_slot;
        ++q_cur;                               // this warp's own chunk is passed
        if (++rs == NR) { rs = 0; phr ^= 1; }
      }
      if (p.rowstats_out != nullptr && row_valid) {      // two partial sums per (row, n tile): one per chunk parity
        float* dst = p.rowstats_out + (((size_t)b * p.L + l0 + row) * (2 * p.n_tiles) + 2 * n_idx + par) * 2;
        dst[0] = rs0 + rs1;
        dst[1] = rq0 + rq1;
      }
    }
    pass_foreign_chunks((uint32_t)(t_end - t_begin) * NCH);     // trailing chunks of the other groups (uniform accounting)
    flush_stats(cur_b, cur_goff);
    if (lane == 0) {                    // every thread that issued stores waits for its bulk groups before exit
      if (et == 0) SK_STAMP(7, 1);
      bulk_wait<0>();
      if (et == 0) SK_STAMP(7, 2);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 3) tmem_dealloc(tmem_base, 2 * BN);
}

// ------------------------------------------------------------------------------------------------ instantiation table
#define SFB_SK_LIST(X) X(128, 4) X(128, 8) X(128, 16) X(256, 8) X(256, 16) X(256, 32) X(256, 64) X(256, 128)

inline int sk_find(int BN, int GS) {
  int i = 0;
#define X(a, b) if (BN == a && GS == b) return i; ++i;
  SFB_SK_LIST(X)
#undef X
  return -1;
}
inline cudaError_t sk_set_attrs() {
  cudaError_t e = cudaSuccess;
#define Y(a, b, c) if (e == cudaSuccess) e = cudaFuncSetAttribute(sk_kernel<a, b, c>, cudaFuncAttributeMaxDynamicSharedMemorySize, SkCfg<a>::kMaxSmem);
#define X(a, b) Y(a, b, 0) Y(a, b, 1) Y(a, b, 2) Y(a, b, 3) Y(a, b, 4) Y(a, b, 5)
  SFB_SK_LIST(X)
#undef X
#undef Y
  return e;
}
inline void sk_launch(int id, int epi, const SkParams& p, int num_sms, cudaStream_t st) {
  const int grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
  int i = 0;
#define Y(a, b, c) if (epi == c) { launch_pdl(sk_kernel<a, b, c>, grid, 512, SkCfg<a>::smem_bytes(p.na, p.nb, p.nr, p.epi12), st, p); return; }
#define X(a, b) if (id == i++) { Y(a, b, 0) Y(a, b, 1) Y(a, b, 2) Y(a, b, 3) Y(a, b, 4) Y(a, b, 5) }
  SFB_SK_LIST(X)
#undef X
#undef Y
}

}  // namespace sfb
