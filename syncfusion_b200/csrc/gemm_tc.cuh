// K1/K4/K5a/K5c/K7/K8: channels-last implicit-GEMM for every dense contraction of the U-Net:
//   ResNet Conv1d k=3 (a6), InjectChannels 1x1 over cat[x, ctx] (a8, two A sources, no cat), attention
//   projections (a9), patchify Down (a11) and un-patchify / nearest+conv3 Up + SkipModulate (a12, a5).
//
//   D[b, l, n] = sum_{tap, k} A[b, l + tap - pad, k] * W[tap * N + n, k]        (fp32 accumulate in TMEM)
//   out        = resid + colscale[n'] * (D + bias[n']) + rowvec[b, n']          n' = n % bias_mod
//
// PERSISTENT kernel: grid = min(#tiles, SMs x CTAs/SM); every CTA walks 128 x BN output tiles (one clip each, n
// fastest so co-resident CTAs share the A rows in L2).  Three overlapped pipelines (blackwell guide, "Anatomy"):
//   warp 0 lane 0  TMA producer: 3-D maps [C, L, B]; the conv halo and the clip boundaries are TMA out-of-bounds zero
//                  fill, so clips never bleed into each other; runs ahead across tiles through the smem stage ring.
//   warp 1 lane 0  tcgen05.mma issuer: SS operands, 128-byte-swizzled K-major smem, fp32 accumulators in TMEM,
//                  double-buffered (2 x BN columns) so tile i+1 accumulates while tile i drains.
//   warps 2-5      epilogue: L2-prefetch of the residual tile, then per 64-column chunk TMEM -> registers -> smem
//                  staging -> coalesced row-major pass (128-bit accesses, 4 rows in flight per thread) that applies
//                  bias / SkipModulate scale / cross-attention bias / residual, writes the fp32 residual-stream copy
//                  and/or the operand-precision copy, and accumulates the GroupNorm statistics (fp64 sum, sum of
//                  squares per (clip, group)) of the OUTPUT, so no separate statistics pass over HBM is ever made.
#pragma once
#include "ptx.cuh"
#undef SFB_FILE_ID
#define SFB_FILE_ID 1   // gemm_tc.cuh (wait-log call sites, ptx.cuh)

namespace sfb {

template <typename T>
struct GemmParams {
  CUtensorMap tmA1;   // [K1, L, B]      box [BK, 128, 1]
  CUtensorMap tmA2;   // [K2, L, B2]     box [BK, 128, 1]   (inject context; unused if k2_chunks == 0)
  CUtensorMap tmW;    // [Kw, taps * N]  box [BK, BN]
  int rows_per_clip;  // L (valid rows per clip)
  int tiles_per_clip; // ceil(L / 128)
  int N;              // total output columns
  int taps;           // 1 or 3 (3 => rows l-1, l, l+1)
  int k1_chunks;      // ceil(K1 / BK)
  int k2_chunks;      // ceil(K2 / BK) or 0
  int K1;             // weight column offset of the A2 segment
  int a2_bmod;        // A2 clip index = b % a2_bmod (CFG branches share the onset pyramid)
  int bias_mod;       // n' = n % bias_mod
  int gs;             // GroupNorm group size in channels (stats group = n' / gs)
  int rowvec_stride;
  int cs_bstride;      // colscale row = (b % cs_bmod) * cs_bstride
  int cs_bmod;
  const float* bias;      // [bias_mod] or null
  const float* colscale;  // [bias_mod] or null   (SkipModulate scale for this step)
  const float* rowvec;    // [B, rowvec_stride] or null (cross-attention bias, M_ctx = 1 fast path)
  const float* resid;     // [B, L, N] fp32 or null
  float* out_r;           // [B, L, N] fp32 residual-stream copy or null
  T* out_t;               // [B, L, N] operand-dtype copy or null
  double* stats;          // [B, 8, 2] or null
  int n_tiles;            // N / BN
  int total_tiles;        // B * tiles_per_clip * n_tiles
  int tag;                // plan op index (wait log)
};

constexpr int kGemmBM = 128;
constexpr int kGemmThreads = 192;
template <int BN> struct GemmCfg {
  // the stage ring spans tiles: narrow (bandwidth-bound) tiles need many stages in flight to cover HBM latency
  static constexpr int kStages = BN == 32 ? 8 : (BN == 64 ? 6 : (BN == 128 ? 4 : 3));
  static constexpr int kCtasPerSm = 1;
  static constexpr int kCW = 32;                          // epilogue chunk width (columns)
  static constexpr int kResBufs = BN == 32 ? 2 : 4;       // residual chunks in flight (cp.async ring)
};

template <typename T, int BN>
__host__ __device__ constexpr int gemm_stage_bytes() { return kGemmBM * 128 + BN * 128; }
template <typename T, int BN>
__host__ __device__ constexpr int gemm_smem_bytes() {
  return GemmCfg<BN>::kStages * gemm_stage_bytes<T, BN>() + kGemmBM * (GemmCfg<BN>::kCW + 4) * 4 +
         GemmCfg<BN>::kResBufs * kGemmBM * GemmCfg<BN>::kCW * 4 + 1024 /*barriers + stats slots*/;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <typename T, int BN>
__global__ void __launch_bounds__(kGemmThreads, GemmCfg<BN>::kCtasPerSm) gemm_tc_kernel(const __grid_constant__ GemmParams<T> p) {
  pdl_trigger();
  using TR = ElemTraits<T>;
  constexpr int BK = TR::kAtomElems;
  constexpr int S = GemmCfg<BN>::kStages;
  constexpr int CW = GemmCfg<BN>::kCW;
  constexpr int kStage = gemm_stage_bytes<T, BN>();
  constexpr int kABytes = kGemmBM * 128;
  constexpr int kBBytes = BN * 128;
  constexpr int LD = CW + 4;               // staging row pitch: 16-byte aligned rows, conflict-free 128-bit accesses

  extern __shared__ __align__(1024) uint8_t smem[];   // 128B-swizzle atoms need 1024-byte alignment
  constexpr int NB = GemmCfg<BN>::kResBufs;
  float* stile = reinterpret_cast<float*>(smem + S * kStage);
  float* rbuf = stile + kGemmBM * LD;            // [NB][128][CW] residual chunks, each thread re-reads only its own copies
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S * kStage + kGemmBM * LD * 4 + NB * kGemmBM * CW * 4);
  uint64_t* empty_bar = full_bar + S;
  uint64_t* tmem_full_bar = empty_bar + S;       // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  double* sacc = reinterpret_cast<double*>(tmem_slot + 2);   // [4 warps][8][2] group partials (fp64: no cancellation in var)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_k = p.taps * p.k1_chunks + p.k2_chunks;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmA1);
    tma_prefetch_desc(&p.tmW);
    if (p.k2_chunks) tma_prefetch_desc(&p.tmA2);
    for (int s = 0; s < S; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full_bar[a], 1); mbar_init(&tmem_empty_bar[a], 128); }
    fence_barrier_init();
  }
  if (threadIdx.x < 64) sacc[threadIdx.x] = 0.0;
  if (warp == 1) { tmem_alloc(tmem_slot, 2 * BN); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  pdl_wait();            // everything above is independent of the previous kernel's output
  mark_progress(p.tag);
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------- TMA producer
    const int pad = (p.taps == 3) ? 1 : 0;
    uint32_t it = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      const int n0 = (t % p.n_tiles) * BN;
      const int mi = t / p.n_tiles;
      const int b = mi / p.tiles_per_clip;
      const int l0 = (mi % p.tiles_per_clip) * kGemmBM;
      for (int tap = 0; tap < p.taps; ++tap) {
        for (int k1 = 0; k1 < p.k1_chunks; ++k1, ++it) {
          const int s = it % S;
          mbar_wait(&empty_bar[s], ((it / S) & 1) ^ 1);
          mbar_expect_tx(&full_bar[s], kABytes + kBBytes);
          uint8_t* sa = smem + s * kStage;
          tma_load_3d(sa, &p.tmA1, &full_bar[s], k1 * BK, l0 + tap - pad, b);
          tma_load_2d(sa + kABytes, &p.tmW, &full_bar[s], k1 * BK, tap * p.N + n0);
        }
      }
      for (int k2 = 0; k2 < p.k2_chunks; ++k2, ++it) {
        const int s = it % S;
        mbar_wait(&empty_bar[s], ((it / S) & 1) ^ 1);
        mbar_expect_tx(&full_bar[s], kABytes + kBBytes);
        uint8_t* sa = smem + s * kStage;
        tma_load_3d(sa, &p.tmA2, &full_bar[s], k2 * BK, l0, b % p.a2_bmod);
        tma_load_2d(sa + kABytes, &p.tmW, &full_bar[s], p.K1 + k2 * BK, n0);
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ------------------------------------------------------------- MMA issuer
    constexpr uint32_t idesc = make_idesc(TR::kFmt, kGemmBM, BN, 0, 0);
    uint32_t it = 0, i = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++i) {
      const uint32_t acc = i & 1;
      mbar_wait(&tmem_empty_bar[acc], ((i >> 1) & 1) ^ 1);   // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t tacc = tmem_base + acc * BN;
      for (int kc = 0; kc < num_k; ++kc, ++it) {
        const int s = it % S;
        mbar_wait(&full_bar[s], (it / S) & 1);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * kStage);
        const uint64_t da = make_smem_desc_sw128(sa, 16, 1024);
        const uint64_t db = make_smem_desc_sw128(sa + kABytes, 16, 1024);
#pragma unroll
        for (int k = 0; k < BK / TR::kUmmaK; ++k) {
          // advance 32 bytes (one UMMA_K slice) inside the 128-byte swizzle row: +2 in the 16-byte address field
          umma_ss<TR::kTF32>(tacc, da + uint64_t(k * 2), db + uint64_t(k * 2), idesc, (kc | k) != 0);
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(&tmem_full_bar[acc]);
    }
  } else if (warp >= 2) {
    // ------------------------------------------------------------- epilogue
    const int q = warp & 3;                  // TMEM lane quarter this warp may read
    const int e = threadIdx.x - 64;          // 0..127
    const int ew = warp - 2;
    constexpr int LPR = CW / 4;              // lanes per row in the coalesced pass
    constexpr int RPI = 32 / LPR;            // rows per warp iteration
    constexpr int ITERS = 32 / RPI;          // 4 warps x RPI rows x ITERS = 128 rows
    constexpr int U = 4;
    const int sub = lane % LPR, rsub = lane / LPR;
    constexpr int NC = BN / CW;              // chunks per tile
    // residual prefetch ring: chunk g (tile-major over this CTA's tiles) -> buffer g % NB; every thread copies exactly
    // the 16-byte pieces it will consume itself, so cp.async.wait_group is the only synchronisation needed.
    auto issue_resid = [&](uint32_t g) {
      if (p.resid) {
        const int t = blockIdx.x + (int)(g / NC) * gridDim.x;
        if (t < p.total_tiles) {
          const int n0 = (t % p.n_tiles) * BN;
          const int mi = t / p.n_tiles;
          const int b = mi / p.tiles_per_clip;
          const int l0 = (mi % p.tiles_per_clip) * kGemmBM;
          const int rows_valid = min(kGemmBM, p.rows_per_clip - l0);
          const float* src = p.resid + ((size_t)b * p.rows_per_clip + l0) * p.N + n0 + (g % NC) * CW + sub * 4;
          float* dst = rbuf + (size_t)(g % NB) * kGemmBM * CW + sub * 4;
#pragma unroll
          for (int it = 0; it < ITERS; ++it) {
            const int r = it * (4 * RPI) + ew * RPI + rsub;
            if (r < rows_valid) cp_async16(dst + r * CW, src + (size_t)r * p.N);
          }
        }
      }
      cp_async_commit();
    };
#pragma unroll
    for (int g = 0; g < NB - 1; ++g) issue_resid(g);
    uint32_t i = 0, g = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++i) {
      const int n0 = (t % p.n_tiles) * BN;
      const int mi = t / p.n_tiles;
      const int b = mi / p.tiles_per_clip;
      const int l0 = (mi % p.tiles_per_clip) * kGemmBM;
      const int rows_valid = min(kGemmBM, p.rows_per_clip - l0);
      const size_t tile_base = ((size_t)b * p.rows_per_clip + l0) * p.N + n0;
      const uint32_t acc = i & 1;
      mbar_wait(&tmem_full_bar[acc], (i >> 1) & 1);
      tc_fence_after();
      const uint32_t trow = tmem_base + acc * BN + (uint32_t(q * 32) << 16);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += CW, ++g) {
        {   // phase 1: TMEM -> registers -> staging chunk (thread = accumulator row)
          const int r = q * 32 + lane;
          uint32_t v[32];
          tmem_ld32(trow + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<uint4*>(&stile[r * LD + j]) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
        if (c0 + CW >= BN) {                 // all TMEM reads of this tile are done: hand the accumulator back
          tc_fence_before();
          mbar_arrive(&tmem_empty_bar[acc]);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        issue_resid(g + NB - 1);             // refill the buffer this thread finished reading one chunk ago
        cp_async_wait<NB - 1>();             // chunk g of the residual has landed (groups retire in order)
        // phase 2: coalesced row-major pass over the chunk, 4 columns per lane
        const int col = c0 + sub * 4;
        const int nm0 = (n0 + col) % p.bias_mod;     // 4 | bias_mod and 4 | (n0 + col): the 4 columns never wrap
        const float4 bias4 = p.bias ? *reinterpret_cast<const float4*>(p.bias + nm0) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 cs4 = p.colscale ? *reinterpret_cast<const float4*>(p.colscale + (size_t)(b % (p.cs_bmod > 0 ? p.cs_bmod : 1)) * p.cs_bstride + nm0)
                                      : make_float4(1.f, 1.f, 1.f, 1.f);
        const float4 rv4 = p.rowvec ? *reinterpret_cast<const float4*>(p.rowvec + (size_t)b * p.rowvec_stride + nm0)
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
        double s1[4] = {0.0, 0.0, 0.0, 0.0}, s2[4] = {0.0, 0.0, 0.0, 0.0};
        const size_t gbase = tile_base + col;
        const float* rb = rbuf + (size_t)(g % NB) * kGemmBM * CW + sub * 4;
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
          const int r = it * (4 * RPI) + ew * RPI + rsub;
          if (r < rows_valid) {
            const float4 a = *reinterpret_cast<const float4*>(&stile[r * LD + sub * 4]);
            const float4 res = p.resid ? *reinterpret_cast<const float4*>(rb + r * CW) : make_float4(0.f, 0.f, 0.f, 0.f);
            float v[4];
            v[0] = (a.x + bias4.x) * cs4.x + rv4.x + res.x;
            v[1] = (a.y + bias4.y) * cs4.y + rv4.y + res.y;
            v[2] = (a.z + bias4.z) * cs4.z + rv4.z + res.z;
            v[3] = (a.w + bias4.w) * cs4.w + rv4.w + res.w;
            const size_t gi = gbase + (size_t)r * p.N;
            if (p.out_r) *reinterpret_cast<float4*>(p.out_r + gi) = make_float4(v[0], v[1], v[2], v[3]);
            if (p.out_t) {
              if constexpr (sizeof(T) == 2) {
                __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
                *reinterpret_cast<uint2*>(p.out_t + gi) = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
              } else {
                *reinterpret_cast<float4*>(p.out_t + gi) = make_float4(from_f32<float>(v[0]), from_f32<float>(v[1]), from_f32<float>(v[2]), from_f32<float>(v[3]));
              }
            }
            if (p.stats) {
#pragma unroll
              for (int k = 0; k < 4; ++k) { const double dv = (double)v[k]; s1[k] += dv; s2[k] += dv * dv; }
            }
          }
        }
        if (p.stats) {
          // deterministic, atomic-free reduction (shared-memory fp64 atomics are CAS loops): first across the RPI
          // row-lanes that share this lane's columns, then into this warp's private slots.
#pragma unroll
          for (int k = 0; k < 4; ++k) {
#pragma unroll
            for (int o = LPR; o < 32; o <<= 1) {
              s1[k] += __shfl_xor_sync(0xffffffffu, s1[k], o);
              s2[k] += __shfl_xor_sync(0xffffffffu, s2[k], o);
            }
          }
          double* slot = sacc + ew * 16;
          if (p.gs >= 4) {       // the lane's 4 columns share a group; gs / 4 adjacent lanes share it too
            double a = s1[0] + s1[1] + s1[2] + s1[3], q2 = s2[0] + s2[1] + s2[2] + s2[3];
            const int lpg = min(p.gs / 4, LPR);
            for (int o = 1; o < lpg; o <<= 1) {
              a += __shfl_xor_sync(0xffffffffu, a, o);
              q2 += __shfl_xor_sync(0xffffffffu, q2, o);
            }
            if (rsub == 0 && (sub % lpg) == 0) {   // group leaders own distinct slots: no conflicts
              const int grp = nm0 / p.gs;
              slot[grp * 2 + 0] += a;
              slot[grp * 2 + 1] += q2;
            }
          } else if (rsub == 0) {   // gs < 4 (8-channel outputs): lanes 0..LPR-1 take turns on the warp's slots
            for (int l = 0; l < LPR; ++l) {
              if (sub == l) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const int grp = (nm0 + k) / p.gs;
                  slot[grp * 2 + 0] += s1[k];
                  slot[grp * 2 + 1] += s2[k];
                }
              }
              __syncwarp((1u << LPR) - 1u);
            }
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");   // staging chunk free again; sacc atomics of this chunk done
      }
      if (p.stats && e < 16) {
        const double v = sacc[e] + sacc[16 + e] + sacc[32 + e] + sacc[48 + e];
        if (v != 0.0) atomicAdd(&p.stats[(size_t)b * 16 + e], v);
        sacc[e] = 0.0; sacc[16 + e] = 0.0; sacc[32 + e] = 0.0; sacc[48 + e] = 0.0;
      }
    }
    cp_async_wait<0>();
  }
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 2 * BN);
}

}  // namespace sfb
