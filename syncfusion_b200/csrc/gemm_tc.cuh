// K1/K4/K5a/K5c/K7/K8: channels-last implicit-GEMM for every dense contraction of the U-Net:
//   ResNet Conv1d k=3 (a6), InjectChannels 1x1 over cat[x, ctx] (a8, two A sources, no cat), attention
//   projections (a9), patchify Down (a11) and un-patchify / nearest+conv3 Up + SkipModulate (a12, a5).
//
//   D[b, l, n] = sum_{tap, k} A[b, l + tap - pad, k] * W[tap * N + n, k]        (fp32 accumulate in TMEM)
//   out        = resid + colscale[n'] * (D + bias[n']) + rowvec[b, n']          n' = n % bias_mod
//
// One CTA = one 128 x BN output tile of one clip.  Warp 0 lane 0: TMA producer (3-D maps [C, L, B]; the conv
// halo and clip boundaries are TMA out-of-bounds zero fill, so clips never bleed into each other).  Warp 1 lane 0:
// tcgen05.mma issuer (SS operands, 128-byte-swizzled K-major smem, accumulator in TMEM).  Warps 2-5: epilogue -
// TMEM -> registers -> (+bias) -> smem staging tile (re-using the drained pipeline buffers) -> coalesced
// column-per-thread pass that adds the residual, writes the fp32 residual-stream copy and/or the operand-dtype
// copy, and accumulates the GroupNorm statistics (sum, sum of squares per (clip, group)) of the *output* for the
// next GroupNorm, so no separate statistics pass over HBM is ever made.
#pragma once
#include "ptx.cuh"

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
};

constexpr int kGemmBM = 128;
constexpr int kGemmStages = 4;
constexpr int kGemmThreads = 192;

template <typename T, int BN>
__host__ __device__ constexpr int gemm_stage_bytes() { return kGemmBM * 128 + BN * 128; }
template <typename T, int BN>
__host__ __device__ constexpr int gemm_smem_bytes() { return kGemmStages * gemm_stage_bytes<T, BN>() + 1024 /*align*/ + 256 /*barriers*/; }

template <typename T, int BN>
__global__ void __launch_bounds__(kGemmThreads, 1) gemm_tc_kernel(const __grid_constant__ GemmParams<T> p) {
  using TR = ElemTraits<T>;
  constexpr int BK = TR::kAtomElems;
  constexpr int kStage = gemm_stage_bytes<T, BN>();
  constexpr int kABytes = kGemmBM * 128;
  constexpr int kBBytes = BN * 128;
  static_assert(kGemmBM * (BN + 1) * 4 <= kGemmStages * kStage, "staging tile must fit in the pipeline buffers");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kGemmStages * kStage);
  uint64_t* empty_bar = full_bar + kGemmStages;
  uint64_t* tmem_full_bar = empty_bar + kGemmStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  float* sacc = reinterpret_cast<float*>(tmem_slot + 2);   // [8][2] group partials

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x / p.tiles_per_clip;
  const int l0 = (blockIdx.x % p.tiles_per_clip) * kGemmBM;
  const int n0 = blockIdx.y * BN;
  const int num_k = p.taps * p.k1_chunks + p.k2_chunks;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmA1);
    tma_prefetch_desc(&p.tmW);
    if (p.k2_chunks) tma_prefetch_desc(&p.tmA2);
    for (int s = 0; s < kGemmStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 16) sacc[threadIdx.x] = 0.f;
  if (warp == 1) { tmem_alloc(tmem_slot, BN); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------- TMA producer
    const int pad = (p.taps == 3) ? 1 : 0;
    int kc = 0;
    for (int tap = 0; tap < p.taps; ++tap) {
      for (int k1 = 0; k1 < p.k1_chunks; ++k1, ++kc) {
        const int s = kc % kGemmStages;
        mbar_wait(&empty_bar[s], ((kc / kGemmStages) & 1) ^ 1);
        mbar_expect_tx(&full_bar[s], kABytes + kBBytes);
        uint8_t* sa = smem + s * kStage;
        tma_load_3d(sa, &p.tmA1, &full_bar[s], k1 * BK, l0 + tap - pad, b);
        tma_load_2d(sa + kABytes, &p.tmW, &full_bar[s], k1 * BK, tap * p.N + n0);
      }
    }
    for (int k2 = 0; k2 < p.k2_chunks; ++k2, ++kc) {
      const int s = kc % kGemmStages;
      mbar_wait(&empty_bar[s], ((kc / kGemmStages) & 1) ^ 1);
      mbar_expect_tx(&full_bar[s], kABytes + kBBytes);
      uint8_t* sa = smem + s * kStage;
      tma_load_3d(sa, &p.tmA2, &full_bar[s], k2 * BK, l0, b % p.a2_bmod);
      tma_load_2d(sa + kABytes, &p.tmW, &full_bar[s], p.K1 + k2 * BK, n0);
    }
  } else if (warp == 1 && lane == 0) {
    // ------------------------------------------------------------- MMA issuer
    constexpr uint32_t idesc = make_idesc(TR::kFmt, kGemmBM, BN, 0, 0);
    for (int kc = 0; kc < num_k; ++kc) {
      const int s = kc % kGemmStages;
      mbar_wait(&full_bar[s], (kc / kGemmStages) & 1);
      tc_fence_after();
      const uint32_t sa = smem_u32(smem + s * kStage);
      const uint64_t da = make_smem_desc_sw128(sa, 16, 1024);
      const uint64_t db = make_smem_desc_sw128(sa + kABytes, 16, 1024);
#pragma unroll
      for (int k = 0; k < BK / TR::kUmmaK; ++k) {
        // advance 32 bytes (one UMMA_K slice) inside the 128-byte swizzle row: +2 in the 16-byte address field
        umma_ss<TR::kTF32>(tmem_base, da + uint64_t(k * 2), db + uint64_t(k * 2), idesc, (kc | k) != 0);
      }
      umma_commit(&empty_bar[s]);
    }
    umma_commit(tmem_full_bar);
  } else if (warp >= 2) {
    // ------------------------------------------------------------- epilogue
    const int q = warp & 3;                  // TMEM lane quarter this warp may read
    const int e = threadIdx.x - 64;          // 0..127
    float* stile = reinterpret_cast<float*>(smem);
    constexpr int LD = BN + 1;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    {
      const int r = q * 32 + lane;
      const uint32_t trow = tmem_base + (uint32_t(q * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t v[32];
        tmem_ld32(trow + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float a = __uint_as_float(v[j]);
          if (p.bias) a += __ldg(&p.bias[(n0 + c + j) % p.bias_mod]);
          stile[r * LD + c + j] = a;
        }
      }
    }
    tc_fence_before();
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const int rows_valid = min(kGemmBM, p.rows_per_clip - l0);
    constexpr int CP = BN < 128 ? BN : 128;  // columns handled per pass
    constexpr int RG = 128 / CP;             // row groups
    const int rg = e / CP;
#pragma unroll 1
    for (int cb = 0; cb < BN; cb += CP) {
      const int col = cb + (e % CP);
      const int n = n0 + col;
      const int nm = n % p.bias_mod;
      const float cs = p.colscale ? __ldg(&p.colscale[(size_t)(b % (p.cs_bmod > 0 ? p.cs_bmod : 1)) * p.cs_bstride + nm]) : 1.f;
      const float rv = p.rowvec ? __ldg(&p.rowvec[(size_t)b * p.rowvec_stride + nm]) : 0.f;
      float s1 = 0.f, s2 = 0.f;
      size_t g = ((size_t)b * p.rows_per_clip + l0 + rg) * p.N + n;
      const size_t gstep = (size_t)RG * p.N;
      for (int r = rg; r < rows_valid; r += RG, g += gstep) {
        float v = stile[r * LD + col] * cs + rv;
        if (p.resid) v += __ldg(&p.resid[g]);
        if (p.out_r) p.out_r[g] = v;
        if (p.out_t) p.out_t[g] = from_f32<T>(v);
        s1 += v;
        s2 += v * v;
      }
      if (p.stats) {
        const int grp = nm / p.gs;
        atomicAdd(&sacc[grp * 2 + 0], s1);
        atomicAdd(&sacc[grp * 2 + 1], s2);
      }
    }
    if (p.stats) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (e < 16 && sacc[e] != 0.f) atomicAdd(&p.stats[(size_t)b * 16 + e], (double)sacc[e]);
    }
  }
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, BN);
}

}  // namespace sfb
