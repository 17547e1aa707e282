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
  static_assert(kGemmBM * (BN + 4) * 4 <= kGemmStages * kStage, "staging tile must fit in the pipeline buffers");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kGemmStages * kStage);
  uint64_t* empty_bar = full_bar + kGemmStages;
  uint64_t* tmem_full_bar = empty_bar + kGemmStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  double* sacc = reinterpret_cast<double*>(tmem_slot + 2);   // [8][2] group partials (fp64: no cancellation in var)

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
  if (threadIdx.x < 16) sacc[threadIdx.x] = 0.0;
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
    constexpr int LD = BN + 4;               // row pitch: 16-byte aligned rows, conflict-free 128-bit accesses
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    {   // phase 1: TMEM -> registers -> staging tile (thread = accumulator row)
      const int r = q * 32 + lane;
      const uint32_t trow = tmem_base + (uint32_t(q * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t v[32];
        tmem_ld32(trow + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<uint4*>(&stile[r * LD + c + j]) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
    }
    tc_fence_before();
    asm volatile("bar.sync 1, 128;" ::: "memory");
    // phase 2: coalesced row-major pass, 4 columns per lane (x2 for BN = 256), U rows in flight per thread
    const int rows_valid = min(kGemmBM, p.rows_per_clip - l0);
    constexpr int LPR = BN >= 128 ? 32 : BN / 4;   // lanes per row
    constexpr int RPI = 32 / LPR;                  // rows per warp iteration
    constexpr int NCH = BN >= 256 ? 2 : 1;         // float4 chunks per lane per row
    constexpr int ITERS = 32 / RPI;                // iterations per warp (4 warps x RPI rows x ITERS = 128 rows)
    constexpr int U = 4;
    const int ew = warp - 2;
    const int sub = lane % LPR, rsub = lane / LPR;
    float4 bias4[NCH], cs4[NCH], rv4[NCH];
    int nm0[NCH];
#pragma unroll
    for (int jj = 0; jj < NCH; ++jj) {
      const int col = jj * 128 + sub * 4;
      nm0[jj] = (n0 + col) % p.bias_mod;           // 4 | bias_mod and 4 | (n0 + col): the 4 columns never wrap
      bias4[jj] = p.bias ? *reinterpret_cast<const float4*>(p.bias + nm0[jj]) : make_float4(0.f, 0.f, 0.f, 0.f);
      cs4[jj] = p.colscale ? *reinterpret_cast<const float4*>(p.colscale + (size_t)(b % (p.cs_bmod > 0 ? p.cs_bmod : 1)) * p.cs_bstride + nm0[jj])
                           : make_float4(1.f, 1.f, 1.f, 1.f);
      rv4[jj] = p.rowvec ? *reinterpret_cast<const float4*>(p.rowvec + (size_t)b * p.rowvec_stride + nm0[jj])
                         : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    double s1[NCH][4], s2[NCH][4];   // fp64 accumulation (only when stats are requested)
#pragma unroll
    for (int jj = 0; jj < NCH; ++jj)
#pragma unroll
      for (int i = 0; i < 4; ++i) { s1[jj][i] = 0.0; s2[jj][i] = 0.0; }
    const size_t gbase = ((size_t)b * p.rows_per_clip + l0) * p.N + n0 + sub * 4;
#pragma unroll 1
    for (int it0 = 0; it0 < ITERS; it0 += U) {
      float4 res[U][NCH];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int r = (it0 + u) * (4 * RPI) + ew * RPI + rsub;
#pragma unroll
        for (int jj = 0; jj < NCH; ++jj)
          res[u][jj] = (p.resid && r < rows_valid) ? *reinterpret_cast<const float4*>(p.resid + gbase + (size_t)r * p.N + jj * 128)
                                                   : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int r = (it0 + u) * (4 * RPI) + ew * RPI + rsub;
        if (r < rows_valid) {
#pragma unroll
          for (int jj = 0; jj < NCH; ++jj) {
            const float4 a = *reinterpret_cast<const float4*>(&stile[r * LD + jj * 128 + sub * 4]);
            float v[4];
            v[0] = (a.x + bias4[jj].x) * cs4[jj].x + rv4[jj].x + res[u][jj].x;
            v[1] = (a.y + bias4[jj].y) * cs4[jj].y + rv4[jj].y + res[u][jj].y;
            v[2] = (a.z + bias4[jj].z) * cs4[jj].z + rv4[jj].z + res[u][jj].z;
            v[3] = (a.w + bias4[jj].w) * cs4[jj].w + rv4[jj].w + res[u][jj].w;
            const size_t g = gbase + (size_t)r * p.N + jj * 128;
            if (p.out_r) *reinterpret_cast<float4*>(p.out_r + g) = make_float4(v[0], v[1], v[2], v[3]);
            if (p.out_t) {
              if constexpr (sizeof(T) == 2) {
                __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
                *reinterpret_cast<uint2*>(p.out_t + g) = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
              } else {
                *reinterpret_cast<float4*>(p.out_t + g) = make_float4(from_f32<float>(v[0]), from_f32<float>(v[1]), from_f32<float>(v[2]), from_f32<float>(v[3]));
              }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if (p.stats) { const double dv = (double)v[i]; s1[jj][i] += dv; s2[jj][i] += dv * dv; }
            }
          }
        }
      }
    }
    if (p.stats) {
#pragma unroll
      for (int jj = 0; jj < NCH; ++jj) {
        if (p.gs >= 4) {     // the lane's 4 columns share a group
          const int grp = nm0[jj] / p.gs;
          atomicAdd(&sacc[grp * 2 + 0], s1[jj][0] + s1[jj][1] + s1[jj][2] + s1[jj][3]);
          atomicAdd(&sacc[grp * 2 + 1], s2[jj][0] + s2[jj][1] + s2[jj][2] + s2[jj][3]);
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int grp = (nm0[jj] + i) / p.gs;
            atomicAdd(&sacc[grp * 2 + 0], s1[jj][i]);
            atomicAdd(&sacc[grp * 2 + 1], s2[jj][i]);
          }
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (e < 16 && sacc[e] != 0.0) atomicAdd(&p.stats[(size_t)b * 16 + e], sacc[e]);
    }
  }
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, BN);
}

}  // namespace sfb
