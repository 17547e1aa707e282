// GPU experiment: can a SWIZZLE_128B K-major UMMA A descriptor start at a row offset that is NOT a multiple of 8 rows
// (start address + r * 128 bytes)?  If yes, a k=3 conv can load its A tile once (rows l0-1 .. l0+128) and issue the
// three taps as row-shifted descriptors of the same smem tile.
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O2 -o tools/exp_shift tools/exp_shift.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "../syncfusion_b200/csrc/ptx.cuh"

using namespace sfb;
typedef __nv_bfloat16 bf16;

constexpr int N = 64, K = 64, ROWS = 136;

__global__ void __launch_bounds__(128) k_shift(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                                                float* out, int l0, int mode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                       // 136 rows x 128 B = 17408 B (17 x 1024)
  uint8_t* sW = smem + 18432;               // 3 x 64 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 18432 + 3 * 8192);
  uint64_t* mbar = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(mbar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(slot, 64); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, ROWS * 128 + 3 * 64 * 128);
    tma_load_2d(sA, &tmA, bar, 0, l0 - 1);
    for (int t = 0; t < 3; ++t) tma_load_2d(sW + t * 8192, &tmW, bar, 0, t * N);
    mbar_wait(bar, 0);
    tc_fence_after();
    constexpr uint32_t idesc = make_idesc(1, 128, N, 0, 0);
    for (int t = 0; t < 3; ++t) {
      const uint32_t a = smem_u32(sA) + t * 128;
      uint64_t da = make_smem_desc_sw128(a, 16, 1024);
      if (mode == 1) da |= (uint64_t)(t & 7) << 49;   // matrix base offset field
      const uint64_t db = make_smem_desc_sw128(smem_u32(sW + t * 8192), 16, 1024);
      for (int k = 0; k < 4; ++k) umma_ss<false>(tm, da + uint64_t(k * 2), db + uint64_t(k * 2), idesc, (t | k) != 0);
    }
    umma_commit(mbar);
  }
  mbar_wait(mbar, 0);
  tc_fence_after();
  uint32_t v[32];
  for (int c = 0; c < N; c += 32) {
    tmem_ld32(tm + (uint32_t(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * N + c + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 64);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)p;
  const int L = 512;
  std::vector<float> A(L * K), W(3 * N * K);
  std::vector<bf16> Ab(L * K), Wb(3 * N * K);
  srand(1);
  for (size_t i = 0; i < A.size(); ++i) { Ab[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); A[i] = __bfloat162float(Ab[i]); }
  for (size_t i = 0; i < W.size(); ++i) { Wb[i] = __float2bfloat16((rand() % 2001 - 1000) / 8000.f); W[i] = __bfloat162float(Wb[i]); }
  bf16 *dA, *dW; float* dO;
  cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dW, W.size() * 2); cudaMalloc(&dO, 128 * N * 4);
  cudaMemcpy(dA, Ab.data(), A.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dW, Wb.data(), W.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap tmA, tmW;
  {
    cuuint64_t dims[2] = {K, (cuuint64_t)L}; cuuint64_t str[1] = {K * 2}; cuuint32_t box[2] = {64, ROWS}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode A: %d\n", (int)r);
  }
  {
    cuuint64_t dims[2] = {K, 3 * N}; cuuint64_t str[1] = {K * 2}; cuuint32_t box[2] = {64, N}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dW, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode W: %d\n", (int)r);
  }
  const int smem = 18432 + 3 * 8192 + 64;
  cudaFuncSetAttribute(k_shift, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int l0 : {64, 0, 384}) {     // interior, left edge (row -1 = TMA zero fill), right edge (rows >= L zero fill)
    for (int mode = 0; mode < 2; ++mode) {
      cudaMemset(dO, 0, 128 * N * 4);
      k_shift<<<1, 128, smem>>>(tmA, tmW, dO, l0, mode);
      cudaError_t e = cudaDeviceSynchronize();
      std::vector<float> O(128 * N);
      cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0, maxref = 0;
      for (int l = 0; l < 128; ++l)
        for (int n = 0; n < N; ++n) {
          double acc = 0;
          for (int t = 0; t < 3; ++t) {
            const int ll = l0 + l + t - 1;
            if (ll < 0 || ll >= L) continue;
            for (int k = 0; k < K; ++k) acc += (double)A[ll * K + k] * W[(t * N + n) * K + k];
          }
          maxerr = fmax(maxerr, fabs(acc - O[l * N + n]));
          maxref = fmax(maxref, fabs(acc));
        }
      printf("l0=%d mode=%d (%s): cuda=%s max|err|=%.3e max|ref|=%.3e -> %s\n", l0, mode, mode ? "base_offset=t" : "base_offset=0",
             cudaGetErrorString(e), maxerr, maxref, maxerr < 1e-3 * maxref ? "OK" : "MISMATCH");
    }
  }
  return 0;
}
