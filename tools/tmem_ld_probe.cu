// Probe: register fragment layout of tcgen05.ld.16x256b.x4 (row/column of every register of every thread).
// Writes value = lane * 1000 + column with tcgen05.st.32x32b, reads back with 16x256b.x4 at lane offsets 0 and 16.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(uint32_t* out) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(32u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot;
  const uint32_t lane_off = uint32_t(warp * 32) << 16;
  // 32x32b.x32 store: thread = lane, register c = column c
  uint32_t v[32];
  for (int c = 0; c < 32; ++c) v[c] = (warp * 32 + lane) * 1000 + c;
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(base + lane_off), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]),
        "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]),
        "r"(v[30]), "r"(v[31]) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  for (int half = 0; half < 2; ++half) {
    uint32_t r[16];
    const uint32_t ta = base + lane_off + (uint32_t(half * 16) << 16);
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(ta) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 16; ++i) out[((warp * 2 + half) * 32 + lane) * 16 + i] = r[i];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(32u) : "memory");
}
int main() {
  uint32_t* d; cudaMalloc(&d, 4 * 2 * 32 * 16 * 4);
  probe<<<1, 128>>>(d);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
  static uint32_t h[4 * 2 * 32 * 16];
  cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int w = 0; w < 4; ++w) for (int half = 0; half < 2; ++half) for (int t = 0; t < 32; ++t) for (int i = 0; i < 16; ++i) {
    const uint32_t val = h[((w * 2 + half) * 32 + t) * 16 + i];
    const int row = val / 1000, col = val % 1000;
    const int erow = w * 32 + half * 16 + t / 4 + ((i >> 1) & 1) * 8, ecol = (i >> 2) * 8 + (t % 4) * 2 + (i & 1);
    if (row != erow || col != ecol) { if (bad < 20) printf("w%d h%d t%d r%d: row %d col %d (expected %d %d)\n", w, half, t, i, row, col, erow, ecol); ++bad; }
  }
  printf("mismatches vs mma-accumulator layout hypothesis: %d\n", bad);
  for (int i = 0; i < 16; ++i) printf("t5 r%d = %u\n", i, h[(0 * 32 + 5) * 16 + i]);
  return 0;
}
