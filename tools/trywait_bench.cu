// How long does one mbarrier.try_wait poll park the thread, as a function of the suspend-time hint?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__global__ void k(unsigned hint, int use_hint, int polls, unsigned long long* out) {
  __shared__ uint64_t bar;
  const unsigned a = (unsigned)__cvta_generic_to_shared(&bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(a));
  __syncthreads();
  const unsigned long long t0 = gtimer();
  unsigned ok = 0;
  for (int i = 0; i < polls; ++i) {
    if (use_hint) asm volatile("{\n.reg .pred P;\nmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\nselp.b32 %0,1,0,P;\n}\n" : "=r"(ok) : "r"(a), "r"(0u), "r"(hint) : "memory");
    else asm volatile("{\n.reg .pred P;\nmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\nselp.b32 %0,1,0,P;\n}\n" : "=r"(ok) : "r"(a), "r"(0u) : "memory");
    if (ok) break;
  }
  out[0] = gtimer() - t0; out[1] = ok;
}
int main() {
  unsigned long long* d; cudaMalloc(&d, 16);
  unsigned hints[] = {0, 1000, 100000, 1000000, 10000000, 100000000};
  for (int h = 0; h < 6; ++h) {
    const int polls = h == 0 ? 100000 : 200;
    k<<<1, 32>>>(hints[h], h != 0, polls, d);
    unsigned long long r[2]; cudaMemcpy(r, d, 16, cudaMemcpyDeviceToHost);
    printf("hint %10u ns (%s): %d polls in %.3f ms -> %.1f ns per poll (err %s)\n", hints[h], h ? "given" : "none", polls, r[0] / 1e6, (double)r[0] / polls, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
