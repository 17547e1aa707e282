// Latency of a SATISFIED mbarrier wait: try_wait (with / without suspend hint) vs test_wait, dependent back-to-back.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k(long long* out) {
  __shared__ uint64_t bar;
  const unsigned a = (unsigned)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(a));
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory");     // phase 0 complete
  }
  __syncthreads();
  const int N = 256;
  unsigned ok, acc = 0;
  long long t0 = clock64();
  for (int i = 0; i < N; ++i) {
    asm volatile("{\n.reg .pred P;\nmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\nselp.b32 %0,1,0,P;\n}\n" : "=r"(ok) : "r"(a + (acc & 0)), "r"(0u), "r"(1000000u) : "memory");
    acc += ok;
  }
  long long t1 = clock64();
  for (int i = 0; i < N; ++i) {
    asm volatile("{\n.reg .pred P;\nmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\nselp.b32 %0,1,0,P;\n}\n" : "=r"(ok) : "r"(a + (acc & 0)), "r"(0u) : "memory");
    acc += ok;
  }
  long long t2 = clock64();
  for (int i = 0; i < N; ++i) {
    asm volatile("{\n.reg .pred P;\nmbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\nselp.b32 %0,1,0,P;\n}\n" : "=r"(ok) : "r"(a + (acc & 0)), "r"(0u) : "memory");
    acc += ok;
  }
  long long t3 = clock64();
  if (threadIdx.x == 0) { out[0] = (t1 - t0) / N; out[1] = (t2 - t1) / N; out[2] = (t3 - t2) / N; out[3] = acc; }
}
int main() {
  long long* d; cudaMalloc(&d, 32);
  for (int threads : {32, 128, 512}) {
    k<<<1, threads>>>(d);
    long long r[4]; cudaMemcpy(r, d, 32, cudaMemcpyDeviceToHost);
    printf("%3d threads: satisfied try_wait+hint %lld cycles, try_wait %lld, test_wait %lld (ok %lld) %s\n", threads, r[0], r[1], r[2], r[3], cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
