// Micro-benchmark: per-SM throughput of the special-function ops the A-transform / softmax warps lean on, as a function
// of warps per SM.   nvcc -arch=sm_100a -O3 -o tools/bin/mufu_bench tools/mufu_bench.cu && tools/bin/mufu_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int OP> __device__ __forceinline__ uint32_t op(uint32_t x) {
  uint32_t y;
  if (OP == 0) asm volatile("tanh.approx.bf16x2 %0, %1;" : "=r"(y) : "r"(x));
  else if (OP == 1) { float f; asm volatile("tanh.approx.f32 %0, %1;" : "=f"(f) : "f"(__uint_as_float(x))); y = __float_as_uint(f); }
  else if (OP == 2) { float f; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(f) : "f"(__uint_as_float(x))); y = __float_as_uint(f); }
  else if (OP == 3) asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x));
  else if (OP == 4) asm volatile("tanh.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
  else if (OP == 5) asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
  else if (OP == 6) { float f; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(f) : "f"(__uint_as_float(x))); y = __float_as_uint(f); }
  else if (OP == 7) asm volatile("fma.rn.bf16x2 %0, %1, %1, %1;" : "=r"(y) : "r"(x));
  else { float f = __uint_as_float(x); f = fmaf(f, f, f); y = __float_as_uint(f); }
  return y;
}

template <int OP>
__global__ void k(uint32_t* out, long long* clk, int iters) {
  uint32_t v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = 0x3c003c00u + threadIdx.x * 16 + i;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = op<OP>(v[i]);
  }
  __syncthreads();
  const long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s ^= v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int OP> void run(const char* name, int elems_per_op) {
  uint32_t* out; long long* clk;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 148 * 8);
  const int iters = 256;
  for (int warps : {4, 8, 16, 32}) {
    k<OP><<<148, warps * 32, 0>>>(out, clk, iters);
    k<OP><<<148, warps * 32, 0>>>(out, clk, iters);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, clk, sizeof h, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    const double ops = (double)warps * 32 * 16 * iters;
    printf("%-22s warps/SM %2d: %8.0f clk  -> %6.2f thread-ops/clk/SM (%6.2f elems/clk/SM), %5.2f clk per warp-instr per SMSP\n", name, warps, avg,
           ops / avg, ops * elems_per_op / avg, avg / (16.0 * iters * warps / 4));
  }
  cudaFree(out); cudaFree(clk);
}

int main() {
  run<0>("tanh.approx.bf16x2", 2);
  run<1>("tanh.approx.f32", 1);
  run<2>("ex2.approx.ftz.f32", 1);
  run<3>("ex2.approx.ftz.bf16x2", 2);
  run<4>("tanh.approx.f16x2", 2);
  run<5>("ex2.approx.f16x2", 2);
  run<6>("rcp.approx.ftz.f32", 1);
  run<7>("fma.rn.bf16x2", 2);
  run<8>("fma.rn.f32", 1);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
