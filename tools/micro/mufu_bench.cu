// Micro-benchmark: MUFU throughput of ex2.approx f32 vs f16x2 vs bf16x2 (results/clk/SM) on the current GPU.
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
template <int MODE>
__global__ void k(float* out, int iters) {
  float a[8];
  uint32_t h[8];
  for (int i = 0; i < 8; ++i) { a[i] = -0.001f * (threadIdx.x + i); h[i] = 0xB800B800u + threadIdx.x + i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 3) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 4) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[i]));
      if (MODE == 5) asm volatile("fma.rn.f16x2 %0, %0, %0, %0;" : "+r"(h[i]));
      if (MODE == 6) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %1;" : "+r"(h[i]) : "f"(a[i]));
    }
  }
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i] + (float)h[i];
  if (s == 123.456f) out[0] = s;
}
template <int MODE> void run(const char* name, int elems) {
  float* d; cudaMalloc(&d, 4);
  int iters = 4096, blocks = 148 * 4, threads = 512;
  k<MODE><<<blocks, threads>>>(d, 16);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); k<MODE><<<blocks, threads>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double ops = (double)blocks * threads * iters * 8 * elems;
  printf("%-18s %8.3f ms  %.1f results/ns  ~%.1f results/clk/SM (at %d MHz nominal)\n", name, ms, ops / ms / 1e6,
         ops / (ms * 1e-3) / 148 / (clk * 1e3), clk / 1000);
}
int main() {
  run<0>("ex2.f32", 1); run<1>("ex2.f16x2", 2); run<2>("ex2.bf16x2", 2); run<3>("tanh.f32", 1);
  run<4>("fma.f32", 1); run<5>("fma.f16x2", 2); run<6>("cvt.bf16x2.f32", 2);
  return 0;
}
