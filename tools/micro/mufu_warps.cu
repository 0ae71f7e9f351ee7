// MUFU.EX2 throughput as a function of resident warps per SM sub-partition (one CTA per SM).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float* out, int iters) {
  float a[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) a[i] = -0.001f * (threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 64; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 64; ++i) s += a[i];
  if (s == 123.456f) out[0] = s;
}
int main() {
  float* d; cudaMalloc(&d, 4);
  for (int threads : {128, 256, 512, 1024}) {
    int iters = 2048;
    k<<<148, threads>>>(d, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<<<148, threads>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = 148.0 * threads * iters * 64;
    printf("warps/SMSP=%d: %.3f ms, %.2f ex2/clk/SM (at 1.965 GHz nominal)\n", threads / 128, ms, ops / (ms * 1e-3) / 148 / 1.965e9);
  }
  return 0;
}
