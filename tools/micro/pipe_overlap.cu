// Do MUFU.EX2 and the packed-fp32 / conversion instructions overlap inside one SM sub-partition?  Each loop iteration
// issues 16 independent MUFU.EX2 plus R x 16 independent "other" instructions (FFMA2, F2FP or FMNMX3); prints cycles per
// iteration per warp set.  Perfect overlap: max(16*8, R*16*2) per warp-iteration; serialised: the sum.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int KIND, int R> __global__ void k(long long* cycles, float* out, int iters) {
  float m[16];
  uint64_t a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { m[i] = -0.001f * (threadIdx.x + i); a[i] = (uint64_t)(threadIdx.x + i) * 0x3f8000003f800000ull; }
  const uint64_t b = 0x3f8000013f800001ull, c = 0x3a8000003a800000ull;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(m[i]));
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (KIND == 0) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(b), "l"(c));
        if (KIND == 1) { uint32_t o; asm volatile("{.reg .f32 x, y; mov.b64 {x, y}, %1; cvt.rn.bf16x2.f32 %0, x, y; mov.b64 %1, {%0, y};}" : "=r"(o), "+l"(a[i])); }
        if (KIND == 2) { float o; asm volatile("{.reg .f32 x, y; mov.b64 {x, y}, %1; max.f32 %0, x, y, x; mov.b64 %1, {%0, y};}" : "=f"(o), "+l"(a[i])); }
        if (KIND == 3) { float o; asm volatile("{.reg .f32 x, y; mov.b64 {x, y}, %1; fma.rn.f32 %0, x, y, y; mov.b64 %1, {%0, y};}" : "=f"(o), "+l"(a[i])); }
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
  uint64_t x = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) { s += m[i]; x ^= a[i]; }
  if (s == 123.456f || x == 0x1234567ull) out[0] = s;
  if (blockIdx.x == 0 && threadIdx.x == 0) cycles[0] = t1 - t0;
}

template <int KIND, int R> void run(const char* name, long long* dc, float* d) {
  printf("%-10s x%d per MUFU:", name, R);
  for (int threads : {128, 256, 512}) {
    const int iters = 512;
    k<KIND, R><<<148, threads>>>(dc, d, 4);
    k<KIND, R><<<148, threads>>>(dc, d, iters);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    printf("  %dw/SMSP: %7.1f cyc/iter", threads / 128, (double)c / iters);
  }
  printf("\n");
}

int main() {
  float* d; cudaMalloc(&d, 4);
  long long* dc; cudaMalloc(&dc, 8);
  printf("cycles per loop iteration (16 MUFU.EX2 + R*16 others per warp); MUFU alone = 128 x warps/SMSP\n");
  run<0, 0>("none", dc, d);
  run<0, 1>("FFMA2", dc, d); run<0, 2>("FFMA2", dc, d); run<0, 4>("FFMA2", dc, d);
  run<1, 1>("F2FP", dc, d);  run<1, 2>("F2FP", dc, d);  run<1, 4>("F2FP", dc, d);
  run<2, 1>("FMNMX3", dc, d); run<2, 2>("FMNMX3", dc, d); run<2, 4>("FMNMX3", dc, d);
  run<3, 2>("FFMA", dc, d); run<3, 4>("FFMA", dc, d); run<3, 8>("FFMA", dc, d);
  return 0;
}
