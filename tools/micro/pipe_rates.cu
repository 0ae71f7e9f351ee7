// Issue rate of the instructions the attention exp2 section is made of, per SM sub-partition, as a function of resident
// warps (one CTA per SM; independent chains, so the number is a throughput, not a latency).  Prints warp-instructions per
// clock per sub-partition, measured with clock64() inside the kernel.
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

template <int OP> __device__ __forceinline__ void op(uint64_t& a, uint64_t b, uint64_t c) {
  if (OP == 0) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a) : "l"(b), "l"(c));
  if (OP == 1) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a) : "l"(b));
  if (OP == 2) { float lo; asm volatile("{.reg .f32 x, y; mov.b64 {x, y}, %1; fma.rn.f32 %0, x, y, y; mov.b64 %1, {%0, y};}" : "=f"(lo), "+l"(a)); }
  if (OP == 3) { float lo; asm volatile("{.reg .f32 x, y; mov.b64 {x, y}, %1; max.f32 %0, x, y, x; mov.b64 %1, {%0, y};}" : "=f"(lo), "+l"(a)); }
  if (OP == 4) { uint32_t r; asm volatile("{.reg .f32 x, y; mov.b64 {x, y}, %1; cvt.rn.bf16x2.f32 %0, x, y; mov.b64 %1, {%0, y};}" : "=r"(r), "+l"(a)); }
  if (OP == 5) { float lo; asm volatile("{.reg .f32 x, y; mov.b64 {x, y}, %1; ex2.approx.ftz.f32 %0, x; mov.b64 %1, {%0, y};}" : "=f"(lo), "+l"(a)); }
  if (OP == 6) { float lo; asm volatile("{.reg .f32 x, y; mov.b64 {x, y}, %1; max.f32 %0, x, y; mov.b64 %1, {%0, y};}" : "=f"(lo), "+l"(a)); }
  if (OP == 7) { uint32_t r; asm volatile("{.reg .b32 x, y; mov.b64 {x, y}, %1; mad.lo.s32 %0, x, 8388608, y; mov.b64 %1, {%0, y};}" : "=r"(r), "+l"(a)); }
}

template <int OP> __global__ void k(long long* cycles, float* out, int iters) {
  uint64_t a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = (uint64_t)(threadIdx.x + i) * 0x3f8000003f800000ull;
  uint64_t b = 0x3f8000013f800001ull, c = 0x3a8000003a800000ull;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) op<OP>(a[i], b, c);
  }
  long long t1 = clock64();
  uint64_t s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s ^= a[i];
  if (s == 0x123456789ull) out[0] = 1.f;
  if (blockIdx.x == 0 && threadIdx.x == 0) cycles[0] = t1 - t0;
}

template <int OP> void run(const char* name, long long* dc, float* d) {
  printf("%-28s", name);
  for (int threads : {128, 256, 512}) {
    int iters = 1024;
    k<OP><<<148, threads>>>(dc, d, 8);
    k<OP><<<148, threads>>>(dc, d, iters);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    double per_smsp = (double)(threads / 128) * iters * 16 / (double)c;
    printf("  %d warp/SMSP: %.3f instr/clk", threads / 128, per_smsp);
  }
  printf("\n");
}

int main() {
  float* d; cudaMalloc(&d, 4);
  long long* dc; cudaMalloc(&dc, 8);
  run<0>("FFMA2 (fma.rn.f32x2)", dc, d);
  run<1>("FADD2 (add.rn.f32x2)", dc, d);
  run<2>("FFMA (scalar)", dc, d);
  run<3>("FMNMX3 (3-input max)", dc, d);
  run<6>("FMNMX (2-input max)", dc, d);
  run<4>("F2FP.BF16 (cvt pack)", dc, d);
  run<5>("MUFU.EX2", dc, d);
  run<7>("IMAD (mad.lo)", dc, d);
  return 0;
}
