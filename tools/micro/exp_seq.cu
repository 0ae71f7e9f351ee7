// The attention kernel's per-sub-block exp2 sequence in isolation: 32 FFMA2 (scale) -> 64 MUFU.EX2 -> 32 F2FP (bf16 pack)
// on 64 register-resident scores per thread, looped; cycles per iteration for 1 / 2 / 4 warps per SM sub-partition.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../pixart_sigma_b200/csrc/ptx.cuh"
using namespace pxa;

template <int MODE> __global__ void k(long long* cycles, uint32_t* out, int iters, float sl2, float nm0) {
  float v[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) v[i] = -0.01f * ((threadIdx.x * 7 + i * 13) % 97);
  uint32_t acc = 0;
  float nm = nm0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const uint64_t sl2x2 = f32x2(sl2, sl2);
    const uint64_t nm2 = f32x2(nm, nm);
    uint32_t pk[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const uint64_t x = fma2(f32x2(v[2 * i], v[2 * i + 1]), sl2x2, nm2);
      float x0, x1;
      f32x2_split(x, x0, x1);
      float e0, e1;
      if (MODE == 0) { e0 = fast_exp2(x0); e1 = fast_exp2(x1); }
      else { e0 = x0 * 1.0001f; e1 = x1 * 1.0001f; }          // MODE 1: no MUFU (what is left of the sequence)
      pk[i] = pack_bf16x2(e0, e1);
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) acc ^= pk[i];
    nm = __uint_as_float(__float_as_uint(nm) + (acc & 1));    // loop-carried dependency: iterations cannot overlap
  }
  long long t1 = clock64();
  if (acc == 0x12345u) out[0] = acc;
  if (blockIdx.x == 0 && threadIdx.x == 0) cycles[0] = t1 - t0;
}

template <int MODE> void run(const char* name, long long* dc, uint32_t* d) {
  printf("%-36s", name);
  for (int threads : {128, 256, 512}) {
    const int iters = 256;
    k<MODE><<<148, threads>>>(dc, d, 4, 0.17f, -3.f);
    k<MODE><<<148, threads>>>(dc, d, iters, 0.17f, -3.f);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    printf("  %dw/SMSP: %7.1f cyc/iter", threads / 128, (double)c / iters);
  }
  printf("\n");
}

int main() {
  uint32_t* d; cudaMalloc(&d, 4);
  long long* dc; cudaMalloc(&dc, 8);
  run<0>("32 FFMA2 + 64 EX2 + 32 F2FP + 32 LOP", dc, d);
  run<1>("same without the EX2", dc, d);
  return 0;
}
