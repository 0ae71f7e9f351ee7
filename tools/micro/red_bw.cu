// Micro-benchmark: L2 fp32 reduce-add throughput for the access pattern a single-pass attention backward would need
// (DESIGN.md 4.8): every CTA adds 128 x 72 fp32 partial dQ tiles (rows 288 B long, 4608 B apart = [token][16 heads][72]) into a
// 4 x 4096 x 1152 fp32 accumulator (75 MB), the tiles of one (sample, head) visited by 32 CTAs at the same time.
//   mode 0: cp.reduce.async.bulk (one 288-byte row per instruction, issued by one thread from smem)
//   mode 1: red.global.add.v4.f32 (one thread per row, 18 vector reductions)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_bw red_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kRows = 128, kCols = 72, kHeads = 16, kTokens = 4096, kBatch = 4;

template <int MODE>
__global__ void __launch_bounds__(128) red_kernel(float* acc, int tiles_per_cta) {
  __shared__ __align__(128) float tile[kRows * kCols];
  for (int i = threadIdx.x; i < kRows * kCols; i += 128) tile[i] = 1.0f;
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  __syncthreads();
  // CTA = (key tile j, head h, sample b); walks the query tiles i starting at a CTA-dependent offset (as concurrent key-tile CTAs would)
  const int j = blockIdx.x % 32, h = (blockIdx.x / 32) % kHeads, b = (blockIdx.x / (32 * kHeads)) % kBatch;
  for (int t = 0; t < tiles_per_cta; ++t) {
    const int i = (t + j) % (kTokens / kRows);
    float* base = acc + ((size_t)(b * kTokens + i * kRows) * kHeads + h) * kCols;
    if (MODE == 0) {
      if (threadIdx.x == 0) {
        for (int r = 0; r < kRows; ++r) {
          asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;\n" ::"l"(base + (size_t)r * kHeads * kCols),
                       "r"((uint32_t)__cvta_generic_to_shared(tile + r * kCols)), "r"(kCols * 4)
                       : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
      }
    } else {
      float* row = base + (size_t)threadIdx.x * kHeads * kCols;
#pragma unroll
      for (int c = 0; c < kCols; c += 4)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(row + c), "f"(1.f), "f"(1.f), "f"(1.f), "f"(1.f) : "memory");
    }
  }
  if (MODE == 0 && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
}

int main() {
  float* acc;
  const size_t n = (size_t)kBatch * kTokens * kHeads * kCols;
  cudaMalloc(&acc, n * 4);
  cudaMemset(acc, 0, n * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int tiles = 32;
  for (int mode = 0; mode < 2; ++mode) {
    for (int ctas : {148, 296, 592, 2048}) {
      float best = 1e9f;
      for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) red_kernel<0><<<ctas, 128>>>(acc, tiles); else red_kernel<1><<<ctas, 128>>>(acc, tiles);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
      }
      const double bytes = (double)ctas * tiles * kRows * kCols * 4;
      printf("%s  %4d CTAs x %d tiles of 128x72 fp32: %.3f ms  %.2f TB/s reduced\n", mode == 0 ? "cp.reduce.async.bulk (288 B rows)" : "red.global.add.v4.f32          ",
             ctas, tiles, best, bytes / best / 1e9);
    }
  }
  cudaError_t err = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(err));
  return 0;
}
