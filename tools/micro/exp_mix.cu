// Softmax exp2 section with a fraction of the exponentials moved from the MUFU pipe to the FMA pipe (round 2).
//
// The MUFU pipe (MUFU.EX2: 8 cycles per warp instruction per SM sub-partition) bounds the attention kernel at head_dim 72.
// Round 1's polynomial exp2 never won because its range reduction ran on the ALU pipe (2 FMNMX + SHL + IADD per element at 2
// cycles each) -- as expensive as the MUFU op it replaced.  This version keeps everything but one LEA on the FMA pipe:
//   z  = fma.rn.sat(s, -scale/128, (m*scale + 8)/128)      in [0, 1]:  x' = -128 z = clamp(x - 8, -128, 0)   (FFMA.SAT: the clamp is free)
//   r  = fma.rm(z, -128, magic)                            magic = 1.5 * 2^23 + 8: floor(x') + 8 lands in the low mantissa bits
//   f  = fma.rn(z, -128, -(r - magic))                     fractional part in [0, 1), exact
//   p  = ((c3 f + c2) f + c1) f + c0                       2^f, degree 3 (max rel. error 8.8e-5, far below bf16 rounding of P)
//   e  = p + (r << 23)  as integers (LEA)                  exponent insertion: e = 2^(x' + 8) = 2^x
// Measures cycles per 64-element row block for 1 / 2 warps per sub-partition and the poly share 0/8 .. 5/8, scalar and packed
// (f32x2) forms, with the row max (32 FMNMX3) and the bf16 packing (32 F2FP) of the real kernel in the loop; and the accuracy
// of the polynomial path against exp2f.
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <cuda_runtime.h>
#include "../../pixart_sigma_b200/csrc/ptx.cuh"
using namespace pxa;

// fma_sat / fma2_rm: ptx.cuh (the attention kernel uses the packed form, pxa::fma_exp2_x2)
__device__ __forceinline__ float fma_rm(float a, float b, float c) { float r; asm("fma.rm.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

constexpr float kMagic = 12582912.0f + 8.0f;     // 1.5 * 2^23 + 8
constexpr float kC0 = 1.0f, kC1 = 0.695146143436431884765625f, kC2 = 0.227564394474029541015625f, kC3 = 0.077119089663028717041015625f;

// scalar polynomial exp2 of x = s * sl2 + nm (x <= 8): 1 FFMA.SAT + 6 FMA-pipe ops + 1 LEA
__device__ __forceinline__ float poly_exp2_sat(float s, float a_sat, float b_sat) {
  const float z = fma_sat(s, a_sat, b_sat);
  const float r = fma_rm(z, -128.0f, kMagic);
  const float f = fmaf(z, -128.0f, kMagic - r);            // (kMagic - r) is exact
  const float p = fmaf(fmaf(fmaf(kC3, f, kC2), f, kC1), f, kC0);
  return __int_as_float(__float_as_int(p) + (__float_as_int(r) << 23));
}
// packed pair: 2 FFMA.SAT + 6 FFMA2/FADD2 + 2 LEA
__device__ __forceinline__ void poly_exp2_sat_x2(float s0, float s1, float a_sat, float b_sat, float& e0, float& e1) {
  const uint64_t z = f32x2(fma_sat(s0, a_sat, b_sat), fma_sat(s1, a_sat, b_sat));
  const uint64_t m128 = f32x2(-128.0f, -128.0f), mg = f32x2(kMagic, kMagic);
  const uint64_t r = fma2_rm(z, m128, mg);
  const uint64_t f = fma2(z, m128, sub2(mg, r));
  uint64_t p = fma2(f32x2(kC3, kC3), f, f32x2(kC2, kC2));
  p = fma2(p, f, f32x2(kC1, kC1));
  p = fma2(p, f, f32x2(kC0, kC0));
  float p0, p1, r0, r1;
  f32x2_split(p, p0, p1);
  f32x2_split(r, r0, r1);
  e0 = __int_as_float(__float_as_int(p0) + (__float_as_int(r0) << 23));
  e1 = __int_as_float(__float_as_int(p1) + (__float_as_int(r1) << 23));
}

// POLY = pairs out of every 8 that take the polynomial; PACKED: f32x2 form of the polynomial
template <int POLY, bool PACKED> __global__ void k(long long* cycles, uint32_t* out, int iters, float sl2, float m0) {
  float v[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) v[i] = -0.05f * ((threadIdx.x * 7 + i * 13) % 97);
  uint32_t acc = 0;
  float m_ref = m0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      mx0 = fmax3(mx0, v[4 * i], v[4 * i + 1]);
      mx1 = fmax3(mx1, v[4 * i + 2], v[4 * i + 3]);
    }
    m_ref = fmaxf(m_ref, fmaxf(mx0, mx1)) * 0.999f;
    const float nm = -m_ref * sl2;
    const uint64_t sl2x2 = f32x2(sl2, sl2), nm2 = f32x2(nm, nm);
    const float a_sat = -sl2 * (1.0f / 128.0f), b_sat = (8.0f - nm) * (1.0f / 128.0f);
    uint32_t pk[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      float e0, e1;
      if (((i * POLY) & 7) < POLY && POLY > 0) {
        if (PACKED) poly_exp2_sat_x2(v[2 * i], v[2 * i + 1], a_sat, b_sat, e0, e1);
        else { e0 = poly_exp2_sat(v[2 * i], a_sat, b_sat); e1 = poly_exp2_sat(v[2 * i + 1], a_sat, b_sat); }
      } else {
        const uint64_t x = fma2(f32x2(v[2 * i], v[2 * i + 1]), sl2x2, nm2);
        float x0, x1;
        f32x2_split(x, x0, x1);
        e0 = fast_exp2(x0);
        e1 = fast_exp2(x1);
      }
      pk[i] = pack_bf16x2(e0, e1);
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) acc ^= pk[i];
    v[it & 63] += __uint_as_float((acc & 1) << 20);           // loop-carried dependency: iterations cannot overlap
  }
  long long t1 = clock64();
  if (acc == 0x12345u) out[0] = acc;
  if (blockIdx.x == 0 && threadIdx.x == 0) cycles[0] = t1 - t0;
}

template <int POLY, bool PACKED> void run(long long* dc, uint32_t* d) {
  printf("poly %d/8 %s", POLY, PACKED ? "packed" : "scalar");
  for (int threads : {128, 256, 384}) {
    const int iters = 512;
    k<POLY, PACKED><<<148, threads>>>(dc, d, 4, 0.17f, -3.f);
    k<POLY, PACKED><<<148, threads>>>(dc, d, iters, 0.17f, -3.f);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    printf("  %dw/SMSP: %7.1f cyc/iter", threads / 128, (double)c / iters);
  }
  printf("\n");
}

__global__ void acc_k(float* err, int n) {
  // relative error of the polynomial path vs exp2f over its clamp range x' = x - 8 in [-128, 0]
  float worst = 0.f, worst_mufu = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float x = -119.9f + 127.9f * (float)i / (float)(n - 1);      // [-119.9, 8]: the clamp range (below it P ~ 2^-120 ~ 0)
    // s * sl2 + nm = x with sl2 = 1, nm = 0
    const float e = poly_exp2_sat(x, -1.0f / 128.0f, 8.0f / 128.0f);
    const float ref = exp2f(x);
    worst = fmaxf(worst, fabsf(e - ref) / ref);
    worst_mufu = fmaxf(worst_mufu, fabsf(fast_exp2(x) - ref) / ref);
  }
  atomicMax(reinterpret_cast<int*>(err), __float_as_int(worst));
  atomicMax(reinterpret_cast<int*>(err) + 1, __float_as_int(worst_mufu));
  if (threadIdx.x == 0) {      // clamp behaviour at the edges
    err[2] = poly_exp2_sat(-500.f, -1.0f / 128.0f, 8.0f / 128.0f);
    err[3] = poly_exp2_sat(-INFINITY, -1.0f / 128.0f, 8.0f / 128.0f);
    err[4] = poly_exp2_sat(8.0f, -1.0f / 128.0f, 8.0f / 128.0f);
    err[5] = poly_exp2_sat(0.0f, -1.0f / 128.0f, 8.0f / 128.0f);
  }
}

int main() {
  uint32_t* d; cudaMalloc(&d, 4);
  long long* dc; cudaMalloc(&dc, 8);
  float* de; cudaMalloc(&de, 32); cudaMemset(de, 0, 32);
  acc_k<<<1, 256>>>(de, 1 << 20);
  float he[8]; cudaMemcpy(he, de, 32, cudaMemcpyDeviceToHost);
  printf("max rel err vs exp2f on [-119.9, 8]: polynomial %.3e, MUFU ex2.approx %.3e; f(-500)=%g f(-inf)=%g f(8)=%g f(0)=%g\n",
         he[0], he[1], he[2], he[3], he[4], he[5]);
  run<0, false>(dc, d);
  run<1, false>(dc, d); run<2, false>(dc, d); run<3, false>(dc, d); run<4, false>(dc, d); run<5, false>(dc, d);
  run<1, true>(dc, d); run<2, true>(dc, d); run<3, true>(dc, d); run<4, true>(dc, d); run<5, true>(dc, d);
  run<8, true>(dc, d);
  return 0;
}
