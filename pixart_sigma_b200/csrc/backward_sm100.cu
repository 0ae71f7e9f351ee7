// HBM-bound kernels of the training backward pass (train_scripts/train.py:180-220 runs
// loss.backward() through the block; these replace the autograd kernels of torch's LayerNorm / GELU / mul / add / sum and
// the transposes its matmul backward performs implicitly):
//   transpose_bf16        A[R, C] -> A^T[C, R]   (operands of the dgrad / wgrad GEMMs, which take K-contiguous inputs)
//   gelu_fwd / gelu_bwd   h = gelu_tanh(pre);  dpre = dh * gelu_tanh'(pre)
//   gate_residual_fwd/bwd out = x + gate[b] * y;  dy = dout * gate[b],  dgate[b] += sum_rows dout * y
//   ln_modulate_bwd       dx = LN-backward((1 + scale[b]) * dxn);  dshift[b] += sum dxn;  dscale[b] += sum dxn * xhat
//   colsum_bf16           out[c] += sum_rows a[r, c]           (bias gradients)
//   attn_delta            delta[b, h, i] = sum_d dO[b,i,h,d] * O[b,i,h,d]   (row term of the softmax backward)
// One pass over the data each, 128-bit accesses, fp32 arithmetic; the per-sample / per-column reductions accumulate
// in registers over a block of rows and finish with fp32 atomics.
#include "host_common.cuh"
#include "ptx.cuh"

namespace pxa {

PXA_DEVICE float warp_sum_b(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------- transpose
// 64 x 64 tile through smem (padded rows); reads and writes are 128-byte row segments.  kPairs: 32-bit accesses (needs
// even R, C, ldi, ldo and 4-byte aligned bases); otherwise element-wise accesses (odd row counts such as packed captions).
template <bool kPairs>
__global__ void __launch_bounds__(256) transpose_bf16_kernel(const __nv_bfloat16* __restrict__ in,
                                                             __nv_bfloat16* __restrict__ out, int R, int C, long long ldi,
                                                             long long ldo) {
  __shared__ __nv_bfloat16 tile[64][66];
  const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 32 x 8
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = r0 + ty + 8 * i;
    const int c = c0 + 2 * tx;
    if constexpr (kPairs) {
      uint32_t v = 0u;
      if (r < R && c < C) v = *reinterpret_cast<const uint32_t*>(in + (size_t)r * ldi + c);
      *reinterpret_cast<uint32_t*>(&tile[ty + 8 * i][2 * tx]) = v;
    } else {
      const __nv_bfloat16 z = __float2bfloat16(0.f);
      tile[ty + 8 * i][2 * tx] = (r < R && c < C) ? in[(size_t)r * ldi + c] : z;
      tile[ty + 8 * i][2 * tx + 1] = (r < R && c + 1 < C) ? in[(size_t)r * ldi + c + 1] : z;
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = c0 + ty + 8 * i;          // output row
    const int r = r0 + 2 * tx;              // output column pair
    if (c < C && r < R) {
      if constexpr (kPairs) {
        __nv_bfloat162 v;
        v.x = tile[2 * tx][ty + 8 * i];
        v.y = tile[2 * tx + 1][ty + 8 * i];
        *reinterpret_cast<__nv_bfloat162*>(out + (size_t)c * ldo + r) = v;
      } else {
        out[(size_t)c * ldo + r] = tile[2 * tx][ty + 8 * i];
        if (r + 1 < R) out[(size_t)c * ldo + r + 1] = tile[2 * tx + 1][ty + 8 * i];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------- GELU(tanh)
template <bool kBwd>
__global__ void __launch_bounds__(256) gelu_kernel(const uint4* __restrict__ pre, const uint4* __restrict__ dh,
                                                   uint4* __restrict__ out, long long n8) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const uint4 p = pre[i];
    uint4 g = make_uint4(0u, 0u, 0u, 0u);
    if (kBwd) g = dh[i];
    const uint32_t pw[4] = {p.x, p.y, p.z, p.w};
    const uint32_t gw[4] = {g.x, g.y, g.z, g.w};
    uint32_t ow[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float a = bf16_lo(pw[k]), b = bf16_hi(pw[k]);
      if (kBwd) ow[k] = pack_bf16x2(bf16_lo(gw[k]) * gelu_tanh_grad(a), bf16_hi(gw[k]) * gelu_tanh_grad(b));
      else ow[k] = pack_bf16x2(gelu_tanh(a), gelu_tanh(b));
    }
    out[i] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
  }
}

// ------------------------------------------------------------------------------------------------- gate + residual
// out[r, :] = x[r, :] + gate[b, :] * y[r, :]        (fp32 stream, bf16 branch output; gate NULL -> 1)
__global__ void __launch_bounds__(256) gate_residual_fwd_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ y,
                                                                const float* __restrict__ gate, float* __restrict__ out,
                                                                long long gate_bs, int rows_per_batch, int M, int C) {
  const int c4 = C >> 2;
  const long long total = (long long)M * c4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int row = (int)(i / c4);
    const int col = (int)(i - (long long)row * c4) * 4;
    const size_t off = (size_t)row * C + col;
    const float4 xv = *reinterpret_cast<const float4*>(x + off);
    const uint2 yv = *reinterpret_cast<const uint2*>(y + off);
    float4 g = make_float4(1.f, 1.f, 1.f, 1.f);
    if (gate != nullptr) g = __ldg(reinterpret_cast<const float4*>(gate + (size_t)(row / rows_per_batch) * gate_bs + col));
    float4 o;
    o.x = fmaf(g.x, bf16_lo(yv.x), xv.x);
    o.y = fmaf(g.y, bf16_hi(yv.x), xv.y);
    o.z = fmaf(g.z, bf16_lo(yv.y), xv.z);
    o.w = fmaf(g.w, bf16_hi(yv.y), xv.w);
    *reinterpret_cast<float4*>(out + off) = o;
  }
}

// dy[r, :] = bf16(dout[r, :] * gate[b, :]);  dgate[b, :] += sum_{r in b} dout[r, :] * y[r, :]
// One CTA = kRows consecutive rows x C columns; thread t owns columns 4t..4t+3 (C/4 threads).
constexpr int kGateRows = 32;
__global__ void __launch_bounds__(288) gate_residual_bwd_kernel(const float* __restrict__ dout, const __nv_bfloat16* __restrict__ y,
                                                                const float* __restrict__ gate, __nv_bfloat16* __restrict__ dy,
                                                                float* __restrict__ dgate, long long gate_bs,
                                                                int rows_per_batch, int M, int C) {
  const int col = threadIdx.x * 4;
  if (col >= C) return;
  const int r0 = blockIdx.x * kGateRows;
  const int r1 = min(r0 + kGateRows, M);
  int cur_b = -1;
  float4 g = make_float4(1.f, 1.f, 1.f, 1.f);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int row = r0; row < r1; ++row) {
    const int b = row / rows_per_batch;
    if (b != cur_b) {
      if (cur_b >= 0 && dgate != nullptr) {
        float* dg = dgate + (size_t)cur_b * C + col;
        atomicAdd(dg + 0, acc.x); atomicAdd(dg + 1, acc.y); atomicAdd(dg + 2, acc.z); atomicAdd(dg + 3, acc.w);
      }
      acc = make_float4(0.f, 0.f, 0.f, 0.f);
      cur_b = b;
      if (gate != nullptr) g = __ldg(reinterpret_cast<const float4*>(gate + (size_t)b * gate_bs + col));
    }
    const size_t off = (size_t)row * C + col;
    const float4 d = *reinterpret_cast<const float4*>(dout + off);
    if (dgate != nullptr) {
      const uint2 yv = *reinterpret_cast<const uint2*>(y + off);
      acc.x = fmaf(d.x, bf16_lo(yv.x), acc.x);
      acc.y = fmaf(d.y, bf16_hi(yv.x), acc.y);
      acc.z = fmaf(d.z, bf16_lo(yv.y), acc.z);
      acc.w = fmaf(d.w, bf16_hi(yv.y), acc.w);
    }
    *reinterpret_cast<uint2*>(dy + off) = make_uint2(pack_bf16x2(d.x * g.x, d.y * g.y), pack_bf16x2(d.z * g.z, d.w * g.w));
  }
  if (cur_b >= 0 && dgate != nullptr) {
    float* dg = dgate + (size_t)cur_b * C + col;
    atomicAdd(dg + 0, acc.x); atomicAdd(dg + 1, acc.y); atomicAdd(dg + 2, acc.z); atomicAdd(dg + 3, acc.w);
  }
}

// ------------------------------------------------------------------------------------------------- LN + modulate backward
// xn = xhat * (1 + scale[b]) + shift[b],  xhat = (x - mean) * rstd.   With g = dxn * (1 + scale[b]):
//   dx = rstd * (g - mean_c(g) - xhat * mean_c(g * xhat));  dshift[b] += dxn;  dscale[b] += dxn * xhat.
// One warp walks kLnRows consecutive rows (same layout as the forward kernel: lane owns float4 groups g*32 + lane), the
// per-sample column sums stay in registers and are flushed with atomics when the sample changes / at the end.
constexpr int kLnRows = 16;
template <int kVec>
__global__ void __launch_bounds__(256) ln_modulate_bwd_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ dxn,
                                                              const float* __restrict__ scale, float* __restrict__ dx,
                                                              float* __restrict__ dshift, float* __restrict__ dscale,
                                                              long long mod_bs, int rows_per_batch, int M, float eps,
                                                              const float* __restrict__ add_in) {
  constexpr int C = kVec * 128;
  const int warp_global = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int r0 = warp_global * kLnRows;
  if (r0 >= M) return;
  const int r1 = min(r0 + kLnRows, M);
  float4 a_sh[kVec], a_sc[kVec];
#pragma unroll
  for (int g = 0; g < kVec; ++g) a_sh[g] = a_sc[g] = make_float4(0.f, 0.f, 0.f, 0.f);
  int cur_b = -1;
  auto flush = [&]() {
    if (cur_b < 0) return;
#pragma unroll
    for (int g = 0; g < kVec; ++g) {
      const int col = (g * 32 + lane) * 4;
      float* ps = dshift + (size_t)cur_b * C + col;
      float* pc = dscale + (size_t)cur_b * C + col;
      atomicAdd(ps + 0, a_sh[g].x); atomicAdd(ps + 1, a_sh[g].y); atomicAdd(ps + 2, a_sh[g].z); atomicAdd(ps + 3, a_sh[g].w);
      atomicAdd(pc + 0, a_sc[g].x); atomicAdd(pc + 1, a_sc[g].y); atomicAdd(pc + 2, a_sc[g].z); atomicAdd(pc + 3, a_sc[g].w);
      a_sh[g] = a_sc[g] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  for (int row = r0; row < r1; ++row) {
    const int b = row / rows_per_batch;
    if (b != cur_b) {
      flush();
      cur_b = b;
    }
    float4 v[kVec], d[kVec];
    const float* xr = x + (size_t)row * C;
    const __nv_bfloat16* dr = dxn + (size_t)row * C;
#pragma unroll
    for (int g = 0; g < kVec; ++g) {
      const int col = (g * 32 + lane) * 4;
      v[g] = *reinterpret_cast<const float4*>(xr + col);
      const uint2 u = *reinterpret_cast<const uint2*>(dr + col);
      d[g] = make_float4(bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y));
    }
    float s = 0.f;
#pragma unroll
    for (int g = 0; g < kVec; ++g) s += (v[g].x + v[g].y) + (v[g].z + v[g].w);
    const float mean = warp_sum_b(s) * (1.0f / C);
    float ss = 0.f;
#pragma unroll
    for (int g = 0; g < kVec; ++g) {
      const float a = v[g].x - mean, bq = v[g].y - mean, c = v[g].z - mean, e = v[g].w - mean;
      ss += (a * a + bq * bq) + (c * c + e * e);
    }
    const float rstd = rsqrtf(warp_sum_b(ss) * (1.0f / C) + eps);
    const float* sc = scale + (size_t)b * mod_bs;
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int g = 0; g < kVec; ++g) {
      const int col = (g * 32 + lane) * 4;
      const float4 a = __ldg(reinterpret_cast<const float4*>(sc + col));
      // v <- xhat, d <- g = dxn * (1 + scale); column sums use dxn
      const float4 xh = make_float4((v[g].x - mean) * rstd, (v[g].y - mean) * rstd, (v[g].z - mean) * rstd, (v[g].w - mean) * rstd);
      a_sh[g].x += d[g].x; a_sh[g].y += d[g].y; a_sh[g].z += d[g].z; a_sh[g].w += d[g].w;
      a_sc[g].x = fmaf(d[g].x, xh.x, a_sc[g].x); a_sc[g].y = fmaf(d[g].y, xh.y, a_sc[g].y);
      a_sc[g].z = fmaf(d[g].z, xh.z, a_sc[g].z); a_sc[g].w = fmaf(d[g].w, xh.w, a_sc[g].w);
      const float4 gg = make_float4(d[g].x * (1.0f + a.x), d[g].y * (1.0f + a.y), d[g].z * (1.0f + a.z), d[g].w * (1.0f + a.w));
      sg += (gg.x + gg.y) + (gg.z + gg.w);
      sgx += (gg.x * xh.x + gg.y * xh.y) + (gg.z * xh.z + gg.w * xh.w);
      v[g] = xh;
      d[g] = gg;
    }
    const float mg = warp_sum_b(sg) * (1.0f / C);
    const float mgx = warp_sum_b(sgx) * (1.0f / C);
    float* orow = dx + (size_t)row * C;
#pragma unroll
    for (int g = 0; g < kVec; ++g) {
      const int col = (g * 32 + lane) * 4;
      float4 o;
      o.x = rstd * (d[g].x - mg - v[g].x * mgx);
      o.y = rstd * (d[g].y - mg - v[g].y * mgx);
      o.z = rstd * (d[g].z - mg - v[g].z * mgx);
      o.w = rstd * (d[g].w - mg - v[g].w * mgx);
      if (add_in != nullptr) {
        const float4 r = *reinterpret_cast<const float4*>(add_in + (size_t)row * C + col);
        o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
      }
      *reinterpret_cast<float4*>(orow + col) = o;
    }
  }
  flush();
}

// ------------------------------------------------------------------------------------------------- column sums
// out[c] += sum_r a[r, c].  CTA = 32 column-threads (8 columns = one 16-byte load each -> 256 columns, 512 contiguous
// bytes per row) x 8 row lanes; each row lane walks its rows of a 256-row slab with 4 independent loads in flight, the 8
// partial sums meet in smem and leave as one fp32 atomic per column per CTA.  (The first version had one thread walk 256
// rows with a single load in flight on 128 CTAs: 97 us per call, 11 % of the training step in profiles/c5_r1f.)
constexpr int kColsumRows = 128;
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ a, float* __restrict__ out, int M,
                                                          int N, long long lda) {
  __shared__ float part[8][256 + 8];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + tx) * 8;
  const int r0 = blockIdx.y * kColsumRows;
  const int r1 = min(r0 + kColsumRows, M);
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  if (col < N) {
    const __nv_bfloat16* base = a + col;
    int r = r0 + ty;
    for (; r + 24 < r1; r += 32) {                         // 4 independent 16-byte loads in flight per thread
      uint4 u[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) u[j] = *reinterpret_cast<const uint4*>(base + (size_t)(r + 8 * j) * lda);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[0] += bf16_lo(u[j].x); acc[1] += bf16_hi(u[j].x); acc[2] += bf16_lo(u[j].y); acc[3] += bf16_hi(u[j].y);
        acc[4] += bf16_lo(u[j].z); acc[5] += bf16_hi(u[j].z); acc[6] += bf16_lo(u[j].w); acc[7] += bf16_hi(u[j].w);
      }
    }
    for (; r < r1; r += 8) {
      const uint4 u = *reinterpret_cast<const uint4*>(base + (size_t)r * lda);
      acc[0] += bf16_lo(u.x); acc[1] += bf16_hi(u.x); acc[2] += bf16_lo(u.y); acc[3] += bf16_hi(u.y);
      acc[4] += bf16_lo(u.z); acc[5] += bf16_hi(u.z); acc[6] += bf16_lo(u.w); acc[7] += bf16_hi(u.w);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) part[ty][tx * 8 + i] = acc[i];
  __syncthreads();
  const int c = threadIdx.x;                               // 256 threads <-> the CTA's 256 columns
  const int gcol = blockIdx.x * 256 + c;
  if (gcol < N) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += part[j][c];
    atomicAdd(out + gcol, s);
  }
}

// ------------------------------------------------------------------------------------------------- attention delta
// delta[(b*H + h)*Nq + i] = sum_d dO[b,i,h,d] * O[b,i,h,d]; one thread per (row, head): 9 x 16-byte loads from each.
__global__ void __launch_bounds__(256) attn_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o,
                                                         float* __restrict__ delta, int B, int H, int Nq, long long ldo,
                                                         long long lddo) {
  const long long total = (long long)B * Nq * H;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int h = (int)(i % H);
  const long long row = i / H;               // b*Nq + q
  const uint4* po = reinterpret_cast<const uint4*>(o + (size_t)row * ldo + h * 72);
  const uint4* pd = reinterpret_cast<const uint4*>(d_o + (size_t)row * lddo + h * 72);
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < 9; ++c) {
    const uint4 a = po[c], g = pd[c];
    s = fmaf(bf16_lo(a.x), bf16_lo(g.x), s); s = fmaf(bf16_hi(a.x), bf16_hi(g.x), s);
    s = fmaf(bf16_lo(a.y), bf16_lo(g.y), s); s = fmaf(bf16_hi(a.y), bf16_hi(g.y), s);
    s = fmaf(bf16_lo(a.z), bf16_lo(g.z), s); s = fmaf(bf16_hi(a.z), bf16_hi(g.z), s);
    s = fmaf(bf16_lo(a.w), bf16_lo(g.w), s); s = fmaf(bf16_hi(a.w), bf16_hi(g.w), s);
  }
  const int b = (int)(row / Nq), q = (int)(row % Nq);
  delta[((size_t)b * H + h) * Nq + q] = s;
}

// ------------------------------------------------------------------------------------------------- KV compress backward
// Backward of kv_compress_kernel (depthwise 2x2 / stride-2 conv + bias, LayerNorm(affine)) for K and V in one launch
// (blockIdx.y).  One warp per OUTPUT token (same lane <-> channel mapping as the forward); the conv output is recomputed
// from the 4 input rows into a per-warp smem row (keeps the register footprint small: the channel loops stay rolled).
// Input gradients go straight to the 4 input rows (stride == kernel: every input token belongs to exactly one output
// token).  The parameter gradients (conv weight [C,4], conv bias, LN weight, LN bias -- shared by K and V) are accumulated
// per CTA in shared memory and flushed with one fp32 atomic per element per CTA.
constexpr int kKvcRows = 4;     // output tokens per warp
constexpr int kKvcSmem = (7 + 8) * 1152 * 4;
PXA_DEVICE void conv_taps(const __nv_bfloat16* conv_w, int col, float (&wt)[4][4]) {
  const uint4 w01 = __ldg(reinterpret_cast<const uint4*>(conv_w + col * 4));
  const uint4 w23 = __ldg(reinterpret_cast<const uint4*>(conv_w + col * 4 + 8));
  wt[0][0] = bf16_lo(w01.x); wt[0][1] = bf16_hi(w01.x); wt[0][2] = bf16_lo(w01.y); wt[0][3] = bf16_hi(w01.y);
  wt[1][0] = bf16_lo(w01.z); wt[1][1] = bf16_hi(w01.z); wt[1][2] = bf16_lo(w01.w); wt[1][3] = bf16_hi(w01.w);
  wt[2][0] = bf16_lo(w23.x); wt[2][1] = bf16_hi(w23.x); wt[2][2] = bf16_lo(w23.y); wt[2][3] = bf16_hi(w23.y);
  wt[3][0] = bf16_lo(w23.z); wt[3][1] = bf16_hi(w23.z); wt[3][2] = bf16_lo(w23.w); wt[3][3] = bf16_hi(w23.w);
}
template <int kVec>
__global__ void __launch_bounds__(256) kv_compress_bwd_kernel(const __nv_bfloat16* __restrict__ k_in, const __nv_bfloat16* __restrict__ v_in,
                                                              const __nv_bfloat16* __restrict__ dk_out, const __nv_bfloat16* __restrict__ dv_out,
                                                              __nv_bfloat16* __restrict__ dk_in, __nv_bfloat16* __restrict__ dv_in,
                                                              const __nv_bfloat16* __restrict__ conv_w, const __nv_bfloat16* __restrict__ conv_b,
                                                              const __nv_bfloat16* __restrict__ ln_w, float* __restrict__ d_conv_w,
                                                              float* __restrict__ d_conv_b, float* __restrict__ d_ln_w,
                                                              float* __restrict__ d_ln_b, int B, int H, int W, int ld_in, int ld_din,
                                                              float eps) {
  constexpr int C = kVec * 128;
  extern __shared__ float kvc_smem[];
  float* acc = kvc_smem;        // [0,4C) conv weight (c*4 + t), [4C,5C) conv bias, [5C,6C) LN weight, [6C,7C) LN bias
  float* wa = kvc_smem + 7 * C + (threadIdx.x >> 5) * C;     // this warp's conv-output row
  for (int i = threadIdx.x; i < 7 * C; i += blockDim.x) acc[i] = 0.f;
  __syncthreads();
  const int Ho = H >> 1, Wo = W >> 1;
  const int n_out = B * Ho * Wo;
  const int lane = threadIdx.x & 31;
  const __nv_bfloat16* in = blockIdx.y == 0 ? k_in : v_in;
  const __nv_bfloat16* dout = blockIdx.y == 0 ? dk_out : dv_out;
  __nv_bfloat16* din = blockIdx.y == 0 ? dk_in : dv_in;
  const int o0 = (blockIdx.x * 8 + (threadIdx.x >> 5)) * kKvcRows;
  for (int orow = o0; orow < min(o0 + kKvcRows, n_out); ++orow) {
    const int b = orow / (Ho * Wo);
    const int p = orow % (Ho * Wo);
    const int py = p / Wo, px = p % Wo;
    const size_t tok0 = (size_t)b * H * W + (size_t)(2 * py) * W + 2 * px;
    float sm = 0.f;
#pragma unroll 1
    for (int g = 0; g < kVec; ++g) {                        // conv output -> smem row
      const int col = (g * 32 + lane) * 4;
      float wt[4][4];
      conv_taps(conv_w, col, wt);
      const uint2 bb = __ldg(reinterpret_cast<const uint2*>(conv_b + col));
      float4 s = make_float4(bf16_lo(bb.x), bf16_hi(bb.x), bf16_lo(bb.y), bf16_hi(bb.y));
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const uint2 u = *reinterpret_cast<const uint2*>(in + (tok0 + (size_t)(t >> 1) * W + (t & 1)) * ld_in + col);
        s.x = fmaf(wt[0][t], bf16_lo(u.x), s.x);
        s.y = fmaf(wt[1][t], bf16_hi(u.x), s.y);
        s.z = fmaf(wt[2][t], bf16_lo(u.y), s.z);
        s.w = fmaf(wt[3][t], bf16_hi(u.y), s.w);
      }
      *reinterpret_cast<float4*>(wa + col) = s;
      sm += (s.x + s.y) + (s.z + s.w);
    }
    const float mean = warp_sum_b(sm) * (1.0f / C);
    float ss = 0.f;
#pragma unroll 1
    for (int g = 0; g < kVec; ++g) {
      const float4 s = *reinterpret_cast<const float4*>(wa + (g * 32 + lane) * 4);
      const float e0 = s.x - mean, e1 = s.y - mean, e2 = s.z - mean, e3 = s.w - mean;
      ss += (e0 * e0 + e1 * e1) + (e2 * e2 + e3 * e3);
    }
    const float rstd = rsqrtf(warp_sum_b(ss) * (1.0f / C) + eps);
    float sg = 0.f, sgx = 0.f;
#pragma unroll 1
    for (int g = 0; g < kVec; ++g) {                        // LN parameter gradients, row statistics of g = dy * gamma
      const int col = (g * 32 + lane) * 4;
      const float4 s = *reinterpret_cast<const float4*>(wa + col);
      const uint2 du = *reinterpret_cast<const uint2*>(dout + (size_t)orow * C + col);
      const float4 dy = make_float4(bf16_lo(du.x), bf16_hi(du.x), bf16_lo(du.y), bf16_hi(du.y));
      const uint2 gw = __ldg(reinterpret_cast<const uint2*>(ln_w + col));
      const float4 xh = make_float4((s.x - mean) * rstd, (s.y - mean) * rstd, (s.z - mean) * rstd, (s.w - mean) * rstd);
      atomicAdd(&acc[5 * C + col + 0], dy.x * xh.x); atomicAdd(&acc[5 * C + col + 1], dy.y * xh.y);
      atomicAdd(&acc[5 * C + col + 2], dy.z * xh.z); atomicAdd(&acc[5 * C + col + 3], dy.w * xh.w);
      atomicAdd(&acc[6 * C + col + 0], dy.x); atomicAdd(&acc[6 * C + col + 1], dy.y);
      atomicAdd(&acc[6 * C + col + 2], dy.z); atomicAdd(&acc[6 * C + col + 3], dy.w);
      const float4 gg = make_float4(dy.x * bf16_lo(gw.x), dy.y * bf16_hi(gw.x), dy.z * bf16_lo(gw.y), dy.w * bf16_hi(gw.y));
      sg += (gg.x + gg.y) + (gg.z + gg.w);
      sgx += (gg.x * xh.x + gg.y * xh.y) + (gg.z * xh.z + gg.w * xh.w);
    }
    const float mg = warp_sum_b(sg) * (1.0f / C), mgx = warp_sum_b(sgx) * (1.0f / C);
#pragma unroll 1
    for (int g = 0; g < kVec; ++g) {                        // gradient of the conv output -> conv parameters, input rows
      const int col = (g * 32 + lane) * 4;
      const float4 s = *reinterpret_cast<const float4*>(wa + col);
      const uint2 du = *reinterpret_cast<const uint2*>(dout + (size_t)orow * C + col);
      const uint2 gw = __ldg(reinterpret_cast<const uint2*>(ln_w + col));
      const float4 gg = make_float4(bf16_lo(du.x) * bf16_lo(gw.x), bf16_hi(du.x) * bf16_hi(gw.x), bf16_lo(du.y) * bf16_lo(gw.y),
                                    bf16_hi(du.y) * bf16_hi(gw.y));
      const float4 da = make_float4(rstd * (gg.x - mg - (s.x - mean) * rstd * mgx), rstd * (gg.y - mg - (s.y - mean) * rstd * mgx),
                                    rstd * (gg.z - mg - (s.z - mean) * rstd * mgx), rstd * (gg.w - mg - (s.w - mean) * rstd * mgx));
      atomicAdd(&acc[4 * C + col + 0], da.x); atomicAdd(&acc[4 * C + col + 1], da.y);
      atomicAdd(&acc[4 * C + col + 2], da.z); atomicAdd(&acc[4 * C + col + 3], da.w);
      float wt[4][4];
      conv_taps(conv_w, col, wt);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const size_t off = (tok0 + (size_t)(t >> 1) * W + (t & 1));
        const uint2 u = *reinterpret_cast<const uint2*>(in + off * ld_in + col);                  // L1 / L2 hit
        atomicAdd(&acc[(col + 0) * 4 + t], da.x * bf16_lo(u.x)); atomicAdd(&acc[(col + 1) * 4 + t], da.y * bf16_hi(u.x));
        atomicAdd(&acc[(col + 2) * 4 + t], da.z * bf16_lo(u.y)); atomicAdd(&acc[(col + 3) * 4 + t], da.w * bf16_hi(u.y));
        *reinterpret_cast<uint2*>(din + off * ld_din + col) =
            make_uint2(pack_bf16x2(da.x * wt[0][t], da.y * wt[1][t]), pack_bf16x2(da.z * wt[2][t], da.w * wt[3][t]));
      }
    }
    __syncwarp();
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 7 * C; i += blockDim.x) {
    const float v = acc[i];
    if (v != 0.f) {
      float* dst = i < 4 * C ? d_conv_w + i : (i < 5 * C ? d_conv_b + (i - 4 * C) : (i < 6 * C ? d_ln_w + (i - 5 * C) : d_ln_b + (i - 6 * C)));
      atomicAdd(dst, v);
    }
  }
}

static inline int grid_for(long long work_items, int threads) {
  long long g = (work_items + threads - 1) / threads;
  const long long cap = (long long)device_info().sms * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace pxa

#define PXA_ALIGNED16(p) ((reinterpret_cast<uintptr_t>(p) & 15) == 0)

extern "C" int pxa_transpose_bf16(const void* in, void* out, int32_t R, int32_t C, int64_t ldi, int64_t ldo, void* stream) {
  using namespace pxa;
  if (!in || !out) return fail(PXA_ERR_ARG, "null pointer");
  if (R <= 0 || C <= 0) return fail(PXA_ERR_ARG, "bad shape R=%d C=%d", R, C);
  PXA_REQUIRE_SM100();
  dim3 grid((C + 63) / 64, (R + 63) / 64);
  const bool pairs = !((R & 1) || (C & 1) || (ldi & 1) || (ldo & 1) ||
                       ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 3));
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const __nv_bfloat16* pi = reinterpret_cast<const __nv_bfloat16*>(in);
  __nv_bfloat16* po = reinterpret_cast<__nv_bfloat16*>(out);
  if (pairs) transpose_bf16_kernel<true><<<grid, 256, 0, s>>>(pi, po, R, C, ldi, ldo);
  else transpose_bf16_kernel<false><<<grid, 256, 0, s>>>(pi, po, R, C, ldi, ldo);
  launch_counter()++;
  PXA_CHECK_CUDA(cudaGetLastError());
  return PXA_OK;
}

extern "C" int pxa_gelu_tanh_bf16(const void* pre, const void* dh, void* out, int64_t n, void* stream) {
  using namespace pxa;
  if (!pre || !out) return fail(PXA_ERR_ARG, "null pointer");
  if (n <= 0 || (n & 7)) return fail(PXA_ERR_ARG, "n must be a positive multiple of 8 (got %lld)", (long long)n);
  if (!PXA_ALIGNED16(pre) || !PXA_ALIGNED16(out) || !PXA_ALIGNED16(dh)) return fail(PXA_ERR_ALIGN, "pointers must be 16-byte aligned");
  PXA_REQUIRE_SM100();
  const long long n8 = n / 8;
  const int grid = grid_for(n8, 256);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (dh)
    gelu_kernel<true><<<grid, 256, 0, s>>>(reinterpret_cast<const uint4*>(pre), reinterpret_cast<const uint4*>(dh),
                                           reinterpret_cast<uint4*>(out), n8);
  else
    gelu_kernel<false><<<grid, 256, 0, s>>>(reinterpret_cast<const uint4*>(pre), nullptr, reinterpret_cast<uint4*>(out), n8);
  launch_counter()++;
  PXA_CHECK_CUDA(cudaGetLastError());
  return PXA_OK;
}

extern "C" int pxa_gate_residual_fwd(const PxaGateResidualArgs* args, void* stream) {
  using namespace pxa;
  if (!args) return fail(PXA_ERR_ARG, "null args");
  const PxaGateResidualArgs& a = *args;
  if (!a.x || !a.y || !a.out) return fail(PXA_ERR_ARG, "null pointer");
  if (a.M <= 0 || a.C <= 0 || (a.C & 3) || a.rows_per_batch <= 0) return fail(PXA_ERR_ARG, "bad M / C / rows_per_batch");
  if (!PXA_ALIGNED16(a.x) || !PXA_ALIGNED16(a.y) || !PXA_ALIGNED16(a.out) || !PXA_ALIGNED16(a.gate) || (a.gate_batch_stride & 3))
    return fail(PXA_ERR_ALIGN, "pointers must be 16-byte aligned, gate_batch_stride a multiple of 4");
  PXA_REQUIRE_SM100();
  const int grid = grid_for((long long)a.M * (a.C / 4), 256);
  gate_residual_fwd_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float*>(a.x), reinterpret_cast<const __nv_bfloat16*>(a.y), a.gate, reinterpret_cast<float*>(a.out),
      a.gate_batch_stride, a.rows_per_batch, a.M, a.C);
  launch_counter()++;
  PXA_CHECK_CUDA(cudaGetLastError());
  return PXA_OK;
}

extern "C" int pxa_gate_residual_bwd(const PxaGateResidualArgs* args, void* stream) {
  using namespace pxa;
  if (!args) return fail(PXA_ERR_ARG, "null args");
  const PxaGateResidualArgs& a = *args;
  // x = dout (fp32), out = dy (bf16); y / dgate only when a gate gradient is wanted
  if (!a.x || !a.out) return fail(PXA_ERR_ARG, "null pointer");
  if (a.dgate && !a.y) return fail(PXA_ERR_ARG, "dgate needs y");
  if (a.M <= 0 || a.C <= 0 || (a.C & 3) || a.C > 288 * 4 || a.rows_per_batch <= 0) return fail(PXA_ERR_ARG, "bad M / C (<= 1152) / rows_per_batch");
  if (!PXA_ALIGNED16(a.x) || !PXA_ALIGNED16(a.y) || !PXA_ALIGNED16(a.out) || !PXA_ALIGNED16(a.gate) || !PXA_ALIGNED16(a.dgate) ||
      (a.gate_batch_stride & 3))
    return fail(PXA_ERR_ALIGN, "pointers must be 16-byte aligned, gate_batch_stride a multiple of 4");
  PXA_REQUIRE_SM100();
  const int grid = (a.M + kGateRows - 1) / kGateRows;
  gate_residual_bwd_kernel<<<grid, 288, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float*>(a.x), reinterpret_cast<const __nv_bfloat16*>(a.y), a.gate,
      reinterpret_cast<__nv_bfloat16*>(a.out), a.dgate, a.gate_batch_stride, a.rows_per_batch, a.M, a.C);
  launch_counter()++;
  PXA_CHECK_CUDA(cudaGetLastError());
  return PXA_OK;
}

extern "C" int pxa_ln_modulate_bwd(const PxaLnModBwdArgs* args, void* stream) {
  using namespace pxa;
  if (!args) return fail(PXA_ERR_ARG, "null args");
  const PxaLnModBwdArgs& a = *args;
  if (!a.x || !a.dxn || !a.scale || !a.dx || !a.dshift || !a.dscale) return fail(PXA_ERR_ARG, "null pointer");
  if (a.C != 1152) return fail(PXA_ERR_ARG, "pxa_ln_modulate_bwd is specialised for C=1152 (got %d)", a.C);
  if (a.M <= 0 || a.rows_per_batch <= 0) return fail(PXA_ERR_ARG, "bad M / rows_per_batch");
  if (a.mod_batch_stride & 3) return fail(PXA_ERR_ALIGN, "mod_batch_stride must be a multiple of 4");
  if (!PXA_ALIGNED16(a.x) || !PXA_ALIGNED16(a.dxn) || !PXA_ALIGNED16(a.scale) || !PXA_ALIGNED16(a.dx) || !PXA_ALIGNED16(a.dshift) ||
      !PXA_ALIGNED16(a.dscale))
    return fail(PXA_ERR_ALIGN, "pointers must be 16-byte aligned");
  PXA_REQUIRE_SM100();
  const int warps = (a.M + kLnRows - 1) / kLnRows;
  const int grid = (warps + 7) / 8;
  ln_modulate_bwd_kernel<9><<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float*>(a.x), reinterpret_cast<const __nv_bfloat16*>(a.dxn), a.scale, reinterpret_cast<float*>(a.dx),
      a.dshift, a.dscale, a.mod_batch_stride, a.rows_per_batch, a.M, a.eps, a.add_in);
  launch_counter()++;
  PXA_CHECK_CUDA(cudaGetLastError());
  return PXA_OK;
}

extern "C" int pxa_colsum_bf16(const void* a, float* out, int32_t M, int32_t N, int64_t lda, void* stream) {
  using namespace pxa;
  if (!a || !out) return fail(PXA_ERR_ARG, "null pointer");
  if (M <= 0 || N <= 0 || (N & 7) || (lda & 7)) return fail(PXA_ERR_ARG, "M > 0, N and lda positive multiples of 8 required");
  if (!PXA_ALIGNED16(a) || (reinterpret_cast<uintptr_t>(out) & 3)) return fail(PXA_ERR_ALIGN, "a must be 16-byte aligned");
  PXA_REQUIRE_SM100();
  dim3 grid((N + 255) / 256, (M + kColsumRows - 1) / kColsumRows);
  colsum_bf16_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(a), out, M, N, lda);
  launch_counter()++;
  PXA_CHECK_CUDA(cudaGetLastError());
  return PXA_OK;
}

extern "C" int pxa_attn_delta_d72(const void* o, const void* d_o, float* delta, int32_t B, int32_t H, int32_t Nq, int64_t ldo,
                                  int64_t lddo, void* stream) {
  using namespace pxa;
  if (!o || !d_o || !delta) return fail(PXA_ERR_ARG, "null pointer");
  if (B <= 0 || H <= 0 || Nq <= 0) return fail(PXA_ERR_ARG, "bad B / H / Nq");
  if ((ldo & 7) || (lddo & 7) || !PXA_ALIGNED16(o) || !PXA_ALIGNED16(d_o)) return fail(PXA_ERR_ALIGN, "o / dO rows must be 16-byte aligned");
  PXA_REQUIRE_SM100();
  const long long total = (long long)B * Nq * H;
  const int grid = (int)((total + 255) / 256);
  attn_delta_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(o), reinterpret_cast<const __nv_bfloat16*>(d_o), delta, B, H, Nq, ldo, lddo);
  launch_counter()++;
  PXA_CHECK_CUDA(cudaGetLastError());
  return PXA_OK;
}

extern "C" int pxa_kv_compress_conv2_ln_bwd(const PxaKvCompressBwdArgs* args, void* stream) {
  using namespace pxa;
  if (!args) return fail(PXA_ERR_ARG, "null args");
  const PxaKvCompressBwdArgs& a = *args;
  if (!a.k_in || !a.v_in || !a.dk_out || !a.dv_out || !a.dk_in || !a.dv_in || !a.conv_w || !a.conv_b || !a.ln_w || !a.d_conv_w ||
      !a.d_conv_b || !a.d_ln_w || !a.d_ln_b)
    return fail(PXA_ERR_ARG, "null pointer");
  if (a.C != 1152) return fail(PXA_ERR_ARG, "specialised for C=1152 (got %d)", a.C);
  if (a.B <= 0 || a.H < 2 || a.W < 2 || (a.H & 1) || (a.W & 1)) return fail(PXA_ERR_ARG, "bad B/H/W (H, W even)");
  if ((a.ld_in & 7) || (a.ld_din & 7)) return fail(PXA_ERR_ALIGN, "ld_in / ld_din must be multiples of 8");
  if (!PXA_ALIGNED16(a.k_in) || !PXA_ALIGNED16(a.v_in) || !PXA_ALIGNED16(a.dk_out) || !PXA_ALIGNED16(a.dv_out) ||
      !PXA_ALIGNED16(a.dk_in) || !PXA_ALIGNED16(a.dv_in) || !PXA_ALIGNED16(a.conv_w) || !PXA_ALIGNED16(a.conv_b) || !PXA_ALIGNED16(a.ln_w))
    return fail(PXA_ERR_ALIGN, "pointers must be 16-byte aligned");
  PXA_REQUIRE_SM100();
  const int rows = a.B * (a.H / 2) * (a.W / 2);
  dim3 grid((rows + 8 * kKvcRows - 1) / (8 * kKvcRows), 2);
  PXA_CHECK_CUDA(cudaFuncSetAttribute(kv_compress_bwd_kernel<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, kKvcSmem));
  kv_compress_bwd_kernel<9><<<grid, 256, kKvcSmem, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(a.k_in), reinterpret_cast<const __nv_bfloat16*>(a.v_in),
      reinterpret_cast<const __nv_bfloat16*>(a.dk_out), reinterpret_cast<const __nv_bfloat16*>(a.dv_out),
      reinterpret_cast<__nv_bfloat16*>(a.dk_in), reinterpret_cast<__nv_bfloat16*>(a.dv_in),
      reinterpret_cast<const __nv_bfloat16*>(a.conv_w), reinterpret_cast<const __nv_bfloat16*>(a.conv_b),
      reinterpret_cast<const __nv_bfloat16*>(a.ln_w), a.d_conv_w, a.d_conv_b, a.d_ln_w, a.d_ln_b, a.B, a.H, a.W, a.ld_in, a.ld_din, a.eps);
  launch_counter()++;
  PXA_CHECK_CUDA(cudaGetLastError());
  return PXA_OK;
}
