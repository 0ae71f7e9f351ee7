// Shared pieces of the 1-CTA and 2-CTA tcgen05 GEMM kernels: parameters, smem/TMEM configuration and the epilogues.
#pragma once
#include "host_common.cuh"
#include "ptx.cuh"

namespace pxa {

constexpr int kBM = 128;
constexpr int kBK = 64;            // 64 bf16 = 128 B = one swizzle atom row
constexpr int kGemmThreads = 256;
constexpr int kEpiWarp0 = 4;       // first epilogue warp
constexpr int kNumEpiThreads = 128;

struct GemmParams {
  const __nv_bfloat16* bias;
  void* out;
  __nv_bfloat16* out_aux;
  const void* residual;
  const float* gate;
  long long gate_batch_stride;
  int rows_per_batch;
  int M, N, K, ldo;
  int num_m_tiles, num_n_tiles;
  long long* trace;     // debug only (NULL in production)
  int k_splits;         // > 1: split-K over tiles (reduce-add epilogue only)
  int aux_branch;       // residual epilogue: the bf16 aux output receives acc + bias (the branch output) instead of out
  int reverse_tiles;    // visit the tiles from the last row block to the first (L2 reuse of a just-written A operand)
  // fused LayerNorm-modulate, producer side (fp32 residual epilogue): aux = out * aux_scale[b], row partial sums of out
  const float* aux_scale;
  long long aux_scale_batch_stride;
  float* row_stats_out;
  // consumer side (PXA_EPI_LN_BIAS / _GELU): out = rstd * (acc - mu * u[b]) + v[b]
  const float* ln_stats;
  const float* ln_u;
  const float* ln_v;
  long long ln_uv_batch_stride;
  float ln_inv_dim, ln_eps;
  // implicit-GEMM 3x3 convolution mode (kConv): A is an NHWC image read through a 4-D tensor map
  int conv_H, conv_W, conv_tile_w, conv_tile_h, conv_cin_blocks;
  // widths that do not tile (neither a power of two <= 128 nor a multiple of 128): the M index runs over a VIRTUAL image of width
  // conv_Wp = W rounded up to 128 -- pixel columns >= W are TMA zero fill on the way in and are not stored on the way out
  int conv_Wp;
};

inline void fill_ln_params(GemmParams& p, const PxaGemmArgs& a) {
  p.aux_scale = a.aux_scale;
  p.aux_scale_batch_stride = a.aux_scale_batch_stride;
  p.row_stats_out = a.row_stats_out;
  p.ln_stats = a.ln_stats;
  p.ln_u = a.ln_u;
  p.ln_v = a.ln_v;
  p.ln_uv_batch_stride = a.ln_uv_batch_stride;
  p.ln_inv_dim = a.ln_dim > 0 ? 1.0f / (float)a.ln_dim : 0.f;
  p.ln_eps = a.ln_eps;
  p.reverse_tiles = a.reverse_tiles ? 1 : 0;
}

constexpr int kResBufs = 3;
constexpr int kResChunkBytes = 128 * 32 * 4;     // 16 KB: 128 rows x 32 fp32 columns
constexpr int kAuxChunkBytes = 128 * 32 * 2;     // 8 KB
constexpr int kResEpiSmem = kResBufs * kResChunkBytes + 2 * kAuxChunkBytes;   // 64 KB
constexpr int kSmemBudget = 227 * 1024;

// Per-tile epilogue constants staged in smem while the tile's MMAs run: bias (fp32) and the gate rows of the (at most
// two) samples a 128-row tile can span.  Reading them per chunk is then an LDS broadcast instead of a chain of L2-latency
// global loads in the epilogue's critical path.  Double-buffered by tile parity.
struct EpiConst {
  float bias[256];     // LN consumer epilogue: v of the first sample
  float gate0[256];    // gate of the sample of the tile's first row (1.0 when there is no gate); LN consumer: v of the last sample
  float gate1[256];    // gate of the sample of the tile's last row
  float ex0[256];      // aux_scale of the first / last sample (1.0 when there is none); LN consumer: u of the first / last sample
  float ex1[256];
  int row_split;       // tile rows >= row_split belong to the second sample
  int pad[3];
};
constexpr int kEpiConstBytes = 2 * sizeof(EpiConst);    // 10272 B

// kTmaRes: fp32 residual epilogue streamed through smem by TMA (see the end of this file)
// kReduceOnly: the weight-gradient form only ever runs the reduce-add epilogue (no residual loads, no aux copy): two chunk
// buffers are enough, which leaves room for one more operand stage (its K loop is thousands of blocks long).
template <int BN, bool kTmaRes = false, bool kReduceOnly = false> struct GemmCfg {
  static constexpr int kStageA = kBM * kBK * 2;                 // 16 KB
  static constexpr int kStageB = BN * kBK * 2;
  static constexpr int kStage = kStageA + kStageB;
  static constexpr int kResRing = kReduceOnly ? 2 : kResBufs;                // chunk buffers of the TMA residual epilogue
  static constexpr int kEpiBufs = kTmaRes ? (kReduceOnly ? 2 * kResChunkBytes : kResEpiSmem)
                                          : 4 * 32 * 32 * 4;                 // else 4 epilogue warps x (32 rows x 128 B)
  static constexpr int kEpiSmem = kEpiBufs + ((kEpiConstBytes + 127) / 128) * 128;
  static constexpr int kBarBytes = 256;
  static constexpr int kMaxStages = (kSmemBudget - kEpiSmem - kBarBytes - 1024) / kStage;
  static constexpr int kStages = kMaxStages > 6 ? 6 : kMaxStages;
  static constexpr int kSmem = kStages * kStage + kEpiSmem + kBarBytes + 1024;  // +1024 alignment slack
  static constexpr uint32_t kTmemCols = (2 * BN <= 256) ? 256 : 512;
};

// ------------------------------------------------------------------------------------------------- epilogues
// Row-per-thread registers `v` hold acc for columns [col0, col0+32) of this thread's row.

// bf16 output (EPI 0/1): math in row-per-thread layout, pack to bf16, transpose through smem (64 B per row).
template <int EPI>
PXA_DEVICE void epilogue_chunk_bf16(uint32_t (&v)[32], const GemmParams& p, uint8_t* stile, int lane, int row0, int col0) {
  // bias: every lane needs the same 32 values -> 4 broadcast 16-byte loads
  float b[32];
  if (p.bias != nullptr) {
    const uint4* bp = reinterpret_cast<const uint4*>(p.bias + col0);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 u = make_uint4(0u, 0u, 0u, 0u);
      if (col0 + 8 * i < p.N) u = __ldg(bp + i);
      b[8 * i + 0] = bf16_lo(u.x); b[8 * i + 1] = bf16_hi(u.x);
      b[8 * i + 2] = bf16_lo(u.y); b[8 * i + 3] = bf16_hi(u.y);
      b[8 * i + 4] = bf16_lo(u.z); b[8 * i + 5] = bf16_hi(u.z);
      b[8 * i + 6] = bf16_lo(u.w); b[8 * i + 7] = bf16_hi(u.w);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) b[i] = 0.f;
  }
  uint32_t pk[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float x0 = __uint_as_float(v[2 * i]) + b[2 * i];
    float x1 = __uint_as_float(v[2 * i + 1]) + b[2 * i + 1];
    if (EPI == PXA_EPI_BIAS_GELU) {
      x0 = gelu_tanh(x0);
      x1 = gelu_tanh(x1);
    }
    pk[i] = pack_bf16x2(x0, x1);
  }
  // smem tile: 32 rows x 64 B; 16-byte chunk c of row r lives at r*64 + ((c ^ ((r >> 1) & 3)) * 16)
  {
    const int sw = (lane >> 1) & 3;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint4 u = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
      *reinterpret_cast<uint4*>(stile + lane * 64 + ((c ^ sw) << 4)) = u;
    }
  }
  __syncwarp();
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out);
  const int c = lane & 3;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int r = it * 8 + (lane >> 2);
    uint4 u = *reinterpret_cast<const uint4*>(stile + r * 64 + ((c ^ ((r >> 1) & 3)) << 4));
    const int grow = row0 + r;
    const int gcol = col0 + c * 8;
    if (grow < p.M && gcol < p.N) {
      *reinterpret_cast<uint4*>(out + (size_t)grow * p.ldo + gcol) = u;
    }
  }
  __syncwarp();
}

// Residual epilogue (EPI 2): fp32 acc transposed through smem (128 B per row), then in the coalesced layout
// out = residual + gate * (acc + bias); optional bf16 aux copy.
// In the coalesced layout lane l owns columns [4*(l&7), +4) of rows 4*it + (l>>3), it = 0..7.
// The residual fragment of chunk c+1 is loaded BEFORE chunk c is processed (software prefetch): the epilogue would
// otherwise serialise one DRAM round trip per 4 rows (loads may not be hoisted over the stores by the compiler because
// `residual` may alias `out`), which made the K=1152 residual GEMMs epilogue-bound (profiles/r1: 11% tensor pipe).
struct ResFrag {
  float4 r[8];
};

template <typename OutT>
PXA_DEVICE void load_residual_frag(ResFrag& f, const GemmParams& p, int lane, int row0, int col0, int m_limit = -1) {
  const int m_end = m_limit >= 0 ? m_limit : p.M;       // convolution with a virtual width: rows [row0, m_limit) exist
  const int gcol = col0 + (lane & 7) * 4;
  const bool col_ok = gcol < p.N;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int grow = row0 + it * 4 + (lane >> 3);
    f.r[it] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (grow < m_end && col_ok) {
      const size_t off = (size_t)grow * p.ldo + gcol;
      if constexpr (sizeof(OutT) == 4) {
        f.r[it] = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.residual) + off);
      } else {
        uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p.residual) + off);
        f.r[it] = make_float4(bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y));
      }
    }
  }
}

template <typename OutT>
PXA_DEVICE void epilogue_chunk_residual(uint32_t (&v)[32], const ResFrag& res, const GemmParams& p, uint8_t* stile,
                                        int lane, int row0, int col0, int m_limit = -1) {
  const int m_end = m_limit >= 0 ? m_limit : p.M;
  // smem tile: 32 rows x 128 B; 16-byte chunk c of row r lives at r*128 + ((c ^ (r & 7)) * 16)
  {
    const int sw = lane & 7;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      uint4 u = make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
      *reinterpret_cast<uint4*>(stile + lane * 128 + ((c ^ sw) << 4)) = u;
    }
  }
  __syncwarp();
  const int c = lane & 7;
  const int gcol = col0 + c * 4;
  const bool col_ok = gcol < p.N;
  float b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
  if (p.bias != nullptr && col_ok) {
    uint2 u = __ldg(reinterpret_cast<const uint2*>(p.bias + gcol));
    b0 = bf16_lo(u.x); b1 = bf16_hi(u.x); b2 = bf16_lo(u.y); b3 = bf16_hi(u.y);
  }
  // gate: all loads first (tiny table, L1/L2 resident), then compute + store
  float4 g[8];
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int grow = row0 + it * 4 + (lane >> 3);
    g[it] = make_float4(1.f, 1.f, 1.f, 1.f);
    if (p.gate != nullptr && grow < m_end && col_ok) {
      const int bidx = grow / p.rows_per_batch;
      g[it] = __ldg(reinterpret_cast<const float4*>(p.gate + (size_t)bidx * p.gate_batch_stride + gcol));
    }
  }
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int r = it * 4 + (lane >> 3);
    const int grow = row0 + r;
    const float4 a = *reinterpret_cast<const float4*>(stile + r * 128 + ((c ^ (r & 7)) << 4));
    if (grow < m_end && col_ok) {
      const size_t off = (size_t)grow * p.ldo + gcol;
      float4 o;
      const float4 yb = make_float4(a.x + b0, a.y + b1, a.z + b2, a.w + b3);     // the branch output
      o.x = fmaf(g[it].x, yb.x, res.r[it].x);
      o.y = fmaf(g[it].y, yb.y, res.r[it].y);
      o.z = fmaf(g[it].z, yb.z, res.r[it].z);
      o.w = fmaf(g[it].w, yb.w, res.r[it].w);
      if constexpr (sizeof(OutT) == 4) {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + off) = o;
      } else {
        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + off) =
            make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
      }
      if (p.out_aux != nullptr) {
        const float4 ax = p.aux_branch ? yb : o;
        *reinterpret_cast<uint2*>(p.out_aux + off) = make_uint2(pack_bf16x2(ax.x, ax.y), pack_bf16x2(ax.z, ax.w));
      }
    }
  }
  __syncwarp();
}


// Per-tile constants in two phases so that their L2 latency hides behind the PREVIOUS tile's epilogue work: the 128
// epilogue threads issue the global loads of tile i + 1 (load_epi_consts) right after tile i's barrier and park the values in
// registers; at the top of tile i + 1 they are stored to the tile's EpiConst buffer (store_epi_consts), followed by a named
// barrier.  (Loading and storing in one go put one L2 round trip, ~1 us, in front of every tile's epilogue.)
template <int BN> struct EpiRegs {
  static constexpr int kCols = (BN + kNumEpiThreads - 1) / kNumEpiThreads;
  float v[kCols][5];
  int row_split;
};

template <int BN, bool kLn = false>
PXA_DEVICE void load_epi_consts(EpiRegs<BN>& r, const GemmParams& p, int tid, int m0, int n0) {
  const int b0 = m0 / p.rows_per_batch;
  const int last = (m0 + kBM - 1 < p.M ? m0 + kBM - 1 : p.M - 1);
  const int b1 = last / p.rows_per_batch;
#pragma unroll
  for (int i = 0; i < EpiRegs<BN>::kCols; ++i) {
    const int col = n0 + tid + i * kNumEpiThreads;
    const bool ok = col < p.N && tid + i * kNumEpiThreads < BN;
    if constexpr (kLn) {
      r.v[i][0] = ok ? __ldg(p.ln_v + (size_t)b0 * p.ln_uv_batch_stride + col) : 0.f;
      r.v[i][1] = ok ? __ldg(p.ln_v + (size_t)b1 * p.ln_uv_batch_stride + col) : 0.f;
      r.v[i][2] = 1.f;
      r.v[i][3] = ok ? __ldg(p.ln_u + (size_t)b0 * p.ln_uv_batch_stride + col) : 0.f;
      r.v[i][4] = ok ? __ldg(p.ln_u + (size_t)b1 * p.ln_uv_batch_stride + col) : 0.f;
    } else {
      r.v[i][0] = (ok && p.bias != nullptr) ? __bfloat162float(p.bias[col]) : 0.f;
      r.v[i][1] = (ok && p.gate != nullptr) ? __ldg(p.gate + (size_t)b0 * p.gate_batch_stride + col) : 1.f;
      r.v[i][2] = (ok && p.gate != nullptr) ? __ldg(p.gate + (size_t)b1 * p.gate_batch_stride + col) : 1.f;
      r.v[i][3] = (ok && p.aux_scale != nullptr) ? __ldg(p.aux_scale + (size_t)b0 * p.aux_scale_batch_stride + col) : 1.f;
      r.v[i][4] = (ok && p.aux_scale != nullptr) ? __ldg(p.aux_scale + (size_t)b1 * p.aux_scale_batch_stride + col) : 1.f;
    }
  }
  r.row_split = (b0 + 1) * p.rows_per_batch - m0;
}

template <int BN>
PXA_DEVICE void store_epi_consts(EpiConst* cb, const EpiRegs<BN>& r, int tid) {
#pragma unroll
  for (int i = 0; i < EpiRegs<BN>::kCols; ++i) {
    const int c = tid + i * kNumEpiThreads;
    if (c < BN) {
      cb->bias[c] = r.v[i][0];
      cb->gate0[c] = r.v[i][1];
      cb->gate1[c] = r.v[i][2];
      cb->ex0[c] = r.v[i][3];
      cb->ex1[c] = r.v[i][4];
    }
  }
  if (tid == 0) cb->row_split = r.row_split;
}

// The PXA_LN_STAT_PARTS partial (sum, sum of squares) pairs of row `row` (fused LayerNorm consumer), loaded one tile ahead.
struct LnStatRegs {
  float4 t[PXA_LN_STAT_PARTS / 2];
};
PXA_DEVICE void load_ln_stats(LnStatRegs& r, const GemmParams& p, int row) {
  const float4* sp = reinterpret_cast<const float4*>(p.ln_stats + (size_t)(row < p.M ? row : p.M - 1) * (2 * PXA_LN_STAT_PARTS));
#pragma unroll
  for (int i = 0; i < PXA_LN_STAT_PARTS / 2; ++i) r.t[i] = __ldg(sp + i);
}
// (rstd, -rstd * mean), so that the normalised accumulator is fma(rstd, acc, fma(-rstd * mean, u, v)).
PXA_DEVICE float2 ln_row_coeffs(const GemmParams& p, const LnStatRegs& r) {
  float s = 0.f, q = 0.f;
#pragma unroll
  for (int i = 0; i < PXA_LN_STAT_PARTS / 2; ++i) {
    s += r.t[i].x + r.t[i].z;
    q += r.t[i].y + r.t[i].w;
  }
  const float mean = s * p.ln_inv_dim;
  const float var = fmaxf(q * p.ln_inv_dim - mean * mean, 0.f);
  const float rstd = rsqrtf(var + p.ln_eps);
  return make_float2(rstd, -rstd * mean);
}

// bf16 output (EPI 0/1/3/4) with staged constants: bias from smem (LDS broadcast).
//   PXA_EPI_BIAS_GELU_AUX: out = gelu(acc + bias) and, through the same smem transpose, out_aux = acc + bias (the
//                          pre-activation the GELU backward needs) -- the MLP's first GEMM in training;
//   PXA_EPI_MUL_DGELU:     out = acc * gelu'(pre), pre = p.residual read as bf16 [M, N] (row stride ldo) -- the dgrad GEMM of
//                          the MLP's second layer producing the gradient of the pre-activation directly.
//   PXA_EPI_LN_BIAS(_GELU): out = [gelu](rstd * (acc - mu * u) + v) -- the fused LayerNorm-modulate (see the header); `ln` =
//                          (rstd, -rstd * mu) of this thread's row, u / v of the row's sample from the staged constants.
template <int EPI>
PXA_DEVICE void epilogue_chunk_bf16_c(uint32_t (&v)[32], const GemmParams& p, const EpiConst* cb, uint8_t* stile, int lane,
                                      int row0, int col0, int ccol, float2 ln = make_float2(1.f, 0.f), bool second = false,
                                      int m_limit = -1) {
  constexpr bool kLn = (EPI == PXA_EPI_LN_BIAS || EPI == PXA_EPI_LN_BIAS_GELU);
  const int m_end = m_limit >= 0 ? m_limit : p.M;
  uint32_t pk[16];
  [[maybe_unused]] uint32_t pk2[16];
  const float4* bp = reinterpret_cast<const float4*>((kLn && second ? cb->gate0 : cb->bias) + ccol);
  [[maybe_unused]] const float4* up = reinterpret_cast<const float4*>((second ? cb->ex1 : cb->ex0) + ccol);
  [[maybe_unused]] uint4 pre[4];
  if constexpr (EPI == PXA_EPI_MUL_DGELU) {
    const __nv_bfloat16* pr = reinterpret_cast<const __nv_bfloat16*>(p.residual) + (size_t)(row0 + lane) * p.ldo + col0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      pre[i] = (row0 + lane < p.M && col0 + 8 * i < p.N) ? *reinterpret_cast<const uint4*>(pr + 8 * i) : make_uint4(0u, 0u, 0u, 0u);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 b = bp[i];
    float x0, x1, x2, x3;
    if constexpr (kLn) {
      const float4 u = up[i];
      x0 = fmaf(ln.x, __uint_as_float(v[4 * i]), fmaf(ln.y, u.x, b.x));
      x1 = fmaf(ln.x, __uint_as_float(v[4 * i + 1]), fmaf(ln.y, u.y, b.y));
      x2 = fmaf(ln.x, __uint_as_float(v[4 * i + 2]), fmaf(ln.y, u.z, b.z));
      x3 = fmaf(ln.x, __uint_as_float(v[4 * i + 3]), fmaf(ln.y, u.w, b.w));
    } else {
      x0 = __uint_as_float(v[4 * i]) + b.x; x1 = __uint_as_float(v[4 * i + 1]) + b.y;
      x2 = __uint_as_float(v[4 * i + 2]) + b.z; x3 = __uint_as_float(v[4 * i + 3]) + b.w;
    }
    if constexpr (EPI == PXA_EPI_BIAS_GELU_AUX) {
      pk2[2 * i] = pack_bf16x2(x0, x1);
      pk2[2 * i + 1] = pack_bf16x2(x2, x3);
    }
    if (EPI == PXA_EPI_BIAS_GELU || EPI == PXA_EPI_BIAS_GELU_AUX || EPI == PXA_EPI_LN_BIAS_GELU) {
      x0 = gelu_tanh(x0); x1 = gelu_tanh(x1); x2 = gelu_tanh(x2); x3 = gelu_tanh(x3);
    }
    if constexpr (EPI == PXA_EPI_MUL_DGELU) {
      const uint32_t w0 = (i & 1) ? pre[i >> 1].z : pre[i >> 1].x, w1 = (i & 1) ? pre[i >> 1].w : pre[i >> 1].y;
      x0 *= gelu_tanh_grad(bf16_lo(w0)); x1 *= gelu_tanh_grad(bf16_hi(w0));
      x2 *= gelu_tanh_grad(bf16_lo(w1)); x3 *= gelu_tanh_grad(bf16_hi(w1));
    }
    pk[2 * i] = pack_bf16x2(x0, x1);
    pk[2 * i + 1] = pack_bf16x2(x2, x3);
  }
  {
    const int sw = (lane >> 1) & 3;
#pragma unroll
    for (int c = 0; c < 4; ++c)
      *reinterpret_cast<uint4*>(stile + lane * 64 + ((c ^ sw) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
  }
  __syncwarp();
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out);
  const int c = lane & 3;
  uint4 u[4];
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int r = it * 8 + (lane >> 2);
    u[it] = *reinterpret_cast<const uint4*>(stile + r * 64 + ((c ^ ((r >> 1) & 3)) << 4));
  }
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int grow = row0 + it * 8 + (lane >> 2);
    const int gcol = col0 + c * 8;
    if (grow < m_end && gcol < p.N) *reinterpret_cast<uint4*>(out + (size_t)grow * p.ldo + gcol) = u[it];
  }
  __syncwarp();
  if constexpr (EPI == PXA_EPI_BIAS_GELU_AUX) {              // second pass of the same transpose for the pre-activation
    const int sw = (lane >> 1) & 3;
#pragma unroll
    for (int cc = 0; cc < 4; ++cc)
      *reinterpret_cast<uint4*>(stile + lane * 64 + ((cc ^ sw) << 4)) = make_uint4(pk2[4 * cc], pk2[4 * cc + 1], pk2[4 * cc + 2], pk2[4 * cc + 3]);
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int r = it * 8 + (lane >> 2);
      const uint4 w = *reinterpret_cast<const uint4*>(stile + r * 64 + ((c ^ ((r >> 1) & 3)) << 4));
      const int grow = row0 + r;
      const int gcol = col0 + c * 8;
      if (grow < p.M && gcol < p.N) *reinterpret_cast<uint4*>(p.out_aux + (size_t)grow * p.ldo + gcol) = w;
    }
    __syncwarp();
  }
}

// One residual chunk for the calling thread (row r of the 128-row tile) with staged constants.
// `reduce`: the chunk buffer receives only the update gate*(acc+bias); the TMA engine adds it into the residual stream
// in global memory (no residual read at all).
// `st` accumulates this row's (sum, sum of squares) of the new residual values over the chunks of a tile (the fused
// LayerNorm's statistics, p.row_stats_out); with p.aux_scale the bf16 aux copy is out * (1 + scale of the next modulate).
PXA_DEVICE void residual_chunk_row_c(uint32_t (&v)[32], const GemmParams& p, const EpiConst* cb, uint8_t* rbuf, uint8_t* abuf,
                                     int r, int ccol, int grow, int gcol0, bool reduce, float2& st) {
  const int sw = r & 7;
  uint8_t* rrow = rbuf + r * 128;
  float4 res[8];
#pragma unroll
  for (int c = 0; c < 8; ++c)                                                                        // all loads first
    res[c] = reduce ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<const float4*>(rrow + ((c ^ sw) << 4));
  const float4* bp = reinterpret_cast<const float4*>(cb->bias + ccol);
  const float4* gp = reinterpret_cast<const float4*>((r >= cb->row_split ? cb->gate1 : cb->gate0) + ccol);
  const float4* ep = reinterpret_cast<const float4*>((r >= cb->row_split ? cb->ex1 : cb->ex0) + ccol);
  const bool scaled_aux = p.aux_scale != nullptr;
  // samples shorter than a tile (rows_per_batch < 128, tiny images): a tile may span more than two samples, so the
  // gate row is looked up per thread in global memory instead of the two staged rows
  const bool per_row_gate = p.gate != nullptr && p.rows_per_batch < kBM;
  const float* gg = nullptr;
  if (per_row_gate) gg = p.gate + (size_t)((grow < p.M ? grow : p.M - 1) / p.rows_per_batch) * p.gate_batch_stride + gcol0;
  uint32_t aux[16];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const float4 b = bp[c];
    float4 g = gp[c];
    if (per_row_gate) g = (gcol0 + 4 * c < p.N) ? __ldg(reinterpret_cast<const float4*>(gg + 4 * c)) : make_float4(1.f, 1.f, 1.f, 1.f);
    float4 o;
    const float4 yb = make_float4(__uint_as_float(v[4 * c + 0]) + b.x, __uint_as_float(v[4 * c + 1]) + b.y,
                                  __uint_as_float(v[4 * c + 2]) + b.z, __uint_as_float(v[4 * c + 3]) + b.w);
    o.x = fmaf(g.x, yb.x, res[c].x);
    o.y = fmaf(g.y, yb.y, res[c].y);
    o.z = fmaf(g.z, yb.z, res[c].z);
    o.w = fmaf(g.w, yb.w, res[c].w);
    res[c] = o;
    st.x += (o.x + o.y) + (o.z + o.w);
    st.y = fmaf(o.x, o.x, fmaf(o.y, o.y, fmaf(o.z, o.z, fmaf(o.w, o.w, st.y))));
    float4 ax = p.aux_branch ? yb : o;
    if (scaled_aux) {
      const float4 e = ep[c];
      ax = make_float4(ax.x * e.x, ax.y * e.y, ax.z * e.z, ax.w * e.w);
    }
    aux[2 * c] = pack_bf16x2(ax.x, ax.y);
    aux[2 * c + 1] = pack_bf16x2(ax.z, ax.w);
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) *reinterpret_cast<float4*>(rrow + ((c ^ sw) << 4)) = res[c];           // then all stores
  if (abuf != nullptr) {
    const int sw2 = (r >> 1) & 3;
    uint8_t* arow = abuf + r * 64;
#pragma unroll
    for (int c = 0; c < 4; ++c)
      *reinterpret_cast<uint4*>(arow + ((c ^ sw2) << 4)) = make_uint4(aux[4 * c], aux[4 * c + 1], aux[4 * c + 2], aux[4 * c + 3]);
  }
}

// ------------------------------------------------------------------------------------------------- TMA residual epilogue
// fp32 residual stream: out = residual + gate * (acc + bias), the residual tile streamed global -> smem and the result
// smem -> global by TMA in 128-row x 32-column chunks (kResBufs chunk buffers, loads issued kResBufs-1 chunks ahead,
// across tile boundaries).  This keeps ~32-48 KB of residual reads in flight per SM without holding registers, which
// the register-prefetch version could not (Little's law: 16 KB in flight per SM capped it at ~2.4 TB/s chip-wide).
// Row-per-thread all the way: no transpose, TMA clips the M / N tails.  Chunk buffers use the 128B TMA swizzle so the
// row-per-thread 16-byte accesses are bank-conflict free; the optional bf16 aux copy uses a 64B-swizzled buffer.
struct ResTileCursor {      // walks the (tile, chunk) sequence of this CTA
  int tile, cc;
};

template <int BN>
PXA_DEVICE int chunks_of_tile(const GemmParams& p, int n0) {
  const int rem = (p.N - n0 + 31) / 32;
  return rem < BN / 32 ? rem : BN / 32;
}

// Arrive on the barrier at the same smem offset in CTA `rank` of the cluster (CTA-pair kernels).
PXA_DEVICE void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}\n" ::"r"(smem_u32(bar)),
      "r"(rank)
      : "memory");
}

// Which output tiles a CTA walks, shared by the single-CTA kernel (128 x BN tiles, one per CTA) and the CTA-pair kernel
// (256 x BN tiles per cluster: m_mult = 256, this CTA's rows start m_off = rank * 128 into the tile).
struct TileWalk {
  int first, stride, count;          // tile ids first, first + stride, ... < count
  int mn_tiles, num_n_tiles;         // tile id % mn_tiles -> (m tile, n tile), n fastest
  int m_mult, m_off;
  int reverse;                       // 1: tile id t stands for tile count - 1 - t
  PXA_DEVICE int eff(int tile) const { return reverse ? count - 1 - tile : tile; }
  PXA_DEVICE int m0(int tile) const { return ((eff(tile) % mn_tiles) / num_n_tiles) * m_mult + m_off; }
  PXA_DEVICE int ntile(int tile) const { return (eff(tile) % mn_tiles) % num_n_tiles; }
};

// The whole epilogue-warp loop of the TMA-streamed fp32 residual epilogue (called by the 4 epilogue warps of either GEMM
// kernel): per tile, per 32-column chunk: accumulator chunk TMEM -> registers (split-phase), residual chunk smem (TMA) ->
// out = residual + gate * (acc + bias) in place in smem -> TMA store (+ bf16 aux chunk), or, with `reduce`, update only +
// TMA reduce-add into global memory.  One elected thread of the first epilogue warp owns every TMA / bulk-group operation.
template <int BN, int kRing, bool kPair>
PXA_DEVICE void tma_res_epilogue(const GemmParams& p, const TileWalk& tw, uint8_t* epi_smem, EpiConst* consts,
                                 uint64_t* tfull_bar, uint64_t* tempty_bar, uint64_t* res_full, uint32_t tmem_base,
                                 const CUtensorMap* tmap_res, const CUtensorMap* tmap_out, const CUtensorMap* tmap_aux,
                                 int warp, int lane) {
  const int q = warp & 3;                                       // TMEM sub-partition of this warp
  const int tid = threadIdx.x - kEpiWarp0 * 32;                 // 0..127
  const int r = q * 32 + lane;                                  // row of the 128-row tile owned by this thread
  uint8_t* rbufs = epi_smem;
  uint8_t* abufs = epi_smem + kRing * kResChunkBytes;           // (unused by the reduce-only configuration)
  // The issuer is picked with elect.sync inside a warp-uniform branch (always the same lane for a full warp), so the TMA /
  // mbarrier operands stay in uniform registers; `issuer_warp` guards the converged regions, `issuer` the elected lane.
  const bool issuer_warp = warp == kEpiWarp0;
  bool issuer = false;
  if (issuer_warp) issuer = elect_one() != 0;
  // In-place update without a bf16 copy or statistics: skip the residual read and let TMA reduce-add the update into x.
  const bool reduce = p.out_aux == nullptr && p.residual == p.out && p.row_stats_out == nullptr;
  int l_tile = tw.first, l_cc = 0, l_g = 0;                     // issuer: next residual chunk to request
  auto request_next = [&]() {
    if (l_tile >= tw.count) return;
    const int lm0 = tw.m0(l_tile);
    const int ln0 = tw.ntile(l_tile) * BN;
    const int buf = l_g % kRing;
    mbar_arrive_expect_tx(&res_full[buf], kResChunkBytes);
    tma_load_2d(rbufs + buf * kResChunkBytes, tmap_res, &res_full[buf], ln0 + l_cc * 32, lm0, kEvictFirst);
    ++l_g;
    if (++l_cc == chunks_of_tile<BN>(p, ln0)) { l_cc = 0; l_tile += tw.stride; }
  };
  if (issuer_warp && !reduce) {
    if (elect_one()) {
      for (int i = 0; i < kRing - 1; ++i) request_next();
    }
  }
  int g = 0, as = 0, titer = 0;
  uint32_t aphase = 0;
  EpiRegs<BN> er;                                               // this tile's constants, loaded one tile ahead
  if (tw.first < tw.count) load_epi_consts<BN>(er, p, tid, tw.m0(tw.first), tw.ntile(tw.first) * BN);
  const bool tracing = issuer && p.trace != nullptr && blockIdx.x == 0;   // debug only (tools/gemm_trace.py)
  int tcnt = 0;
  auto stamp = [&]() {
    if (tracing && tcnt < 4096) p.trace[tcnt++] = clock64();
  };

  for (int tile = tw.first; tile < tw.count; tile += tw.stride, ++titer) {
    const int m0 = tw.m0(tile);
    const int nt = tw.ntile(tile);
    const int n0 = nt * BN;
    const int nch = chunks_of_tile<BN>(p, n0);
    EpiConst* cb = consts + (titer & 1);
    store_epi_consts<BN>(cb, er, tid);
    named_bar_sync(2, kNumEpiThreads);
    if (tile + tw.stride < tw.count)                            // next tile's constants: in flight during this tile's chunks
      load_epi_consts<BN>(er, p, tid, tw.m0(tile + tw.stride), tw.ntile(tile + tw.stride) * BN);
    stamp();                                                    // tile: start waiting for the accumulator
    mbar_wait(&tfull_bar[as], aphase);
    stamp();                                                    // tile: accumulator ready
    tc_fence_after();
    const uint32_t t_acc = tmem_base + as * BN + (static_cast<uint32_t>(q * 32) << 16);
    float2 st = make_float2(0.f, 0.f);                          // this row's (sum, sum of squares) over the tile

    auto process = [&](uint32_t (&v)[32], int cc) {
      const int buf = g % kRing;
      uint8_t* rb = rbufs + buf * kResChunkBytes;
      uint8_t* ab = p.out_aux != nullptr ? abufs + (g & 1) * kAuxChunkBytes : nullptr;
      stamp();                                                  // chunk: acc in registers
      if (!reduce) mbar_wait(&res_full[buf], (g / kRing) & 1);  // residual chunk has landed in smem
      stamp();                                                  // chunk: residual landed
      residual_chunk_row_c(v, p, cb, rb, ab, r, cc * 32, m0 + r, n0 + cc * 32, reduce, st);
      fence_proxy_async_smem();                                 // generic-proxy writes -> visible to the TMA store
      stamp();                                                  // chunk: computed
      if (issuer_warp) {
        if (elect_one()) tma_store_wait_read<0>();              // earlier stores have drained their buffers
      }
      stamp();                                                  // chunk: previous store drained
      named_bar_sync(1, kNumEpiThreads);
      stamp();                                                  // chunk: barrier passed
      if (issuer_warp) {
        if (elect_one()) {
          if (reduce) {
            tma_reduce_add_2d(tmap_out, rb, n0 + cc * 32, m0);
            tma_store_commit();
          } else {
            tma_store_2d(tmap_out, rb, n0 + cc * 32, m0);
            if (ab != nullptr) tma_store_2d(tmap_aux, ab, n0 + cc * 32, m0);
            tma_store_commit();
            request_next();                                     // refills the buffer chunk g-1 has just left
          }
        }
      }
      ++g;
    };
    auto release_acc = [&]() {                                  // all TMEM reads of this accumulator are done
      tc_fence_before();
      if constexpr (kPair) mbar_arrive_cluster(&tempty_bar[as], 0);   // the leader's MMA thread owns the hand-off
      else mbar_arrive(&tempty_bar[as]);
    };

    // software pipeline: the TMEM load of chunk c+1 is in flight while chunk c is processed
    uint32_t va[32], vb[32];
    tmem_ld_32x32b_x32_nowait(t_acc, va);
#pragma unroll 1
    for (int cc = 0; cc < nch; cc += 2) {
      tmem_ld_wait_x32(va);
      if (cc + 1 < nch) tmem_ld_32x32b_x32_nowait(t_acc + (cc + 1) * 32, vb); else release_acc();
      process(va, cc);
      if (cc + 1 < nch) {
        tmem_ld_wait_x32(vb);
        if (cc + 2 < nch) tmem_ld_32x32b_x32_nowait(t_acc + (cc + 2) * 32, va); else release_acc();
        process(vb, cc + 1);
      }
    }
    if (p.row_stats_out != nullptr && m0 + r < p.M) {           // partial LayerNorm statistics of this row: part = n tile
      float2* sp = reinterpret_cast<float2*>(p.row_stats_out) + (size_t)(m0 + r) * PXA_LN_STAT_PARTS;
      sp[nt] = st;
      if (nt == tw.num_n_tiles - 1)
        for (int i = tw.num_n_tiles; i < PXA_LN_STAT_PARTS; ++i) sp[i] = make_float2(0.f, 0.f);
    }
    as ^= 1;
    if (as == 0) aphase ^= 1;
  }
  if (issuer_warp) {
    if (elect_one()) tma_store_wait_all<0>();
  }
}

}  // namespace pxa
