// pxa_t5_attn_d64_bf16: softmax(Q K^T * scale + bias[h] + key_bias[b]) V for head_dim 64 and short sequences (L <= 384) on tcgen05.
//
// The self-attention of the T5-v1.1-XXL caption encoder (transformers `T5Attention.forward`; reference call site
// diffusion/model/t5.py:107-110): 64 heads of 64, 300 tokens, NO 1/sqrt(d) scale, a learned relative-position bias
// `position_bias[h, i, j]` added to the logits, padded keys masked additively.  The DiT's flash kernel (attn_sm100.cu) takes neither a
// bias nor head_dim 64; with <= 384 keys there is no need for its key loop or online softmax either: the whole S row fits TMEM.
//
// One CTA = 128 query rows of one (sample, head); 192 threads:
//   warp 0      one elected thread: TMA loads (Q tile, all K / V rows of the sample: <= 3 boxes of 128 keys each), then the MMAs:
//               S[128 x 128 j..] = Q K_j^T (SS, 4 K-steps of 16 per key box), later O = P V (TS: P read from TMEM, V MN-major in its
//               natural [key, d] layout), one N = 64 MMA per 16 keys
//   warp 1      TMEM allocator (512 columns: S at 0..383, O at 384..447)
//   warps 2-5   softmax, one thread per query row: pass 1 adds the bias (T5's relative-position bias by offset j - i from a 2 L - 1
//               entry smem vector, or a dense [H, L, L] table from global memory) to S in TMEM and finds the row max, pass 2 forms
//               P = exp2((s - max) log2 e), accumulates the row sum and writes P as bf16 over the S columns already consumed
//               (P chunk c lands in columns 16 c .. 16 c + 15, inside S chunk c / 2 <= c); then O / sum -> bf16 -> global.
// Every barrier is used once (phase 0): the kernel is a straight line, not a pipeline -- at 300 keys a CTA lives ~10 us and 5 waves of
// 768 CTAs (4 captions x 64 heads x 3 query tiles) overlap each other's loads and epilogues across the SMs.
// Algorithmic work: 4 L^2 64 FLOP per (sample, head); bytes: q, k, v, out once + the fp32 bias rows (L2-resident, 23 MB at L = 300).
#include "host_common.cuh"
#include "ptx.cuh"

namespace pxa {

constexpr int kT5Threads = 192;
constexpr int kT5D = 64;
constexpr int kT5Tile = 128;                       // query rows per CTA, keys per TMA box
constexpr int kT5MaxKeyBoxes = 3;                  // L <= 384
constexpr int kT5BoxBytes = kT5Tile * kT5D * 2;    // 16 KB: 128 rows x 128 B, 128B-swizzled
constexpr int kT5OffQ = 0;
constexpr int kT5OffK = kT5OffQ + kT5BoxBytes;
constexpr int kT5OffV = kT5OffK + kT5MaxKeyBoxes * kT5BoxBytes;
constexpr int kT5OffBars = kT5OffV + kT5MaxKeyBoxes * kT5BoxBytes;
constexpr int kT5MaxL = kT5Tile * kT5MaxKeyBoxes;  // 384
constexpr int kT5OffRel = kT5OffBars + 128;        // fp32 [2 * 384]: the head's relative-position bias by offset j - i + L - 1
constexpr int kT5OffKeyBias = kT5OffRel + 2 * kT5MaxL * 4;   // fp32 [384]: the sample's additive key mask
constexpr int kT5Smem = kT5OffKeyBias + kT5MaxL * 4 + 1024;  // + alignment slack
constexpr uint32_t kT5ColO = 384;

struct T5AttnParams {
  __nv_bfloat16* out;
  const float* bias;        // [H, L, L], or nullptr when rel_bias is given
  const float* rel_bias;    // [H, 2 L - 1] or nullptr: bias[h, i, j] = rel_bias[h, j - i + L - 1] (T5: the bias is Toeplitz)
  const float* key_bias;    // [B, L] or nullptr
  int B, H, L, ldo;
  float scale_log2;         // scale * log2(e)
  float log2e;
};

__global__ void __launch_bounds__(kT5Threads, 1)
t5_attn_d64_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                   const __grid_constant__ CUtensorMap tm_v, const T5AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kT5OffBars);
  uint64_t* qk_full = bars;          // TMA -> MMA: Q tile + all K boxes landed
  uint64_t* v_full = bars + 1;       // TMA -> MMA: all V boxes landed
  uint64_t* s_full = bars + 2;       // MMA -> softmax: S complete
  uint64_t* p_full = bars + 3;       // softmax (128 arrivals) -> MMA: P written
  uint64_t* o_full = bars + 4;       // MMA -> softmax: O complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  const int warp = warp_idx_sync();
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kT5Tile, h = blockIdx.y, b = blockIdx.z;
  const int L = p.L;
  const int nkb = (L + kT5Tile - 1) / kT5Tile;         // 128-key boxes
  const int n16 = (L + 15) / 16;                       // 16-key MMA steps of P V
  const int nch = (L + 31) / 32;                       // 32-key softmax chunks

  if (threadIdx.x == 0) {
    prefetch_tmap(&tm_q); prefetch_tmap(&tm_k); prefetch_tmap(&tm_v);
    mbar_init(qk_full, 1); mbar_init(v_full, 1); mbar_init(s_full, 1); mbar_init(p_full, 128); mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      const int row0 = b * L;
      mbar_arrive_expect_tx(qk_full, (1 + nkb) * kT5BoxBytes);
      // rows past the sample's end read the next sample's tokens (finite, masked / never stored) or TMA zero fill
      tma_load_3d(smem + kT5OffQ, &tm_q, qk_full, 0, h, row0 + q0, kEvictFirst);
      for (int j = 0; j < nkb; ++j) tma_load_3d(smem + kT5OffK + j * kT5BoxBytes, &tm_k, qk_full, 0, h, row0 + j * kT5Tile, kEvictLast);
      mbar_arrive_expect_tx(v_full, nkb * kT5BoxBytes);
      for (int j = 0; j < nkb; ++j) tma_load_3d(smem + kT5OffV + j * kT5BoxBytes, &tm_v, v_full, 0, h, row0 + j * kT5Tile, kEvictLast);

      constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, kT5D, 0, 1);       // V is MN-major: [key, d] rows of 128 B
      const uint32_t sbase = smem_u32(smem);
      mbar_wait(qk_full, 0);
      tc_fence_after();
      const uint64_t qd = make_smem_desc(sbase + kT5OffQ, 16, 1024, kLayoutSW128);
      for (int j = 0; j < nkb; ++j) {
        const uint64_t kd = make_smem_desc(sbase + kT5OffK + j * kT5BoxBytes, 16, 1024, kLayoutSW128);
#pragma unroll
        for (int k = 0; k < kT5D / 16; ++k) umma_ss(tmem_base + j * kT5Tile, qd + 2 * k, kd + 2 * k, idesc_qk, k != 0 ? 1u : 0u);
      }
      umma_commit(s_full);
      mbar_wait(v_full, 0);
      mbar_wait(p_full, 0);
      tc_fence_after();
      // the key boxes are contiguous in smem, so key row r sits at r * 128 B (inside its 1024-byte swizzle group) across boxes
      const uint64_t vd = make_smem_desc(sbase + kT5OffV, kT5BoxBytes, 1024, kLayoutSW128);
      for (int k = 0; k < n16; ++k)
        umma_ts(tmem_base + kT5ColO, tmem_base + 8 * k, vd + (uint64_t)(k * (2048 >> 4)), idesc_pv, k != 0 ? 1u : 0u);
      umma_commit(o_full);
    }
  } else if (warp >= 2) {
    const int qd = warp & 3;                          // TMEM sub-partition this warp may access
    const int row = qd * 32 + lane;
    const int qi = q0 + row;                          // query index within the sample
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(qd * 32) << 16);
    const int qc = min(qi, L - 1);                    // rows >= L are computed on a clamped bias row, never stored
    const float* brow = p.bias ? p.bias + ((size_t)h * L + qc) * L : nullptr;
    const float sl2 = p.scale_log2, l2e = p.log2e;
    // The bias of a query row, one float per key.  Dense form: thread-per-row global loads touch 32 cache lines per instruction
    // (46 us per CTA at L = 300, session 34).  T5's bias depends on j - i only: 2 L - 1 floats per head staged in smem, where the 32
    // rows of a warp read 32 consecutive words; the key mask (one float per key, the same for every row) sits next to it.
    float* rel_s = reinterpret_cast<float*>(smem + kT5OffRel);
    float* kb_s = reinterpret_cast<float*>(smem + kT5OffKeyBias);
    {
      const int t0 = threadIdx.x - 64;                // 0..127
      if (p.rel_bias != nullptr)
        for (int t = t0; t < 2 * L - 1; t += 128) rel_s[t] = __ldg(p.rel_bias + (size_t)h * (2 * L - 1) + t);
      for (int t = t0; t < L; t += 128) kb_s[t] = p.key_bias ? __ldg(p.key_bias + (size_t)b * L + t) : 0.f;
      named_bar_sync(1, 128);
    }
    const float* rrow = rel_s + (L - 1 - qc);         // rrow[j] = rel_bias[h, j - i + L - 1]
    const bool use_rel = p.rel_bias != nullptr;

    mbar_wait(s_full, 0);
    tc_fence_after();
    // pass 1: t = (s * scale + bias) * log2(e) back into TMEM, row max
    float m = -INFINITY;
    for (int c = 0; c < nch; ++c) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(t_row + 32 * c, v);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int j = 32 * c + i;
        float t = -INFINITY;
        if (j < L) {
          const float add = (use_rel ? rrow[j] : __ldg(brow + j)) + kb_s[j];
          t = fmaf(__uint_as_float(v[i]), sl2, add * l2e);
        }
        m = fmaxf(m, t);
        v[i] = __float_as_uint(t);
      }
      tmem_st_32x32b_x32(t_row + 32 * c, v);
    }
    tmem_st_wait();
    if (m == -INFINITY) m = 0.f;                       // every key masked to -inf: P = 0, output zeros (no NaN)
    // pass 2: P = exp2(t - m) as bf16 over the consumed S columns, row sum of the fp32 values
    float sum = 0.f;
    for (int c = 0; c < nch; ++c) {
      uint32_t v[32], pk[16];
      tmem_ld_32x32b_x32(t_row + 32 * c, v);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float e0 = fast_exp2(__uint_as_float(v[2 * i]) - m);
        const float e1 = fast_exp2(__uint_as_float(v[2 * i + 1]) - m);
        sum += e0 + e1;
        pk[i] = pack_bf16x2(e0, e1);
      }
      tmem_st_32x32b_x16(t_row + 16 * c, pk);
    }
    tmem_st_wait();
    tc_fence_before();
    mbar_arrive(p_full);

    mbar_wait(o_full, 0);
    tc_fence_after();
    uint32_t oa[32], ob[32];
    tmem_ld_32x32b_x32_pair(t_row + kT5ColO, oa, t_row + kT5ColO + 32, ob);
    if (qi < L) {
      const float inv = sum > 0.f ? 1.0f / sum : 0.f;
      uint4* dst = reinterpret_cast<uint4*>(p.out + (size_t)(b * L + qi) * p.ldo + h * kT5D);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        dst[g] = make_uint4(pack_bf16x2(__uint_as_float(oa[8 * g]) * inv, __uint_as_float(oa[8 * g + 1]) * inv),
                            pack_bf16x2(__uint_as_float(oa[8 * g + 2]) * inv, __uint_as_float(oa[8 * g + 3]) * inv),
                            pack_bf16x2(__uint_as_float(oa[8 * g + 4]) * inv, __uint_as_float(oa[8 * g + 5]) * inv),
                            pack_bf16x2(__uint_as_float(oa[8 * g + 6]) * inv, __uint_as_float(oa[8 * g + 7]) * inv));
        dst[4 + g] = make_uint4(pack_bf16x2(__uint_as_float(ob[8 * g]) * inv, __uint_as_float(ob[8 * g + 1]) * inv),
                                pack_bf16x2(__uint_as_float(ob[8 * g + 2]) * inv, __uint_as_float(ob[8 * g + 3]) * inv),
                                pack_bf16x2(__uint_as_float(ob[8 * g + 4]) * inv, __uint_as_float(ob[8 * g + 5]) * inv),
                                pack_bf16x2(__uint_as_float(ob[8 * g + 6]) * inv, __uint_as_float(ob[8 * g + 7]) * inv));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

static int make_t5_map(CUtensorMap* map, const void* base, int H, long long rows, long long s_row, long long s_head) {
  uint64_t dims[3] = {(uint64_t)kT5D, (uint64_t)H, (uint64_t)rows};
  uint64_t str[2] = {(uint64_t)s_head * 2, (uint64_t)s_row * 2};
  uint32_t box[3] = {(uint32_t)kT5D, 1, (uint32_t)kT5Tile};
  return make_tmap_bf16(map, base, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

}  // namespace pxa

extern "C" int pxa_t5_attn_d64_bf16(const PxaT5AttnArgs* args, void* stream) {
  using namespace pxa;
  if (!args) return fail(PXA_ERR_ARG, "null args");
  const PxaT5AttnArgs& a = *args;
  if (!a.q || !a.k || !a.v || !a.out) return fail(PXA_ERR_ARG, "null q / k / v / out");
  if (!a.bias && !a.rel_bias) return fail(PXA_ERR_ARG, "one of bias [H, L, L] / rel_bias [H, 2L-1] is required");
  if (a.B <= 0 || a.H <= 0 || a.L <= 0) return fail(PXA_ERR_ARG, "bad B / H / L");
  if (a.L > kT5Tile * kT5MaxKeyBoxes) return fail(PXA_ERR_ARG, "L = %d: at most %d keys (the S row lives in TMEM)", a.L, kT5Tile * kT5MaxKeyBoxes);
  if ((a.q_sn & 7) || (a.q_sh & 7) || (a.k_sn & 7) || (a.k_sh & 7) || (a.v_sn & 7) || (a.v_sh & 7) || (a.ldo & 7))
    return fail(PXA_ERR_ALIGN, "strides must be multiples of 8 elements");
  if ((reinterpret_cast<uintptr_t>(a.out) & 15) || (reinterpret_cast<uintptr_t>(a.bias) & 3) || (reinterpret_cast<uintptr_t>(a.key_bias) & 3) ||
      (reinterpret_cast<uintptr_t>(a.rel_bias) & 3))
    return fail(PXA_ERR_ALIGN, "out must be 16-byte aligned, bias / key_bias 4-byte aligned");
  if ((long long)a.H * kT5D > a.ldo) return fail(PXA_ERR_ARG, "ldo smaller than H * 64");
  PXA_REQUIRE_SM100();
  CUtensorMap qm, km, vm;
  int rc;
  const long long rows = (long long)a.B * a.L;
  if ((rc = make_t5_map(&qm, a.q, a.H, rows, a.q_sn, a.q_sh))) return rc;
  if ((rc = make_t5_map(&km, a.k, a.H, rows, a.k_sn, a.k_sh))) return rc;
  if ((rc = make_t5_map(&vm, a.v, a.H, rows, a.v_sn, a.v_sh))) return rc;
  T5AttnParams p;
  p.out = reinterpret_cast<__nv_bfloat16*>(a.out);
  p.bias = a.rel_bias ? nullptr : a.bias;
  p.rel_bias = a.rel_bias;
  p.key_bias = a.key_bias;
  p.B = a.B; p.H = a.H; p.L = a.L; p.ldo = (int)a.ldo;
  p.log2e = 1.4426950408889634f;
  p.scale_log2 = a.scale * p.log2e;
  PXA_CHECK_CUDA(cudaFuncSetAttribute(t5_attn_d64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kT5Smem));
  dim3 grid((a.L + kT5Tile - 1) / kT5Tile, a.H, a.B);
  t5_attn_d64_kernel<<<grid, kT5Threads, kT5Smem, reinterpret_cast<cudaStream_t>(stream)>>>(qm, km, vm, p);
  launch_counter()++;
  PXA_CHECK_CUDA(cudaGetLastError());
  return PXA_OK;
}
