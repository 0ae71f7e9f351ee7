// pxa_flash_attn_d72_bf16, THREE-TILE variant (round 2): softmax(Q K^T * scale) V for head_dim 72 with three 128-row query
// tiles per CTA, i.e. THREE softmax warps per SM sub-partition instead of two.
//
// Why: the two-tile kernel (attn_sm100.cu) is bound by the exp2 section of its softmax warps.  Its cycle trace
// (profiles/r2_attn_trace.txt) shows the MUFU pipe saturated while the two warps of a sub-partition are both inside their
// exp2 sections, and idle in between: a warp alone cannot keep the MUFU pipe busy (tools/micro/exp_seq.cu: 671 cycles per 64
// exponentials for one warp, 1071 for two, MUFU floor 512 per warp), and each warp spends ~350 of its ~1400 cycles per
// sub-block outside the exp2 section (row max, P store, fences, barrier round trips).  With a third warp there are always two
// in the exp2 section (tools/micro/exp_mix.cu: 1577 cycles per 3 x 64 exponentials vs 1144 per 2 x 64).
//
// What it costs: TMEM.  512 columns hold 3 x (S 64 | O 80 | L 16) = 480, so S is a SINGLE 64-key buffer per tile (the
// two-tile kernel double-buffers it): a tile's softmax waits for its own P V (n) + Q K^T (n+1) round trip (~550 cycles) --
// during which the other two tiles keep the MUFU pipe busy.  K / V stream through 64-key stages (5-deep ring).
//
//   warp 0       TMA producer: Q (3 tiles) once, then K / V stages of 64 keys
//   warp 1       MMA issuer: per sub-block, per tile: O_t += P_t V (TS), L_t += P_t 1, then S_t = Q_t K^T of the next sub-block
//   warp 2       TMEM allocator; warp 3 initialises the ones tile
//   warps 4-15   softmax: tile t = (warp - 4) / 4, one thread per query row, online softmax with lazy rescale as in the
//                two-tile kernel (same numerics: bf16 P written over the S columns it came from, row sums on the tensor pipe)
// Everything else (head_dim-72 staging as main + tail, MN-major V, var-len keys, lse output) is attn_sm100.cu's.
#include <type_traits>

#include "host_common.cuh"
#include "ptx.cuh"

namespace pxa {

constexpr int k3Tiles = 3;
constexpr int k3Threads = 128 + k3Tiles * 128;          // 512
constexpr int k3D = 72;
constexpr int k3TileQ = 128;
constexpr int k3Sub = 64;                               // keys per stage = per MMA / softmax step
constexpr int k3Stages = 5;
constexpr int k3QMain = 128 * 128, k3QTail = 128 * 32;
constexpr int k3KMain = k3Sub * 128, k3KTail = k3Sub * 32;          // 8 KB + 2 KB
constexpr int k3VMain = k3Sub * 128, k3VTail = k3Sub * 128;         // 8 KB + 8 KB (wide tail: d 64..127, zero-filled past 71)
constexpr int k3OffQMain = 0;
constexpr int k3OffQTail = k3OffQMain + k3Tiles * k3QMain;          // tails behind the 1024-aligned mains
constexpr int k3OffKMain = k3OffQTail + k3Tiles * k3QTail;          // 61440: 1024-aligned
constexpr int k3OffVMain = k3OffKMain + k3Stages * k3KMain;
constexpr int k3OffVTail = k3OffVMain + k3Stages * k3VMain;
constexpr int k3OffKTail = k3OffVTail + k3Stages * k3VTail;
constexpr int k3OffOnes = k3OffKTail + k3Stages * k3KTail;
constexpr int k3OffBars = k3OffOnes + 2048;
constexpr int k3Smem = k3OffBars + 512 + 1024;
static_assert(k3OffKMain % 1024 == 0 && k3OffVMain % 1024 == 0 && k3OffVTail % 1024 == 0 && k3OffKTail % 256 == 0, "swizzle atoms");
static_assert(k3Smem <= 227 * 1024, "shared memory budget");

constexpr uint32_t k3TileCols = 160;                    // per tile: S 64 | O 80 | L 16
constexpr uint32_t k3ColO = 64, k3ColL = 144;

struct Attn3Params {
  __nv_bfloat16* out;
  float* lse;
  const int* kv_len;
  const int* kv_off;
  int B, H, Nq, Nk, ldo;
  float scale_log2;
  int reverse_batch;
};

__global__ void __launch_bounds__(k3Threads, 1)
flash_attn_d72_x3_kernel(const __grid_constant__ CUtensorMap tm_q_main, const __grid_constant__ CUtensorMap tm_q_tail,
                         const __grid_constant__ CUtensorMap tm_k_main, const __grid_constant__ CUtensorMap tm_k_tail,
                         const __grid_constant__ CUtensorMap tm_v_main, const __grid_constant__ CUtensorMap tm_v_tail,
                         const Attn3Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + k3OffBars);
  uint64_t* q_full = bars;                       // [1]
  uint64_t* k_full = bars + 1;                   // [k3Stages]
  uint64_t* k_empty = k_full + k3Stages;
  uint64_t* v_full = k_empty + k3Stages;
  uint64_t* v_empty = v_full + k3Stages;
  uint64_t* s_full = v_empty + k3Stages;         // [3] MMA -> softmax: S_t of the current sub-block ready
  uint64_t* p_full = s_full + k3Tiles;           // [3] softmax -> MMA: P_t written over S_t
  uint64_t* pv_done = p_full + k3Tiles;          // [3] MMA -> softmax: P_t V done (lazy-rescale path only)
  uint64_t* o_full = pv_done + k3Tiles;          // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = warp_idx_sync();
  const int lane = threadIdx.x & 31;
  const int b = p.reverse_batch ? p.B - 1 - (int)blockIdx.z : (int)blockIdx.z, h = blockIdx.y;
  const int q0 = blockIdx.x * (k3Tiles * k3TileQ);
  // tiles of this CTA that hold at least one query row (the last CTA of a sample may own fewer than three)
  const int n_tiles = min(k3Tiles, (p.Nq - q0 + k3TileQ - 1) / k3TileQ);

  int kv_len = p.kv_len ? p.kv_len[b] : p.Nk;
  kv_len = min(max(kv_len, 0), p.Nk);
  const int kv_row0 = p.kv_off ? p.kv_off[b] : b * p.Nk;
  const int n_sub = (kv_len + k3Sub - 1) / k3Sub;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tm_q_main); prefetch_tmap(&tm_q_tail);
    prefetch_tmap(&tm_k_main); prefetch_tmap(&tm_k_tail);
    prefetch_tmap(&tm_v_main); prefetch_tmap(&tm_v_tail);
    mbar_init(q_full, 1);
    for (int s = 0; s < k3Stages; ++s) {
      mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1);
    }
    for (int t = 0; t < k3Tiles; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t], 128);
      mbar_init(&pv_done[t], 1);
    }
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  if (warp == 3) {
    uint4* ones = reinterpret_cast<uint4*>(smem + k3OffOnes);
    for (int i = lane; i < 2048 / 16; i += 32) ones[i] = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================================================ TMA producer
    if (n_sub > 0 && elect_one()) {
      const int qrow = b * p.Nq + q0;
      mbar_arrive_expect_tx(q_full, n_tiles * (k3QMain + k3QTail));
      for (int t = 0; t < n_tiles; ++t) {
        tma_load_3d(smem + k3OffQMain + t * k3QMain, &tm_q_main, q_full, 0, h, qrow + t * k3TileQ, kEvictFirst);
        tma_load_3d(smem + k3OffQTail + t * k3QTail, &tm_q_tail, q_full, 64, h, qrow + t * k3TileQ, kEvictFirst);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_sub; ++j) {
        const int krow = kv_row0 + j * k3Sub;
        mbar_wait(&k_empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&k_full[stage], k3KMain + k3KTail);
        tma_load_3d(smem + k3OffKMain + stage * k3KMain, &tm_k_main, &k_full[stage], 0, h, krow, kEvictLast);
        tma_load_3d(smem + k3OffKTail + stage * k3KTail, &tm_k_tail, &k_full[stage], 64, h, krow, kEvictLast);
        mbar_wait(&v_empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&v_full[stage], k3VMain + k3VTail);
        tma_load_3d(smem + k3OffVMain + stage * k3VMain, &tm_v_main, &v_full[stage], 0, h, krow, kEvictLast);
        tma_load_3d(smem + k3OffVTail + stage * k3VTail, &tm_v_tail, &v_full[stage], 64, h, krow, kEvictLast);
        if (++stage == k3Stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    if (n_sub > 0 && elect_one()) {
      constexpr uint32_t idesc_qk = make_idesc_bf16(128, k3Sub, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, 80, 0, 1);        // V MN-major, N = d 0..79
      constexpr uint32_t idesc_l = make_idesc_bf16(128, 16, 0, 0);
      const uint32_t sbase = smem_u32(smem);
      auto issue_qk = [&](int t, int stage) {
        const uint64_t qd = make_smem_desc(sbase + k3OffQMain + t * k3QMain, 16, 1024, kLayoutSW128);
        const uint64_t kd = make_smem_desc(sbase + k3OffKMain + stage * k3KMain, 16, 1024, kLayoutSW128);
        const uint32_t d = tmem_base + t * k3TileCols;
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(d, qd + 2 * k, kd + 2 * k, idesc_qk, k != 0 ? 1u : 0u);
        const uint64_t qt = make_smem_desc(sbase + k3OffQTail + t * k3QTail, 16, 256, kLayoutSW32);
        const uint64_t kt = make_smem_desc(sbase + k3OffKTail + stage * k3KTail, 16, 256, kLayoutSW32);
        umma_ss(d, qt, kt, idesc_qk, 1u);
      };
      auto issue_pv = [&](int t, int stage, bool first) {
        // main (d 0..63) and wide tail (d 64..127) atoms of the stage are LBO apart
        const uint64_t vd = make_smem_desc(sbase + k3OffVMain + stage * k3VMain,
                                           (k3OffVTail - k3OffVMain) + stage * (k3VTail - k3VMain), 1024, kLayoutSW128);
        const uint32_t pt = tmem_base + t * k3TileCols;
        const uint32_t om = pt + k3ColO, lm = pt + k3ColL;
#pragma unroll
        for (int k = 0; k < k3Sub / 16; ++k) umma_ts(om, pt + 8 * k, vd + (uint64_t)(k * (2048 >> 4)), idesc_pv, (first && k == 0) ? 0u : 1u);
        const uint64_t od = make_smem_desc(sbase + k3OffOnes, 16, 1024, kLayoutSW128);
#pragma unroll
        for (int k = 0; k < k3Sub / 16; ++k) umma_ts(lm, pt + 8 * k, od, idesc_l, (first && k == 0) ? 0u : 1u);
      };

      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      for (int t = 0; t < n_tiles; ++t) {
        issue_qk(t, 0);
        umma_commit(&s_full[t]);
      }
      umma_commit(&k_empty[0]);
      for (int n = 0; n < n_sub; ++n) {
        const int stage = n % k3Stages;
        mbar_wait(&v_full[stage], (n / k3Stages) & 1);
        const bool has_next = n + 1 < n_sub;
        const int nstage = (n + 1) % k3Stages;
        if (has_next) mbar_wait(&k_full[nstage], ((n + 1) / k3Stages) & 1);
        for (int t = 0; t < n_tiles; ++t) {
          mbar_wait(&p_full[t], n & 1);
          tc_fence_after();
          issue_pv(t, stage, n == 0);
          umma_commit(&pv_done[t]);
          if (has_next) {
            issue_qk(t, nstage);
            umma_commit(&s_full[t]);
          }
        }
        umma_commit(&v_empty[stage]);
        if (has_next) umma_commit(&k_empty[nstage]);
      }
      umma_commit(o_full);
    }
  } else if (warp >= 4) {
    // ================================================================ softmax + epilogue (one thread per query row)
    const int t = (warp - 4) >> 2;
    const int qd = warp & 3;
    const int row_in_tile = qd * 32 + lane;
    const int qrow = q0 + t * k3TileQ + row_in_tile;
    const uint32_t lane_sel = static_cast<uint32_t>(qd * 32) << 16;
    const uint32_t t_s = tmem_base + t * k3TileCols + lane_sel;
    const uint32_t t_o = t_s + k3ColO, t_l = t_s + k3ColL;
    const float sl2 = p.scale_log2;
    const uint64_t sl2x2 = f32x2(sl2, sl2);
    const bool active = t < n_tiles;
    float m_ref = -INFINITY;

    auto softmax_sub = [&](const int n, auto masked_tag) {
      uint32_t va[32], vb[32];
      mbar_wait(&s_full[t], n & 1);
      tc_fence_after();
      tmem_ld_32x32b_x32_nowait(t_s, va);
      tmem_ld_32x32b_x32_nowait(t_s + 32, vb);
      tmem_ld_wait_x32(va);
      tmem_ld_wait_x32(vb);
      if constexpr (decltype(masked_tag)::value) {
        const int rem = kv_len - n * k3Sub;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i >= rem) va[i] = 0xff800000u;
          if (32 + i >= rem) vb[i] = 0xff800000u;
        }
      }
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        mx0 = fmax3(mx0, __uint_as_float(va[4 * i]), __uint_as_float(va[4 * i + 1]));
        mx1 = fmax3(mx1, __uint_as_float(va[4 * i + 2]), __uint_as_float(va[4 * i + 3]));
        mx2 = fmax3(mx2, __uint_as_float(vb[4 * i]), __uint_as_float(vb[4 * i + 1]));
        mx3 = fmax3(mx3, __uint_as_float(vb[4 * i + 2]), __uint_as_float(vb[4 * i + 3]));
      }
      const float m_new = fmaxf(fmax3(mx0, mx1, mx2), fmaxf(mx3, m_ref));
      // lazy rescale (see attn_sm100.cu): keep the old reference max unless it is stale by more than 2^8; warp-uniform
      const bool stale = (m_new - m_ref) * sl2 > 8.0f;
      if (__any_sync(0xffffffffu, stale)) {
        const float factor = (m_ref == -INFINITY) ? 0.f : fast_exp2((m_ref - m_new) * sl2);
        if (n > 0) {
          mbar_wait(&pv_done[t], (n - 1) & 1);
          tc_fence_after();
          for (int piece = 0; piece < 10; ++piece) {
            const uint32_t ta = piece < 9 ? t_o + 8 * piece : t_l;
            uint32_t o[8];
            tmem_ld_32x32b_x8(ta, o);
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
            tmem_st_32x32b_x8(ta, o);
          }
        }
        m_ref = m_new;
      }
      const float nm = -m_ref * sl2;
      const uint64_t nm2 = f32x2(nm, nm);
      uint32_t pk[32];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const uint64_t x = fma2(f32x2(__uint_as_float(va[2 * i]), __uint_as_float(va[2 * i + 1])), sl2x2, nm2);
        float x0, x1;
        f32x2_split(x, x0, x1);
        pk[i] = pack_bf16x2(fast_exp2(x0), fast_exp2(x1));
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const uint64_t x = fma2(f32x2(__uint_as_float(vb[2 * i]), __uint_as_float(vb[2 * i + 1])), sl2x2, nm2);
        float x0, x1;
        f32x2_split(x, x0, x1);
        pk[16 + i] = pack_bf16x2(fast_exp2(x0), fast_exp2(x1));
      }
      tmem_st_32x32b_x32(t_s, pk);                 // P (bf16, 64 keys = 32 packed columns) over the S columns it came from
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[t]);
    };
    if (active) {
      for (int n = 0; n + 1 < n_sub; ++n) softmax_sub(n, std::false_type{});
      if (n_sub > 0) {
        if (kv_len % k3Sub != 0) softmax_sub(n_sub - 1, std::true_type{});
        else softmax_sub(n_sub - 1, std::false_type{});
      }
    }

    // ---- epilogue: O / row_sum -> bf16 -> out[(b*Nq + qrow), h*72 .. h*72+71]
    float row_sum = 0.f;
    if (active && n_sub > 0) {
      mbar_wait(o_full, 0);
      tc_fence_after();
      uint32_t l[8];
      tmem_ld_32x32b_x8(t_l, l);
      row_sum = __uint_as_float(l[0]);
    }
    const float inv = row_sum > 0.f ? 1.0f / row_sum : 0.f;
    const bool row_ok = active && qrow < p.Nq;
    if (p.lse != nullptr && row_ok)
      p.lse[((size_t)b * p.H + h) * p.Nq + qrow] = row_sum > 0.f ? fmaf(m_ref, sl2, log2f(row_sum)) : 0.f;
    uint4* d4 = reinterpret_cast<uint4*>(p.out + (size_t)(b * p.Nq + (row_ok ? qrow : 0)) * p.ldo + h * k3D);
#pragma unroll
    for (int piece = 0; piece < 2; ++piece) {
      uint32_t o[32];
      if (active && n_sub > 0) {
        tmem_ld_32x32b_x32(t_o + 32 * piece, o);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = 0u;
      }
      if (row_ok) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          d4[4 * piece + c] = make_uint4(pack_bf16x2(__uint_as_float(o[8 * c]) * inv, __uint_as_float(o[8 * c + 1]) * inv),
                                         pack_bf16x2(__uint_as_float(o[8 * c + 2]) * inv, __uint_as_float(o[8 * c + 3]) * inv),
                                         pack_bf16x2(__uint_as_float(o[8 * c + 4]) * inv, __uint_as_float(o[8 * c + 5]) * inv),
                                         pack_bf16x2(__uint_as_float(o[8 * c + 6]) * inv, __uint_as_float(o[8 * c + 7]) * inv));
        }
      }
    }
    {
      uint32_t o1[8];
      if (active && n_sub > 0) {
        tmem_ld_32x32b_x8(t_o + 64, o1);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) o1[i] = 0u;
      }
      if (row_ok) {
        d4[8] = make_uint4(pack_bf16x2(__uint_as_float(o1[0]) * inv, __uint_as_float(o1[1]) * inv),
                           pack_bf16x2(__uint_as_float(o1[2]) * inv, __uint_as_float(o1[3]) * inv),
                           pack_bf16x2(__uint_as_float(o1[4]) * inv, __uint_as_float(o1[5]) * inv),
                           pack_bf16x2(__uint_as_float(o1[6]) * inv, __uint_as_float(o1[7]) * inv));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

static int make_qkv_maps3(CUtensorMap* main_map, CUtensorMap* tail_map, const void* base, int H, long long rows, long long s_row,
                          long long s_head, int box_rows, bool wide_tail) {
  uint64_t dims[3] = {(uint64_t)k3D, (uint64_t)H, (uint64_t)rows};
  uint64_t str[2] = {(uint64_t)s_head * 2, (uint64_t)s_row * 2};
  uint32_t box_main[3] = {64, 1, (uint32_t)box_rows};
  int rc = make_tmap_bf16(main_map, base, 3, dims, str, box_main, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  if (wide_tail) return make_tmap_bf16(tail_map, base, 3, dims, str, box_main, CU_TENSOR_MAP_SWIZZLE_128B);
  uint32_t box_tail[3] = {16, 1, (uint32_t)box_rows};
  return make_tmap_bf16(tail_map, base, 3, dims, str, box_tail, CU_TENSOR_MAP_SWIZZLE_32B);
}

// Called by pxa_flash_attn_d72_bf16 (attn_sm100.cu) after argument validation when the three-tile variant is selected.
int flash_attn_d72_x3_launch(const PxaAttnArgs& a, cudaStream_t stream) {
  CUtensorMap qm, qt, km, kt, vm, vt;
  int rc;
  if ((rc = make_qkv_maps3(&qm, &qt, a.q, a.H, (long long)a.B * a.Nq, a.q_sn, a.q_sh, k3TileQ, false))) return rc;
  if ((rc = make_qkv_maps3(&km, &kt, a.k, a.H, a.kv_rows, a.k_sn, a.k_sh, k3Sub, false))) return rc;
  if ((rc = make_qkv_maps3(&vm, &vt, a.v, a.H, a.kv_rows, a.v_sn, a.v_sh, k3Sub, true))) return rc;
  Attn3Params p;
  p.out = reinterpret_cast<__nv_bfloat16*>(a.out);
  p.lse = a.lse;
  p.kv_len = a.kv_len;
  p.kv_off = a.kv_off;
  p.B = a.B; p.H = a.H; p.Nq = a.Nq; p.Nk = a.Nk; p.ldo = a.ldo;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  p.reverse_batch = a.reverse_batch ? 1 : 0;
  PXA_CHECK_CUDA(cudaFuncSetAttribute(flash_attn_d72_x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, k3Smem));
  dim3 grid((a.Nq + k3Tiles * k3TileQ - 1) / (k3Tiles * k3TileQ), a.H, a.B);
  flash_attn_d72_x3_kernel<<<grid, k3Threads, k3Smem, stream>>>(qm, qt, km, kt, vm, vt, p);
  launch_counter()++;
  PXA_CHECK_CUDA(cudaGetLastError());
  return PXA_OK;
}

}  // namespace pxa
