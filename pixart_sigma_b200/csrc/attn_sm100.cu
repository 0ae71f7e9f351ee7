// pxa_flash_attn_d72_bf16: softmax(Q K^T * scale) V for head_dim 72 on tcgen05 tensor cores (sm_100a).
//
// One CTA = one (sample, head) x 256 query rows, processed as two 128-row tiles A / B that ping-pong between the
// tensor pipe and the softmax warps (while the softmax warps of A run exp2 on S_A, the tensor pipe does P_B V and
// the next Q_B K^T).  640 threads:
//   warp 0      TMA producer: Q once, then K / V blocks of 128 keys into two 3-deep smem rings
//   warp 1      MMA issuer (one thread): S = Q K^T (SS), O += P V (TS: P read from TMEM)
//   warp 2      TMEM allocator (512 columns: S_A | S_B | O_A | O_B)
//   warps 4-19  softmax: per tile two warpgroups, "lo" owns keys 0..63 and "hi" keys 64..127 of every block, so a
//               query row is shared by a thread pair (64 S values each; row max exchanged through smem once per
//               block).  4 softmax warps per SM sub-partition keep the MUFU fed while others wait on TMEM / barriers.
//               Online softmax in fp32 with lazy rescaling of O (only when the running max grows by > 2^8), P written
//               back to TMEM as bf16 over each half's own S columns, final O / rowsum -> bf16 -> global
//
// head_dim 72 is not a multiple of the 64-element swizzle atom, nor of UMMA K=16 / N%16: each Q/K/V tile is staged
// as a "main" part (d 0..63, 128B swizzle) plus a "tail" part (d 64..79, 32B swizzle) whose d 72..79 are zero-filled
// by TMA out-of-bounds handling (the tensor map's innermost extent is 72).  QK^T = 4 main K-steps + 1 tail K-step;
// P V = one N=80 MMA per 16 keys, V consumed MN-major straight from its natural [key, d] layout (two swizzle atoms).
//
// Algorithmic work: 4 * Nq * Nk * 72 FLOP per (sample, head); the MUFU (exp2) pipe, not the tensor pipe, is the
// tighter bound at head_dim 72: 128x128 exp2 per tile-block = 1024 cycles/SM vs 640 cycles of MMA.
#include <type_traits>

#include "host_common.cuh"
#include "ptx.cuh"

// Experiment switches (tools/attn_variants.sh builds and times the combinations; defaults = the fastest measured).
#ifndef PXA_EXP_CHUNKED
#define PXA_EXP_CHUNKED 1     // cut the exp2 section into four dependency-chained 16-element chunks
#endif
#ifndef PXA_EXP_POLY
#define PXA_EXP_POLY 1        // 1 in 4 exponentials on the FMA pipe (poly_exp2)
#endif
#ifndef PXA_STAGGER
#define PXA_STAGGER 0         // delay tile B's first exp2 section until tile A's first one is done
#endif

namespace pxa {

constexpr int kAttnThreads = 640;
constexpr int kD = 72;
constexpr int kTileQ = 128;
constexpr int kTileKV = 128;
constexpr int kKVStages = 3;
constexpr int kMainBytes = 128 * 128;   // 128 rows x 64 bf16
constexpr int kTailBytes = 128 * 32;    // 128 rows x 16 bf16
constexpr int kTileBytes = kMainBytes + kTailBytes;

// smem carve-up (offsets from the 1024-aligned base)
constexpr int kOffQMain = 0;                                       // 2 tiles
constexpr int kOffKMain = kOffQMain + 2 * kMainBytes;              // kKVStages
constexpr int kOffVMain = kOffKMain + kKVStages * kMainBytes;
constexpr int kOffQTail = kOffVMain + kKVStages * kMainBytes;
constexpr int kOffKTail = kOffQTail + 2 * kTailBytes;
constexpr int kOffVTail = kOffKTail + kKVStages * kTailBytes;      // V tails are full 128 B-row tiles (see below)
constexpr int kVTailBytes = kMainBytes;                            // 128 keys x 128 B: d 64..71 valid, rest zero
constexpr int kVTileBytes = kMainBytes + kVTailBytes;
constexpr int kOffBars = kOffVTail + kKVStages * kVTailBytes;
constexpr int kOffXchg = kOffBars + 256;                            // float [2 parity][2 tile][2 half][128 row]
constexpr int kAttnSmem = kOffXchg + 2 * 2 * 2 * 128 * 4 + 1024;

// TMEM columns
constexpr uint32_t kColS = 0;       // S_A at 0, S_B at 128; P (bf16) aliases S: keys 0..63 -> cols [0,32), keys 64..127 -> [64,96)
constexpr uint32_t kColO = 256;     // O_A at 256 (main 64 + tail 16), O_B at 384

struct AttnParams {
  __nv_bfloat16* out;
  const int* kv_len;
  const int* kv_off;
  int B, H, Nq, Nk, ldo;
  float scale_log2;
  long long* trace;     // debug only (NULL in production): cycle stamps of CTA (0,0,0), see PXA_TRACE
};

constexpr int kTraceMax = 512;
// Slot `who` (0..15 softmax warps, 16 = MMA thread, 17 = TMA thread) appends clock64() stamps.
#define PXA_TRACE(who, cnt)                                                                        \
  do {                                                                                             \
    if (tracing && (cnt) < kTraceMax) p.trace[(who) * kTraceMax + (cnt)++] = clock64();            \
  } while (0)

__global__ void __launch_bounds__(kAttnThreads, 1)
flash_attn_d72_kernel(const __grid_constant__ CUtensorMap tm_q_main, const __grid_constant__ CUtensorMap tm_q_tail,
                      const __grid_constant__ CUtensorMap tm_k_main, const __grid_constant__ CUtensorMap tm_k_tail,
                      const __grid_constant__ CUtensorMap tm_v_main, const __grid_constant__ CUtensorMap tm_v_tail,
                      const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBars);
  uint64_t* q_full = bars;                    // [1]
  uint64_t* k_full = bars + 1;                // [kKVStages]
  uint64_t* k_empty = k_full + kKVStages;     // [kKVStages]
  uint64_t* v_full = k_empty + kKVStages;     // [kKVStages]
  uint64_t* v_empty = v_full + kKVStages;     // [kKVStages]
  uint64_t* s_full = v_empty + kKVStages;     // [2]  MMA -> softmax (S ready; also implies previous PV done)
  uint64_t* p_full = s_full + 2;              // [2]  softmax -> MMA (P written, S consumed)
  uint64_t* o_full = p_full + 2;              // [1]  MMA -> softmax (all PV done)
  uint64_t* stagger_bar = o_full + 1;         // [1]  PXA_STAGGER experiment only
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(stagger_bar + 1);

  const int warp = warp_idx_sync();
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y;
  const int q0 = blockIdx.x * (2 * kTileQ);

  int kv_len = p.kv_len ? p.kv_len[b] : p.Nk;
  kv_len = min(max(kv_len, 0), p.Nk);
  const int kv_row0 = p.kv_off ? p.kv_off[b] : b * p.Nk;
  const int n_blocks = (kv_len + kTileKV - 1) / kTileKV;
  const bool tracing = p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0;
  int tcnt = 0;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tm_q_main); prefetch_tmap(&tm_q_tail);
    prefetch_tmap(&tm_k_main); prefetch_tmap(&tm_k_tail);
    prefetch_tmap(&tm_v_main); prefetch_tmap(&tm_v_tail);
    mbar_init(q_full, 1);
    for (int s = 0; s < kKVStages; ++s) {
      mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t], 256);
    }
    mbar_init(o_full, 1);
    mbar_init(stagger_bar, 256);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================================================ TMA producer
    if (n_blocks > 0 && elect_one()) {
      const int qrow = b * p.Nq + q0;
      mbar_arrive_expect_tx(q_full, 2 * kTileBytes);
      for (int t = 0; t < 2; ++t) {
        // NB: a tile whose rows run past this sample's Nq reads the next sample's rows (finite garbage, never
        // stored) or TMA zero fill past the end of the tensor.
        tma_load_3d(smem + kOffQMain + t * kMainBytes, &tm_q_main, q_full, 0, h, qrow + t * kTileQ, kEvictFirst);
        tma_load_3d(smem + kOffQTail + t * kTailBytes, &tm_q_tail, q_full, 64, h, qrow + t * kTileQ, kEvictFirst);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_blocks; ++j) {
        const int krow = kv_row0 + j * kTileKV;
        mbar_wait(&k_empty[stage], phase ^ 1);
        PXA_TRACE(17, tcnt);
        mbar_arrive_expect_tx(&k_full[stage], kTileBytes);
        tma_load_3d(smem + kOffKMain + stage * kMainBytes, &tm_k_main, &k_full[stage], 0, h, krow, kEvictLast);
        tma_load_3d(smem + kOffKTail + stage * kTailBytes, &tm_k_tail, &k_full[stage], 64, h, krow, kEvictLast);
        mbar_wait(&v_empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&v_full[stage], kVTileBytes);
        tma_load_3d(smem + kOffVMain + stage * kMainBytes, &tm_v_main, &v_full[stage], 0, h, krow, kEvictLast);
        tma_load_3d(smem + kOffVTail + stage * kVTailBytes, &tm_v_tail, &v_full[stage], 64, h, krow, kEvictLast);
        if (++stage == kKVStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    if (n_blocks > 0 && elect_one()) {
      constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, 80, 0, 1);        // V is MN-major; N = 80 = d 0..79
      const uint32_t sbase = smem_u32(smem);

      // S_t = Q_t K^T : 4 K-steps from the 128B-swizzled main buffers + 1 from the 32B-swizzled tails
      auto issue_qk = [&](int t, int stage) {
        const uint64_t qd = make_smem_desc(sbase + kOffQMain + t * kMainBytes, 16, 1024, kLayoutSW128);
        const uint64_t kd = make_smem_desc(sbase + kOffKMain + stage * kMainBytes, 16, 1024, kLayoutSW128);
        const uint32_t d = tmem_base + kColS + t * 128;
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(d, qd + 2 * k, kd + 2 * k, idesc_qk, k != 0 ? 1u : 0u);
        const uint64_t qt = make_smem_desc(sbase + kOffQTail + t * kTailBytes, 16, 256, kLayoutSW32);
        const uint64_t kt = make_smem_desc(sbase + kOffKTail + stage * kTailBytes, 16, 256, kLayoutSW32);
        umma_ss(d, qt, kt, idesc_qk, 1u);
      };
      // O_t += P_t V : ONE N=80 MMA per 16 keys.  V is MN-major; its d extent spans two 64-element swizzle atoms: the main
      // tile (d 0..63) and the tail tile (d 64..127, of which 64..71 are data and the rest TMA zero fill), LBO apart.
      // (Two MMAs, N=64 + N=16, cost about twice the issue + pipeline overhead of one N=80 MMA.)  P from TMEM.
      auto issue_pv = [&](int t, int stage, bool first) {
        const uint64_t vd = make_smem_desc(sbase + kOffVMain + stage * kMainBytes, kOffVTail - kOffVMain, 1024, kLayoutSW128);
        const uint32_t pt = tmem_base + kColS + t * 128;
        const uint32_t om = tmem_base + kColO + t * 128;
#pragma unroll
        for (int k = 0; k < kTileKV / 16; ++k) {
          const uint32_t acc = (first && k == 0) ? 0u : 1u;
          const uint32_t pa = pt + (k < 4 ? 8 * k : 64 + 8 * (k - 4));   // lo half at cols 0..31, hi half at 64..95
          umma_ts(om, pa, vd + (uint64_t)(k * (2048 >> 4)), idesc_pv, acc);
        }
      };

      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_qk(0, 0);
      umma_commit(&s_full[0]);
      issue_qk(1, 0);
      umma_commit(&s_full[1]);
      umma_commit(&k_empty[0]);

      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_blocks; ++j) {
        int nstage = stage + 1;
        uint32_t nphase = phase;
        if (nstage == kKVStages) { nstage = 0; nphase ^= 1; }
        const bool more = (j + 1 < n_blocks);
        mbar_wait(&v_full[stage], phase);
        // ---- tile A
        PXA_TRACE(16, tcnt);                       // [4j+0] start waiting for P_A
        mbar_wait(&p_full[0], j & 1);
        PXA_TRACE(16, tcnt);                       // [4j+1] P_A ready
        tc_fence_after();
        issue_pv(0, stage, j == 0);
        if (more) {
          mbar_wait(&k_full[nstage], nphase);
          tc_fence_after();
          issue_qk(0, nstage);
          umma_commit(&s_full[0]);
        }
        // ---- tile B
        PXA_TRACE(16, tcnt);                       // [4j+2] A's MMAs issued, start waiting for P_B
        mbar_wait(&p_full[1], j & 1);
        PXA_TRACE(16, tcnt);                       // [4j+3] P_B ready
        tc_fence_after();
        issue_pv(1, stage, j == 0);
        umma_commit(&v_empty[stage]);
        if (more) {
          issue_qk(1, nstage);
          umma_commit(&s_full[1]);
          umma_commit(&k_empty[nstage]);
        }
        stage = nstage;
        phase = nphase;
      }
      umma_commit(o_full);
    }
  } else if (warp >= 4) {
    // ================================================================ softmax + epilogue (two threads per query row)
    const int w = warp - 4;
    const int t = w >> 3;                          // tile 0 (A) / 1 (B)
    const int hf = (w >> 2) & 1;                   // 0: keys 0..63 of each block, 1: keys 64..127
    const int qd = w & 3;                          // TMEM sub-partition (= warp % 4)
    const int row_in_tile = qd * 32 + lane;
    const int qrow = q0 + t * kTileQ + row_in_tile;             // query index within the sample
    const uint32_t lane_sel = static_cast<uint32_t>(qd * 32) << 16;
    const uint32_t t_s = tmem_base + kColS + t * 128 + hf * 64 + lane_sel;   // this half's 64 S columns (P over the first 32)
    const uint32_t t_o = tmem_base + kColO + t * 128 + hf * 32 + lane_sel;   // lo: O cols 0..31, hi: O cols 32..71
    const float sl2 = p.scale_log2;
    float* xchg = reinterpret_cast<float*>(smem + kOffXchg);
    const uint32_t bar_id = 1 + t;                 // named barrier of this tile's 256 softmax threads

    float m_ref = -INFINITY;     // reference max used in the exponent (raw S units); identical in both halves
    float row_sum = 0.f;         // partial: this half's keys only

    auto softmax_block = [&](const int j, auto masked_tag) {
      PXA_TRACE(w, tcnt);                          // [7j+0] start waiting for S
      mbar_wait(&s_full[t], j & 1);
      PXA_TRACE(w, tcnt);                          // [7j+1] S ready
      tc_fence_after();
      uint32_t v0[32], v1[32];
      tmem_ld_32x32b_x32_pair(t_s, v0, t_s + 32, v1);
      PXA_TRACE(w, tcnt);                          // [7j+2] S in registers
      const int rem = kv_len - j * kTileKV - hf * 64;           // valid keys among this half's 64
      // Only the last block of a sample can be partial.  The masking selects are compiled into a separate copy of the
      // block body: if-converted into the common path they cost ~250 extra issue slots per thread per block (the
      // softmax warps are issue-bound: 4 warps per sub-partition share one issue port).
      if constexpr (decltype(masked_tag)::value) {
        const uint32_t ninf = 0xff800000u;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i >= rem) v0[i] = ninf;
          if (32 + i >= rem) v1[i] = ninf;
        }
      }
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        mx0 = fmaxf(mx0, __uint_as_float(v0[2 * i]));
        mx1 = fmaxf(mx1, __uint_as_float(v0[2 * i + 1]));
        mx2 = fmaxf(mx2, __uint_as_float(v1[2 * i]));
        mx3 = fmaxf(mx3, __uint_as_float(v1[2 * i + 1]));
      }
      const float m_half = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      float* xb = xchg + ((j & 1) * 4 + t * 2) * 128;
      xb[hf * 128 + row_in_tile] = m_half;
      named_bar_sync(bar_id, 256);
      PXA_TRACE(w, tcnt);                          // [7j+3] row max exchanged
      const float m_new = fmaxf(fmaxf(m_half, xb[(hf ^ 1) * 128 + row_in_tile]), m_ref);
      // Lazy rescale: keep the old reference max unless it is stale by more than 2^8 (P stays <= 256, exact in the
      // fp32 accumulators).  The decision is warp-uniform (the TMEM round trip below is warp-collective) and identical
      // in the partner warp of the other half, which sees the same m_new / m_ref for the same 32 rows.
      const bool stale = (m_new - m_ref) * sl2 > 8.0f;          // true on the first block (m_ref = -inf)
      if (__any_sync(0xffffffffu, stale)) {
        const float factor = (m_ref == -INFINITY) ? 0.f : fast_exp2((m_ref - m_new) * sl2);
        if (j > 0) {
          // PV of block j-1 has completed (it was issued before the QK^T that produced this S)
          uint32_t o0[32], o1[8];
          tmem_ld_32x32b_x32(t_o, o0);
#pragma unroll
          for (int i = 0; i < 32; ++i) o0[i] = __float_as_uint(__uint_as_float(o0[i]) * factor);
          tmem_st_32x32b_x32(t_o, o0);
          if (hf == 1) {                         // warp-uniform: d 64..71 (the 8 zero pad columns 72..79 need no scaling)
            tmem_ld_32x32b_x8(t_o + 32, o1);
#pragma unroll
            for (int i = 0; i < 8; ++i) o1[i] = __float_as_uint(__uint_as_float(o1[i]) * factor);
            tmem_st_32x32b_x8(t_o + 32, o1);
          }
        }
        row_sum *= factor;
        m_ref = m_new;
      }
      const float neg_m = -m_ref * sl2;
      float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
      uint32_t pk[32];
      // exp2 section.  Per 4 elements: 3 on the MUFU pipe, 1 on the FMA/ALU pipes (poly_exp2).  The work is cut into four
      // 16-element chunks chained through a register dependency, so the compiler cannot cluster all MUFU instructions
      // into one burst: warps issue in order, and when the 4 warps of a sub-partition all sit in a MUFU burst at the same
      // time the FMA pipe idles (and vice versa).  Fine-grained chunks let one warp's MUFU work overlap another's FMA work.
      float nm = neg_m;
#if PXA_STAGGER
      if (j == 0 && t == 1) {
        mbar_wait(stagger_bar, 0);
        asm volatile("" : "+f"(nm)::"memory");       // the exp2 below may not be hoisted above the wait
      }
#endif
#pragma unroll
      for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int i = 4 * c; i < 4 * c + 4; ++i) {
          const float a0 = fast_exp2(fmaf(__uint_as_float(v0[2 * i]), sl2, nm));
          const float a1 = fast_exp2(fmaf(__uint_as_float(v0[2 * i + 1]), sl2, nm));
          const float b0 = fast_exp2(fmaf(__uint_as_float(v1[2 * i]), sl2, nm));
#if PXA_EXP_POLY
          const float b1 = poly_exp2(fmaf(__uint_as_float(v1[2 * i + 1]), sl2, nm));
#else
          const float b1 = fast_exp2(fmaf(__uint_as_float(v1[2 * i + 1]), sl2, nm));
#endif
          sum0 += a0; sum1 += a1; sum2 += b0; sum3 += b1;
          pk[i] = pack_bf16x2(a0, a1);
          pk[16 + i] = pack_bf16x2(b0, b1);
        }
#if PXA_EXP_CHUNKED
        if (c < 3) asm volatile("" : "+f"(nm) : "f"(sum0), "f"(sum1), "f"(sum2), "f"(sum3));
#endif
      }
#if PXA_STAGGER
      if (j == 0 && t == 0) {
        asm volatile("" ::"r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]),
                     "r"(pk[8]), "r"(pk[9]), "r"(pk[10]), "r"(pk[11]), "r"(pk[12]), "r"(pk[13]), "r"(pk[14]), "r"(pk[15]),
                     "r"(pk[16]), "r"(pk[17]), "r"(pk[18]), "r"(pk[19]), "r"(pk[20]), "r"(pk[21]), "r"(pk[22]),
                     "r"(pk[23]), "r"(pk[24]), "r"(pk[25]), "r"(pk[26]), "r"(pk[27]), "r"(pk[28]), "r"(pk[29]),
                     "r"(pk[30]), "r"(pk[31])
                     : "memory");
        mbar_arrive(stagger_bar);
      }
#endif
      if (tracing) {   // debug only: pin the end of the exp2 section for the cycle trace
        asm volatile("" ::"r"(pk[0]), "r"(pk[3]), "r"(pk[7]), "r"(pk[11]), "r"(pk[15]), "r"(pk[16]), "r"(pk[19]), "r"(pk[23]),
                     "r"(pk[27]), "r"(pk[31]), "f"(sum0), "f"(sum1), "f"(sum2), "f"(sum3) : "memory");
      }
      PXA_TRACE(w, tcnt);                          // [7j+4] exp2 section done
      // P (bf16, this half's 64 keys = 32 packed columns) over the first 32 of this half's own S columns
      tmem_st_32x32b_x32(t_s, pk);
      tmem_st_wait();
      PXA_TRACE(w, tcnt);                          // [7j+5] P in TMEM
      tc_fence_before();
      mbar_arrive(&p_full[t]);
      PXA_TRACE(w, tcnt);                          // [7j+6] P published
      row_sum += (sum0 + sum1) + (sum2 + sum3);
    };
    for (int j = 0; j + 1 < n_blocks; ++j) softmax_block(j, std::false_type{});
    if (n_blocks > 0) {
      if (kv_len % kTileKV != 0) softmax_block(n_blocks - 1, std::true_type{});
      else softmax_block(n_blocks - 1, std::false_type{});
    }

    // ---- epilogue: O / row_sum -> bf16 -> out[(b*Nq + qrow), h*72 + hf*32 .. ]   (lo: d 0..31, hi: d 32..71)
    float* xs = xchg + (t * 2) * 128;               // parity-0 slots are free again after the last block's barrier
    named_bar_sync(bar_id, 256);                    // everybody is done reading the max exchange buffers
    xs[hf * 128 + row_in_tile] = row_sum;
    named_bar_sync(bar_id, 256);
    const float total = row_sum + xs[(hf ^ 1) * 128 + row_in_tile];
    uint32_t o0[32], o1[8];
    if (n_blocks > 0) {
      mbar_wait(o_full, 0);
      tc_fence_after();
      tmem_ld_32x32b_x32(t_o, o0);
      if (hf == 1) tmem_ld_32x32b_x8(t_o + 32, o1);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) o0[i] = 0u;
#pragma unroll
      for (int i = 0; i < 8; ++i) o1[i] = 0u;
    }
    if (qrow < p.Nq) {
      const float inv = total > 0.f ? 1.0f / total : 0.f;
      __nv_bfloat16* dst = p.out + (size_t)(b * p.Nq + qrow) * p.ldo + h * kD + hf * 32;
      uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        d4[c] = make_uint4(pack_bf16x2(__uint_as_float(o0[8 * c]) * inv, __uint_as_float(o0[8 * c + 1]) * inv),
                           pack_bf16x2(__uint_as_float(o0[8 * c + 2]) * inv, __uint_as_float(o0[8 * c + 3]) * inv),
                           pack_bf16x2(__uint_as_float(o0[8 * c + 4]) * inv, __uint_as_float(o0[8 * c + 5]) * inv),
                           pack_bf16x2(__uint_as_float(o0[8 * c + 6]) * inv, __uint_as_float(o0[8 * c + 7]) * inv));
      }
      if (hf == 1) {
        d4[4] = make_uint4(pack_bf16x2(__uint_as_float(o1[0]) * inv, __uint_as_float(o1[1]) * inv),
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

static int make_qkv_maps(CUtensorMap* main_map, CUtensorMap* tail_map, const void* base, int H, long long rows,
                         long long s_row, long long s_head, bool wide_tail = false) {
  uint64_t dims[3] = {(uint64_t)kD, (uint64_t)H, (uint64_t)rows};
  uint64_t str[2] = {(uint64_t)s_head * 2, (uint64_t)s_row * 2};
  uint32_t box_main[3] = {64, 1, 128};
  uint32_t box_tail[3] = {16, 1, 128};
  int rc = make_tmap_bf16(main_map, base, 3, dims, str, box_main, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  if (wide_tail) {
    // V: the tail is a full 64-wide 128B-swizzled tile (d 64..127, zero-filled past 71) so that main + tail form one
    // two-atom MN-major operand for a single N=80 MMA
    uint32_t box_wide[3] = {64, 1, 128};
    return make_tmap_bf16(tail_map, base, 3, dims, str, box_wide, CU_TENSOR_MAP_SWIZZLE_128B);
  }
  return make_tmap_bf16(tail_map, base, 3, dims, str, box_tail, CU_TENSOR_MAP_SWIZZLE_32B);
}

}  // namespace pxa

extern "C" int pxa_flash_attn_d72_bf16(const PxaAttnArgs* args, void* stream) {
  using namespace pxa;
  if (!args) return fail(PXA_ERR_ARG, "null args");
  const PxaAttnArgs& a = *args;
  if (!a.q || !a.k || !a.v || !a.out) return fail(PXA_ERR_ARG, "null q / k / v / out");
  if (a.B <= 0 || a.H <= 0 || a.Nq <= 0 || a.Nk < 0) return fail(PXA_ERR_ARG, "bad B/H/Nq/Nk");
  if (a.kv_rows <= 0) return fail(PXA_ERR_ARG, "kv_rows must be > 0");
  if ((a.q_sn & 7) || (a.q_sh & 7) || (a.k_sn & 7) || (a.k_sh & 7) || (a.v_sn & 7) || (a.v_sh & 7) || (a.ldo & 7))
    return fail(PXA_ERR_ALIGN, "strides must be multiples of 8 elements");
  if (reinterpret_cast<uintptr_t>(a.out) & 15) return fail(PXA_ERR_ALIGN, "out must be 16-byte aligned");
  if (a.H * kD > a.ldo) return fail(PXA_ERR_ARG, "ldo smaller than H*72");
  PXA_REQUIRE_SM100();
  CUtensorMap qm, qt, km, kt, vm, vt;
  int rc;
  if ((rc = make_qkv_maps(&qm, &qt, a.q, a.H, (long long)a.B * a.Nq, a.q_sn, a.q_sh))) return rc;
  if ((rc = make_qkv_maps(&km, &kt, a.k, a.H, a.kv_rows, a.k_sn, a.k_sh))) return rc;
  if ((rc = make_qkv_maps(&vm, &vt, a.v, a.H, a.kv_rows, a.v_sn, a.v_sh, true))) return rc;
  AttnParams p;
  p.out = reinterpret_cast<__nv_bfloat16*>(a.out);
  p.kv_len = a.kv_len;
  p.kv_off = a.kv_off;
  p.B = a.B; p.H = a.H; p.Nq = a.Nq; p.Nk = a.Nk; p.ldo = a.ldo;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  p.trace = reinterpret_cast<long long*>(a.debug_trace);
  PXA_CHECK_CUDA(cudaFuncSetAttribute(flash_attn_d72_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem));
  dim3 grid((a.Nq + 2 * kTileQ - 1) / (2 * kTileQ), a.H, a.B);
  flash_attn_d72_kernel<<<grid, kAttnThreads, kAttnSmem, reinterpret_cast<cudaStream_t>(stream)>>>(qm, qt, km, kt, vm,
                                                                                                  vt, p);
  launch_counter()++;
  PXA_CHECK_CUDA(cudaGetLastError());
  return PXA_OK;
}
