// pxa_flash_attn_d72_bf16: softmax(Q K^T * scale) V for head_dim 72 on tcgen05 tensor cores (sm_100a).
//
// One work item = one (sample, head) x 256 query rows, processed as two 128-row tiles A / B.  The grid is PERSISTENT (one
// CTA per SM walks the items w = blockIdx.x, blockIdx.x + gridDim.x, ...): barrier set-up, the TMEM allocation and the CTA launch
// happen once, and the producer / MMA warps run ahead into the next item (Q, the first K / V stages and the first two Q K^T)
// while the softmax threads still write the current item's output -- the ~6 us of per-CTA prologue + epilogue that the
// one-CTA-per-item launch exposed 14 times per SM (12 % of the self-attention at 4096 keys, half of the 300-key
// cross-attention) shrink to the output store.  `variant = 2` launches one CTA per item (the round-1 behaviour).  384 threads:
//   warp 0      TMA producer: Q per item, then K / V stages of 128 keys into two kKVStages-deep smem rings
//   warp 1      MMA issuer (one elected thread): S = Q K^T (SS), O += P V (TS: P read from TMEM)
//   warp 2      TMEM allocator (512 columns: S_A | S_B | O_A | O_B)
//   warps 4-11  softmax, one thread per query row (tile A: warps 4-7, tile B: warps 8-11; 3 warps per SM sub-partition
//               leave 168 registers per thread).  Online softmax in fp32 with lazy rescaling of O (only when the
//               running max grows by > 2^8), P written back to TMEM as bf16 over the S columns it came from, final
//               O / rowsum -> bf16 -> global.
//
// The unit of work is a 64-key SUB-BLOCK (half a TMA stage).  The 128 S columns of a tile are two 64-column halves used
// as a double buffer: while the softmax threads work on S[half h] of sub-block n, Q K^T of sub-block n+1 is already in
// (or on its way into) the other half, and as soon as P(n) is published the tensor pipe runs P(n) V and then Q K^T of
// sub-block n+2 into half h.  The MMA round trip (P published -> P V -> next Q K^T -> S ready, ~900 cycles) therefore
// overlaps the exp2 work of the other half instead of sitting between two exp2 sections of the same tile, which is what
// bounded the earlier one-S-buffer-per-tile version of this kernel (profiles/r1_attn_variants.txt: 3900 cycles per 128
// keys, 40 % of them with the MUFU pipe idle).
//
// exp2 runs on packed fp32 pairs (FFMA2 / FADD2 / FMNMX3 halve the issue slots of the scale, row-sum and row-max steps)
// and a fraction of the pairs takes a polynomial exp2 on the FMA pipe instead of MUFU (see PXA_POLY_OF8).
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

// Experiment switch (tools/attn_variants.sh builds and times the alternatives; default = the fastest measured).
#ifndef PXA_SUM_MMA
// Row sums of P on the tensor pipe (P x ones, N = 16) instead of 32 FADD2 per thread per sub-block.  In isolation (clocks not
// power-capped) the tensor-pipe form is ~5 % faster (profiles/r1_attn_variants2.txt), but inside the c3 step the board sits at
// its 1 kW cap and the four extra MMAs per sub-block and tile cost more than the FADD2s: in-step attention 685 vs 650 TFLOP/s,
// step 60.88 vs 62.0 ms (profiles/r2_instep_ab.txt).  The step is what is judged: default 0.
#define PXA_SUM_MMA 0
#endif
#ifndef PXA_PREFETCH_S
#define PXA_PREFETCH_S 1      // read S of sub-block n+1 from TMEM behind the exp2 work of sub-block n
#endif
#ifndef PXA_POLY_OF8
#define PXA_POLY_OF8 0        // how many of every 8 score pairs take the polynomial exp2 (FMA pipe) instead of MUFU
#endif

namespace pxa {

constexpr int kAttnThreads = 384;   // warps 0..3: TMA, MMA, TMEM alloc, idle; warps 4..11: softmax (3 warps / sub-partition -> 168 regs)
constexpr int kD = 72;
constexpr int kTileQ = 128;
constexpr int kTileKV = 128;            // keys per TMA stage
constexpr int kSub = 64;                // keys per MMA / softmax sub-block (half a stage)
#ifndef PXA_BULK_OUT
#define PXA_BULK_OUT 1        // output rows leave through smem + one TMA bulk copy per row (see kOffOut); 0: direct stores
#endif
#ifndef PXA_KV_STAGES
#define PXA_KV_STAGES (PXA_BULK_OUT ? 2 : 3)   // 2 x 128 keys in flight = 5.6 k cycles of prefetch distance; 3 do not fit next to kOffOut
#endif
constexpr int kKVStages = PXA_KV_STAGES;
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
constexpr int kOffOnes = kOffVTail + kKVStages * kVTailBytes;      // 2 KB of bf16 1.0: B operand of the row-sum MMA
// Output staging (PXA_BULK_OUT): the softmax threads of a tile park their finished rows (72 bf16 = 144 B each) here and ONE
// elected thread hands the whole 128 x 72 tile to the TMA engine as a single tensor store -- the 5-9 store instructions per row of
// the direct path go through the LSU, where 32 lanes writing 32 different rows cost 32 transactions per instruction: 20 us of an
// 816 us self-attention call and 17 of a 79 us cross-attention call (profiles/r2_attn_epilogue.txt; one 144-byte bulk copy per
// ROW was measured too: slower than the direct stores).  2 tiles x 128 rows x 144 B = 36 KB, paid for with the third K / V stage.
constexpr int kOutRowBytes = kD * 2;
constexpr int kOffOut = kOffOnes + 2048;
constexpr int kOutTileBytes = kTileQ * kOutRowBytes;
constexpr int kOffBars = kOffOut + (PXA_BULK_OUT ? 2 * kOutTileBytes : 0);
constexpr int kAttnSmem = kOffBars + 256 + 1024;                    // + alignment slack

// TMEM columns
constexpr uint32_t kColS = 0;       // S_A at 0, S_B at 128, each two 64-column halves; P (bf16) over the first 32 columns of its half
constexpr uint32_t kColO = 256;     // O_A at 256 (main 64 + tail 16), O_B at 384
constexpr uint32_t kColL = 336;     // L_A at 336, L_B at 464: row sums of P (16 identical fp32 columns, PXA_SUM_MMA)

struct AttnParams {
  __nv_bfloat16* out;
  float* lse;           // optional [B, H, Nq]: log2-domain log-sum-exp of the scaled scores (training: consumed by the backward)
  const int* kv_len;
  const int* kv_off;
  int B, H, Nq, Nk, ldo;
  int nx;               // 256-row work items per (sample, head)
  float scale_log2;
  int reverse_batch;    // CTAs take the samples from the last to the first (L2 reuse of the freshly written qkv rows)
  long long* trace;     // debug only (NULL in production): cycle stamps of CTA (0,0,0), see PXA_TRACE
  int item_trace;       // debug only: stamp item-level events of CTA 0 instead (PXA_ITRACE; variant 5 + debug_trace)
  int trace_item;       // debug only: the work item (of CTA 0) whose sub-block stamps PXA_TRACE records
  int stagger;          // persistent grid: tile B starts every item half an exp2 section behind tile A
  int wide_stores;      // out and its row stride are 32-byte aligned: 256-bit output stores
  int tile_store;       // persistent grid: output tiles leave through smem + one TMA tensor store (drains under the next item)
};

constexpr int kTraceMax = 512;
// Slot `who` (0..15 softmax warps, 16 = MMA thread, 17 = TMA thread) appends clock64() stamps.
#define PXA_TRACE(who, cnt)                                                                        \
  do {                                                                                             \
    if (tracing && (cnt) < kTraceMax) p.trace[(who) * kTraceMax + (cnt)++] = clock64();            \
  } while (0)

// Item-level stamps of CTA 0 (slot 0 = first softmax warp, 16 = MMA thread, 17 = TMA thread): 8 per item with keys, 64 items.
#define PXA_ITRACE(slot, k)                                                                                   \
  do {                                                                                                        \
    if (p.item_trace && blockIdx.x == 0 && it < 64) p.trace[(slot) * kTraceMax + 8 * it + (k)] = clock64();  \
  } while (0)

// kSplitP = "fp32_attention" grade P V (PxaAttnArgs.p_precision = 1; PixArt_blocks.py:145-147 keeps q / k / v and therefore P in
// fp32): the softmax threads publish P as TWO bf16 terms, P_hi = bf16(P) over the first 32 columns of the S half (as always) and
// P_lo = bf16(P - P_hi) over the other 32 columns -- free once the scores are in registers -- and the MMA thread accumulates
// P_hi V + P_lo V (same V descriptor, A operand 32 TMEM columns further).  P then carries 16 mantissa bits (2^-17 relative
// instead of 2^-8), V is bf16-valued in the reference too (the qkv GEMM output), the products are exact and the accumulation is
// fp32: the result is the fp32 attention of the same bf16 q / k / v up to summation order.  Tensor work per 64-key sub-block pair
// goes 640 -> 960 cycles, still under the 1024-cycle exp2 floor, and no smem or TMEM is added (a kind::tf32 P V would need fp32
// V tiles -- twice the smem traffic -- for the same tensor-pipe time).
template <bool kSplitP>
__global__ void __launch_bounds__(kAttnThreads, 1)
flash_attn_d72_kernel(const __grid_constant__ CUtensorMap tm_q_main, const __grid_constant__ CUtensorMap tm_q_tail,
                      const __grid_constant__ CUtensorMap tm_k_main, const __grid_constant__ CUtensorMap tm_k_tail,
                      const __grid_constant__ CUtensorMap tm_v_main, const __grid_constant__ CUtensorMap tm_v_tail,
                      const __grid_constant__ CUtensorMap tm_out, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBars);
  uint64_t* q_full = bars;                    // [1]
  uint64_t* k_full = bars + 1;                // [kKVStages]
  uint64_t* k_empty = k_full + kKVStages;     // [kKVStages]
  uint64_t* v_full = k_empty + kKVStages;     // [kKVStages]
  uint64_t* v_empty = v_full + kKVStages;     // [kKVStages]
  uint64_t* s_full = v_empty + kKVStages;     // [2 tiles][2 halves]  MMA -> softmax: S of a 64-key sub-block ready
  uint64_t* p_full = s_full + 4;              // [2][2]  softmax -> MMA: P written over that S half
  uint64_t* pv_done = p_full + 4;             // [2]     MMA -> softmax: P V of a sub-block done (lazy-rescale path only)
  uint64_t* o_full = pv_done + 2;             // [1]     MMA -> softmax: all P V of the item done
  uint64_t* q_empty = o_full + 1;             // [1]     MMA -> TMA: the item's last Q K^T has read the Q tiles
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_empty + 1);

  const int warp = warp_idx_sync();
  const int lane = threadIdx.x & 31;
  // work item w -> (sample, head, first query row); x fastest, so the CTAs running side by side share K / V through L2
  const int total_items = p.nx * p.H * p.B;
  struct Item {
    int b, h, q0, kv_len, kv_row0, n_blocks, n_sub;
  };
  auto decode = [&](int w) {
    Item im;
    const int x = w % p.nx, yz = w / p.nx;
    im.h = yz % p.H;
    const int z = yz / p.H;
    im.b = p.reverse_batch ? p.B - 1 - z : z;
    im.q0 = x * (2 * kTileQ);
    int kv_len = p.kv_len ? p.kv_len[im.b] : p.Nk;
    im.kv_len = min(max(kv_len, 0), p.Nk);
    im.kv_row0 = p.kv_off ? p.kv_off[im.b] : im.b * p.Nk;
    im.n_blocks = (im.kv_len + kTileKV - 1) / kTileKV;     // 128-key TMA stages
    im.n_sub = (im.kv_len + kSub - 1) / kSub;              // 64-key sub-blocks (the MMA / softmax unit)
    return im;
  };
  int tcnt = 0;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tm_q_main); prefetch_tmap(&tm_q_tail);
    prefetch_tmap(&tm_k_main); prefetch_tmap(&tm_k_tail);
    prefetch_tmap(&tm_v_main); prefetch_tmap(&tm_v_tail);
    prefetch_tmap(&tm_out);
    mbar_init(q_full, 1);
    for (int s = 0; s < kKVStages; ++s) {
      mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 128);
    }
    mbar_init(&pv_done[0], 1);
    mbar_init(&pv_done[1], 1);
    mbar_init(o_full, 1);
    mbar_init(q_empty, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
#if PXA_SUM_MMA
  if (warp == 3) {
    uint4* ones = reinterpret_cast<uint4*>(smem + kOffOnes);
    for (int i = lane; i < 2048 / 16; i += 32) ones[i] = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
    fence_proxy_async_smem();                    // generic-proxy stores -> visible to the tensor core's smem reads
  }
#endif
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================================================ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0, it = 0;                  // `it` counts the items with keys (the others touch no barrier)
      for (int w = blockIdx.x; w < total_items; w += gridDim.x) {
        const Item im = decode(w);
        if (im.n_blocks == 0) continue;
        const bool tracing = p.trace != nullptr && w == p.trace_item && !p.item_trace;
        const int h = im.h, qrow = im.b * p.Nq + im.q0;
        PXA_ITRACE(17, 0);
        mbar_wait(q_empty, (it & 1) ^ 1);          // the previous item's Q K^T are done with the Q tiles
        PXA_ITRACE(17, 1);
        mbar_arrive_expect_tx(q_full, 2 * kTileBytes);
        for (int t = 0; t < 2; ++t) {
          // NB: a tile whose rows run past this sample's Nq reads the next sample's rows (finite garbage, never
          // stored) or TMA zero fill past the end of the tensor.
          tma_load_3d(smem + kOffQMain + t * kMainBytes, &tm_q_main, q_full, 0, h, qrow + t * kTileQ, kEvictFirst);
          tma_load_3d(smem + kOffQTail + t * kTailBytes, &tm_q_tail, q_full, 64, h, qrow + t * kTileQ, kEvictFirst);
        }
        for (int j = 0; j < im.n_blocks; ++j) {
          const int krow = im.kv_row0 + j * kTileKV;
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
        PXA_ITRACE(17, 2);
        ++it;
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc_qk = make_idesc_bf16(128, kSub, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, 80, 0, 1);        // V is MN-major; N = 80 = d 0..79
      const uint32_t sbase = smem_u32(smem);

      // S[t][hh] = Q_t K_hh^T for the 64 keys of half hh of a stage: 4 K-steps from the 128B-swizzled main buffers + 1
      // from the 32B-swizzled tails.  K rows 64 hh .. 64 hh + 63 start 64 rows into the (K-major) tiles.
      auto issue_qk = [&](int t, int stage, int hh) {
        const uint64_t qd = make_smem_desc(sbase + kOffQMain + t * kMainBytes, 16, 1024, kLayoutSW128);
        const uint64_t kd = make_smem_desc(sbase + kOffKMain + stage * kMainBytes + hh * (kSub * 128), 16, 1024, kLayoutSW128);
        const uint32_t d = tmem_base + kColS + t * 128 + hh * kSub;
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(d, qd + 2 * k, kd + 2 * k, idesc_qk, k != 0 ? 1u : 0u);
        const uint64_t qt = make_smem_desc(sbase + kOffQTail + t * kTailBytes, 16, 256, kLayoutSW32);
        const uint64_t kt = make_smem_desc(sbase + kOffKTail + stage * kTailBytes + hh * (kSub * 32), 16, 256, kLayoutSW32);
        umma_ss(d, qt, kt, idesc_qk, 1u);
      };
      // O_t += P[t][hh] V_hh : ONE N=80 MMA per 16 keys.  V is MN-major; its d extent spans two 64-element swizzle atoms:
      // the main tile (d 0..63) and the tail tile (d 64..127, of which 64..71 are data and the rest TMA zero fill), LBO
      // apart.  (Two MMAs, N=64 + N=16, cost about twice the issue + pipeline overhead of one N=80 MMA.)  P from TMEM.
      auto issue_pv = [&](int t, int stage, int hh, bool first) {
        const uint64_t vd = make_smem_desc(sbase + kOffVMain + stage * kMainBytes, kOffVTail - kOffVMain, 1024, kLayoutSW128);
        const uint32_t pt = tmem_base + kColS + t * 128 + hh * kSub;
        const uint32_t om = tmem_base + kColO + t * 128;
#pragma unroll
        for (int k = 0; k < kSub / 16; ++k) {
          const uint32_t acc = (first && k == 0) ? 0u : 1u;
          umma_ts(om, pt + 8 * k, vd + (uint64_t)((4 * hh + k) * (2048 >> 4)), idesc_pv, acc);
        }
        if constexpr (kSplitP) {                     // + P_lo V: the low-order bf16 term of P sits 32 columns behind P_hi
#pragma unroll
          for (int k = 0; k < kSub / 16; ++k)
            umma_ts(om, pt + 32 + 8 * k, vd + (uint64_t)((4 * hh + k) * (2048 >> 4)), idesc_pv, 1u);
        }
#if PXA_SUM_MMA
        // L_t += P[t][hh] 1 : the row sums of the (bf16-rounded) P that P V multiplies, for free on the tensor pipe.  The
        // B operand is all ones, so any valid descriptor over the 2 KB ones region will do.
        constexpr uint32_t idesc_l = make_idesc_bf16(128, 16, 0, 0);
        const uint64_t od = make_smem_desc(sbase + kOffOnes, 16, 1024, kLayoutSW128);
        const uint32_t lm = tmem_base + kColL + t * 128;
#pragma unroll
        for (int k = 0; k < kSub / 16; ++k) umma_ts(lm, pt + 8 * k, od, idesc_l, (first && k == 0) ? 0u : 1u);
        if constexpr (kSplitP) {
#pragma unroll
          for (int k = 0; k < kSub / 16; ++k) umma_ts(lm, pt + 32 + 8 * k, od, idesc_l, 1u);
        }
#endif
      };

      // Running barrier counts across items: K / V blocks loaded so far (ring stage / phase of block j of this item), completed
      // uses of S half 0 / 1 (the same for both tiles), items with keys.
      uint32_t it = 0, kbase = 0, hb0 = 0, hb1 = 0;
      for (int w = blockIdx.x; w < total_items; w += gridDim.x) {
        const Item im = decode(w);
        if (im.n_blocks == 0) continue;
        const bool tracing = p.trace != nullptr && w == p.trace_item && !p.item_trace;
        const int n_sub = im.n_sub;
        PXA_ITRACE(16, 0);
        auto kstage = [&](int j) { return (int)((kbase + j) % kKVStages); };
        auto kphase = [&](int j) { return ((kbase + j) / kKVStages) & 1u; };
        // prologue: the first two sub-blocks' S for both tiles (the two S halves of a tile are a double buffer).  The S halves
        // are free: every P of the previous item was consumed by a P V issued earlier on the (in-order) tensor pipe.
        mbar_wait(q_full, it & 1);
        mbar_wait(&k_full[kstage(0)], kphase(0));
        PXA_ITRACE(16, 1);
        tc_fence_after();
        for (int hh = 0; hh < 2 && hh < n_sub; ++hh) {
          for (int t = 0; t < 2; ++t) {
            issue_qk(t, kstage(0), hh);
            umma_commit(&s_full[2 * t + hh]);
          }
        }
        umma_commit(&k_empty[kstage(0)]);
        if (n_sub <= 2) umma_commit(q_empty);
        PXA_ITRACE(16, 2);

        for (int n = 0; n < n_sub; ++n) {
          const int j = n >> 1, hh = n & 1;
          const int stage = kstage(j);
          if (hh == 0) mbar_wait(&v_full[stage], kphase(j));
          const int nn = n + 2;                      // the sub-block that reuses this S half
          const bool has_next = nn < n_sub;
          const int nstage = kstage(j + 1);
          if (has_next && hh == 0) mbar_wait(&k_full[nstage], kphase(j + 1));
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            PXA_TRACE(16, tcnt);                     // [4n+2t]   start waiting for P[t]
            mbar_wait(&p_full[2 * t + hh], ((hh ? hb1 : hb0) + j) & 1);
            PXA_TRACE(16, tcnt);                     // [4n+2t+1] P[t] ready
            tc_fence_after();
            // n == 0 overwrites O / L of the previous item: its softmax threads published this P after their output stores
            issue_pv(t, stage, hh, n == 0);
            umma_commit(&pv_done[t]);
            if (has_next) {
              issue_qk(t, nstage, hh);
              umma_commit(&s_full[2 * t + hh]);
            }
          }
          if (hh == 1 || n + 1 == n_sub) umma_commit(&v_empty[stage]);            // last P V reading this V stage
          if (has_next && (hh == 1 || nn + 1 == n_sub)) umma_commit(&k_empty[nstage]);   // last Q K^T reading that K stage
          if (has_next && nn + 1 == n_sub) umma_commit(q_empty);                   // ... and the item's last read of Q
        }
        umma_commit(o_full);
        PXA_ITRACE(16, 3);
        kbase += im.n_blocks;
        hb0 += (n_sub + 1) >> 1;
        hb1 += n_sub >> 1;
        ++it;
      }
    }
  } else if (warp >= 4) {
    // ================================================================ softmax + epilogue (one thread per query row)
    const int sw = warp - 4;                       // 0..7
    const int t = sw >> 2;                         // tile 0 (A) / 1 (B)
    const int qd = warp & 3;                       // TMEM sub-partition a warp may access (= warp % 4)
    const int row_in_tile = qd * 32 + lane;
    const uint32_t lane_sel = static_cast<uint32_t>(qd * 32) << 16;
    const uint32_t t_s = tmem_base + kColS + t * 128 + lane_sel;   // two 64-column S halves; P (bf16) over the first 32 of each
    const uint32_t t_o = tmem_base + kColO + t * 128 + lane_sel;   // O columns 0..79 (d 0..71 + 8 zero pad)
    const float sl2 = p.scale_log2;
    const uint64_t sl2x2 = f32x2(sl2, sl2);

    const uint32_t t_l = tmem_base + kColL + t * 128 + lane_sel;
    // running barrier counts across items (see the MMA issuer): uses of S half 0 / 1, P V commits, items with keys
    uint32_t it = 0, hb0 = 0, hb1 = 0, pvb = 0;

    for (int w = blockIdx.x; w < total_items; w += gridDim.x) {
    const Item im = decode(w);
    const bool tracing = p.trace != nullptr && w == p.trace_item && lane == 0 && !p.item_trace;
    const bool itr = sw == 0 && lane == 0 && im.n_blocks > 0;
    const int b = im.b, h = im.h, kv_len = im.kv_len, n_blocks = im.n_blocks, n_sub = im.n_sub;
    if (itr) PXA_ITRACE(0, 0);
    const int qrow = im.q0 + t * kTileQ + row_in_tile;          // query index within the sample
    float m_ref = -INFINITY;     // reference max used in the exponent (raw S units)
    uint64_t sa = f32x2(0.f, 0.f), sb = f32x2(0.f, 0.f);        // running row sum, four partial lanes (!PXA_SUM_MMA)

    uint32_t va[32], vb[32];       // the 64 scores of this row in the current sub-block
    if (n_sub > 0 && t == 1 && p.stagger) named_bar_sync(1, 256);           // tile B: half an exp2 section behind tile A (see softmax_sub)
#if PXA_PREFETCH_S
    if (n_sub > 0) {
      mbar_wait(&s_full[2 * t], hb0 & 1);
      tc_fence_after();
      tmem_ld_32x32b_x32_nowait(t_s, va);
      tmem_ld_32x32b_x32_nowait(t_s + 32, vb);
      tmem_ld_wait_x32(va);
      tmem_ld_wait_x32(vb);
    }
#endif
    if (itr) PXA_ITRACE(0, 1);
    auto softmax_sub = [&](const int n, auto masked_tag) {
      [[maybe_unused]] const int j = n >> 1;
      const int hh = n & 1;
      const uint32_t ts = t_s + hh * kSub;
      PXA_TRACE(sw, tcnt);                          // [7n+0] start waiting for S
#if !PXA_PREFETCH_S
      mbar_wait(&s_full[2 * t + hh], ((hh ? hb1 : hb0) + j) & 1);
      PXA_TRACE(sw, tcnt);                          // [7n+1] S ready
      tc_fence_after();
      tmem_ld_32x32b_x32_nowait(ts, va);
      tmem_ld_32x32b_x32_nowait(ts + 32, vb);
      tmem_ld_wait_x32(va);
      tmem_ld_wait_x32(vb);
#else
      PXA_TRACE(sw, tcnt);                          // [7n+1] (S was fetched during the previous sub-block)
#endif
      PXA_TRACE(sw, tcnt);                          // [7n+2] S in registers
      // Only the last sub-block of a sample can be partial.  The masking selects are compiled into a separate copy of
      // the body: if-converted into the common path they would cost 64 extra issue slots per thread per sub-block.
      if constexpr (decltype(masked_tag)::value) {
        const int rem = kv_len - n * kSub;         // valid keys in this sub-block
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i >= rem) va[i] = 0xff800000u;       // -inf
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
      PXA_TRACE(sw, tcnt);                          // [7n+3] row max known
      // Lazy rescale: keep the old reference max unless it is stale by more than 2^8 (P stays <= 256, exact in the
      // fp32 accumulators).  The decision is warp-uniform because the TMEM round trip below is warp-collective.
      const bool stale = (m_new - m_ref) * sl2 > 8.0f;          // true on the first sub-block (m_ref = -inf)
      if (__any_sync(0xffffffffu, stale)) {
        const float factor = (m_ref == -INFINITY) ? 0.f : fast_exp2((m_ref - m_new) * sl2);
        if (n > 0) {
          // O may only be touched between P V (n-1) and P V (n): the latter waits for this thread's P, the former is
          // awaited here (the S double buffer lets Q K^T run ahead, so "S ready" no longer implies "P V done").
          mbar_wait(&pv_done[t], (pvb + n - 1) & 1);
          tc_fence_after();
          // d 0..71 in 9 pieces of 8 columns (rare path: a short register footprint matters more than TMEM round
          // trips); the pad columns 72..79 hold zeros and need no scaling.
          for (int piece = 0; piece < 9 + PXA_SUM_MMA; ++piece) {
            const uint32_t ta = piece < 9 ? t_o + 8 * piece : t_l;       // the row-sum columns scale like O
            uint32_t o[8];
            tmem_ld_32x32b_x8(ta, o);
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
            tmem_st_32x32b_x8(ta, o);
          }
        }
        const uint64_t f2 = f32x2(factor, factor);
        sa = mul2(sa, f2);
        sb = mul2(sb, f2);
        m_ref = m_new;
      }
      const float nm = -m_ref * sl2;
      const uint64_t nm2 = f32x2(nm, nm);
      [[maybe_unused]] const float a_sat = -sl2 * (1.0f / 128.0f), b_sat = (8.0f - nm) * (1.0f / 128.0f);   // fma_exp2_x2
      // exp2 on packed fp32 pairs (FFMA2 / FADD2: one issue slot for two lane-ops).  PXA_POLY_OF8 of every 8 pairs take
      // the polynomial exp2 (FMA pipe) instead of MUFU (16 results/clk/SM), so that both pipes and the issue port end up
      // about equally loaded.
      uint32_t pk[32];
      [[maybe_unused]] uint32_t pl[16];             // kSplitP: P_lo words of 32 keys at a time
      // one pair of scores -> exp2 -> running packed sum + bf16x2 P word; `pair` (0..31) selects MUFU or polynomial
      auto exp_pair = [&](float s0, float s1, int pair, [[maybe_unused]] uint64_t& acc, uint32_t& packed,
                          [[maybe_unused]] uint32_t& packed_lo) {
        [[maybe_unused]] uint64_t e;
        float e0, e1;
        if (PXA_POLY_OF8 > 0 && ((pair * PXA_POLY_OF8) & 7) < PXA_POLY_OF8) {
          fma_exp2_x2(s0, s1, a_sat, b_sat, e0, e1);              // FMA-pipe exp2: keeps the MUFU pipe for the other pairs
#if !PXA_SUM_MMA
          e = f32x2(e0, e1);
#endif
        } else {
          const uint64_t x = fma2(f32x2(s0, s1), sl2x2, nm2);
          float x0, x1;
          f32x2_split(x, x0, x1);
          e0 = fast_exp2(x0);
          e1 = fast_exp2(x1);
          e = f32x2(e0, e1);
        }
#if !PXA_SUM_MMA
        acc = add2(acc, e);
#endif
        packed = pack_bf16x2(e0, e1);
        if constexpr (kSplitP) packed_lo = pack_bf16x2(e0 - bf16_lo(packed), e1 - bf16_hi(packed));   // exact differences
      };
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        exp_pair(__uint_as_float(va[2 * i]), __uint_as_float(va[2 * i + 1]), i, sa, pk[i], pl[i]);
        exp_pair(__uint_as_float(va[2 * i + 2]), __uint_as_float(va[2 * i + 3]), i + 1, sb, pk[i + 1], pl[i + 1]);
      }
      // P_lo of keys 0..31 -> columns 32..47 of this S half (their scores were read into vb long ago; completion is covered by
      // the tcgen05.wait::st in front of the P barrier)
      if constexpr (kSplitP) tmem_st_32x32b_x16(ts + 32, pl);
      // Persistent grid: at an item boundary both tiles find their first S complete and would start IN PHASE -- and stay
      // there (1650 instead of 1395 cycles per sub-block, tools/attn_itrace.py): the kernel's good operating point is tile B
      // about half an exp2 section behind tile A, which is what a fresh CTA falls into because S_A(0) is issued first.  So tile B
      // starts every item when tile A is half-way through its first exp2 section.
      if (n == 0 && t == 0 && p.stagger) named_bar_arrive(1, 256);
#if PXA_PREFETCH_S
      // S of the next sub-block sits in the other S half (normally complete long ago): fetch it into the registers the
      // first 32 scores have just left, behind the second half of this sub-block's exp2 work.
      const bool has_next = n + 1 < n_sub;
      const uint32_t tn = t_s + (hh ^ 1) * kSub;
      if (has_next) {
        mbar_wait(&s_full[2 * t + (hh ^ 1)], ((hh ? hb0 : hb1) + ((n + 1) >> 1)) & 1);
        tc_fence_after();
        tmem_ld_32x32b_x32_nowait(tn, va);
      }
#endif
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        exp_pair(__uint_as_float(vb[2 * i]), __uint_as_float(vb[2 * i + 1]), 16 + i, sa, pk[16 + i], pl[i]);
        exp_pair(__uint_as_float(vb[2 * i + 2]), __uint_as_float(vb[2 * i + 3]), 16 + i + 1, sb, pk[16 + i + 1], pl[i + 1]);
      }
      if constexpr (kSplitP) tmem_st_32x32b_x16(ts + 48, pl);
#if PXA_PREFETCH_S
      if (has_next) tmem_ld_32x32b_x32_nowait(tn + 32, vb);
#endif
      if (tracing) {   // debug only: pin the end of the exp2 section for the cycle trace
        asm volatile("" ::"r"(pk[0]), "r"(pk[7]), "r"(pk[15]), "r"(pk[23]), "r"(pk[31]), "l"(sa), "l"(sb) : "memory");
      }
      PXA_TRACE(sw, tcnt);                          // [7n+4] exp2 section done
      // P (bf16, 64 keys = 32 packed columns) over the first 32 columns of this S half
      tmem_st_32x32b_x32(ts, pk);
      tmem_st_wait();
      PXA_TRACE(sw, tcnt);                          // [7n+5] P in TMEM
      tc_fence_before();
      mbar_arrive(&p_full[2 * t + hh]);
#if PXA_PREFETCH_S
      if (has_next) {
        tmem_ld_wait_x32(va);
        tmem_ld_wait_x32(vb);
      }
#endif
      PXA_TRACE(sw, tcnt);                          // [7n+6] P published (and the next S in registers)
    };
    for (int n = 0; n + 1 < n_sub; ++n) softmax_sub(n, std::false_type{});
    if (n_sub > 0) {
      if (kv_len % kSub != 0) softmax_sub(n_sub - 1, std::true_type{});
      else softmax_sub(n_sub - 1, std::false_type{});
    }
    if (itr) PXA_ITRACE(0, 2);
    float row_sum;
    {
      float s0, s1, s2, s3;
      f32x2_split(sa, s0, s1);
      f32x2_split(sb, s2, s3);
      row_sum = (s0 + s1) + (s2 + s3);
    }

    // ---- epilogue: O / row_sum -> bf16 -> out[(b*Nq + qrow), h*72 .. h*72+71]
    if (n_blocks > 0) {
      mbar_wait(o_full, it & 1);
      if (itr) PXA_ITRACE(0, 3);
      tc_fence_after();
#if PXA_SUM_MMA
      uint32_t l[8];
      tmem_ld_32x32b_x8(t_l, l);
      row_sum = __uint_as_float(l[0]);
#endif
    }
    const float inv = row_sum > 0.f ? 1.0f / row_sum : 0.f;      // n_blocks == 0 (no keys): zeros
    if (p.lse != nullptr && qrow < p.Nq)                          // P = exp2(S * scale_log2 - lse) in the backward
      p.lse[((size_t)b * p.H + h) * p.Nq + qrow] = row_sum > 0.f ? fmaf(m_ref, sl2, log2f(row_sum)) : 0.f;
    // The row's 72 outputs (144 B) as 36 packed words.  Stores: 32 lanes write 32 different rows, so every store instruction
    // costs one LSU sector operation per lane whatever its width -- 256-bit stores (4 x 32 B + 1 x 16 B per row, the 16-byte
    // piece first when h is odd: h * 144 B is then only 16-byte aligned) need 5 per row instead of the 9 of 128-bit stores.
    const bool row_ok = qrow < p.Nq;
    uint32_t ow[36];
    if (n_blocks > 0) {
      uint32_t o8[8];
      tmem_ld_32x32b_x32_nowait(t_o, va);
      tmem_ld_32x32b_x32_nowait(t_o + 32, vb);
      tmem_ld_32x32b_x8(t_o + 64, o8);             // tcgen05.wait::ld covers all three
      tmem_ld_wait_x32(va);
      tmem_ld_wait_x32(vb);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        ow[i] = pack_bf16x2(__uint_as_float(va[2 * i]) * inv, __uint_as_float(va[2 * i + 1]) * inv);
        ow[16 + i] = pack_bf16x2(__uint_as_float(vb[2 * i]) * inv, __uint_as_float(vb[2 * i + 1]) * inv);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) ow[32 + i] = pack_bf16x2(__uint_as_float(o8[2 * i]) * inv, __uint_as_float(o8[2 * i + 1]) * inv);
    } else {
#pragma unroll
      for (int i = 0; i < 36; ++i) ow[i] = 0u;
    }
#if PXA_BULK_OUT
    if (p.tile_store) {
      uint8_t* stile = smem + kOffOut + t * kOutTileBytes;
      const bool issuer = qd == 0;                 // first warp of the tile: one elected lane talks to the TMA engine
      if (issuer) {
        if (elect_one()) tma_store_wait_read<0>(); // the previous item's store has read the staging tile
      }
      named_bar_sync(2 + t, 128);
      uint4* srow = reinterpret_cast<uint4*>(stile + row_in_tile * kOutRowBytes);
#pragma unroll
      for (int c = 0; c < 9; ++c) srow[c] = make_uint4(ow[4 * c], ow[4 * c + 1], ow[4 * c + 2], ow[4 * c + 3]);
      fence_proxy_async_smem();                    // the generic-proxy stores above -> visible to the TMA engine's read
      named_bar_sync(2 + t, 128);
      if (issuer && im.q0 + t * kTileQ < p.Nq) {   // rows >= Nq of a partial tile are clipped by the tensor map
        if (elect_one()) {
          tma_store_4d(&tm_out, stile, 0, h, im.q0 + t * kTileQ, b);
          tma_store_commit();
        }
      }
    }
#endif
#ifdef PXA_DEBUG_NO_STORE     // experiment only: how much of the epilogue is the output stores (results are wrong)
    if (row_ok && ow[0] == 0x12345678u) {
#else
    if (row_ok && !(PXA_BULK_OUT && p.tile_store)) {
#endif
      __nv_bfloat16* dst = p.out + (size_t)(b * p.Nq + qrow) * p.ldo + h * kD;
      if (p.wide_stores) {
        const int lead = (h & 1) ? 4 : 0;          // words in front of the first 32-byte boundary
        if (lead) *reinterpret_cast<uint4*>(dst) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (lead)
            st_global_v8(dst + 8 + 16 * c, ow[4 + 8 * c], ow[5 + 8 * c], ow[6 + 8 * c], ow[7 + 8 * c], ow[8 + 8 * c], ow[9 + 8 * c],
                         ow[10 + 8 * c], ow[11 + 8 * c]);
          else
            st_global_v8(dst + 16 * c, ow[8 * c], ow[8 * c + 1], ow[8 * c + 2], ow[8 * c + 3], ow[8 * c + 4], ow[8 * c + 5],
                         ow[8 * c + 6], ow[8 * c + 7]);
        }
        if (!lead) *reinterpret_cast<uint4*>(dst + 64) = make_uint4(ow[32], ow[33], ow[34], ow[35]);
      } else {
#pragma unroll
        for (int c = 0; c < 9; ++c)
          reinterpret_cast<uint4*>(dst)[c] = make_uint4(ow[4 * c], ow[4 * c + 1], ow[4 * c + 2], ow[4 * c + 3]);
      }
    }
    if (itr) PXA_ITRACE(0, 4);
    if (n_blocks > 0) {
      hb0 += (n_sub + 1) >> 1;
      hb1 += n_sub >> 1;
      pvb += n_sub;
      ++it;
    }
    }   // items
#if PXA_BULK_OUT
    // The staging tile must outlive the store's READ of it; the write to global memory completes on its own (waiting for it too
    // cost the one-CTA-per-item launch 4.5 us per CTA: 880 instead of 816 us at 4096 keys).
    tma_store_wait_read<0>();
#endif
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

int flash_attn_d72_x3_launch(const PxaAttnArgs& a, cudaStream_t stream);   // attn3_sm100.cu

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
  if (a.variant != 0 && (a.variant < 2 || a.variant > 7)) return fail(PXA_ERR_ARG, "variant must be 0 or 2..7");
  if (a.p_precision != 0 && a.p_precision != 1) return fail(PXA_ERR_ARG, "p_precision must be 0 (bf16 P) or 1 (bf16 hi + lo P)");
  if (a.variant == 3 && a.p_precision) return fail(PXA_ERR_ARG, "variant 3 has no p_precision = 1 form");
  if (a.variant == 3 && !a.debug_trace) return flash_attn_d72_x3_launch(a, reinterpret_cast<cudaStream_t>(stream));
  CUtensorMap qm, qt, km, kt, vm, vt;
  int rc;
  if ((rc = make_qkv_maps(&qm, &qt, a.q, a.H, (long long)a.B * a.Nq, a.q_sn, a.q_sh))) return rc;
  if ((rc = make_qkv_maps(&km, &kt, a.k, a.H, a.kv_rows, a.k_sn, a.k_sh))) return rc;
  if ((rc = make_qkv_maps(&vm, &vt, a.v, a.H, a.kv_rows, a.v_sn, a.v_sh, true))) return rc;
  AttnParams p;
  p.out = reinterpret_cast<__nv_bfloat16*>(a.out);
  p.lse = a.lse;
  p.kv_len = a.kv_len;
  p.kv_off = a.kv_off;
  p.B = a.B; p.H = a.H; p.Nq = a.Nq; p.Nk = a.Nk; p.ldo = a.ldo;
  p.nx = (a.Nq + 2 * kTileQ - 1) / (2 * kTileQ);
  p.scale_log2 = a.scale * 1.4426950408889634f;
  p.reverse_batch = a.reverse_batch ? 1 : 0;
  p.trace = reinterpret_cast<long long*>(a.debug_trace);
  p.wide_stores = ((reinterpret_cast<uintptr_t>(a.out) & 31) == 0 && (a.ldo & 15) == 0) ? 1 : 0;
  p.item_trace = (a.debug_trace && a.variant == 5) ? 1 : 0;
  auto kernel = a.p_precision ? flash_attn_d72_kernel<true> : flash_attn_d72_kernel<false>;
  PXA_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem));
  // variant 0 / 4: persistent grid, one CTA per SM walks the items; 2: one CTA per item (round-1 behaviour).  In-step A/B under
  // the power cap (profiles/r2_instep_ab.txt, r2_attn_epilogue.txt): persistent wins for every key count on the boxes of sessions
  // 19-29 (-0.9 ... -1.5 % of the c3 step; isolated 794 vs 819 us at 4096 keys, 67 vs 107 us for the 300-token cross-attention).
  const long long items = (long long)p.nx * a.H * a.B;
  long long grid = items;
  const bool persistent = a.variant != 2;
  if (persistent && grid > device_info().sms) grid = device_info().sms;
  p.stagger = (grid < items && a.variant != 5) ? 1 : 0;      // variant 5 (experiments): persistent without the stagger
  // Output tiles through smem + one TMA tensor store: pays for short key sets, where the epilogue is a large part of an item
  // (cross-attention 80 -> 67 us); at 4096 keys the direct 256-bit stores are faster (794 vs 821 us).  With one CTA per item the
  // CTA could not retire before the TMA engine has read the staging tile (870 vs 816 us): never there.
  p.tile_store = (PXA_BULK_OUT && grid < items && ((a.Nk <= 1024 && a.variant != 6) || a.variant == 7)) ? 1 : 0;
  p.trace_item = grid < items ? (int)(2 * grid) : 0;          // sub-block trace: CTA 0's third item when persistent
  CUtensorMap om;
  {
    // out as [B][Nq][H][72]: one store per 128-row tile of one head; rows >= Nq (partial last tile) are clipped
    uint64_t dims[4] = {(uint64_t)kD, (uint64_t)a.H, (uint64_t)a.Nq, (uint64_t)a.B};
    uint64_t str[3] = {(uint64_t)kD * 2, (uint64_t)a.ldo * 2, (uint64_t)a.Nq * a.ldo * 2};
    uint32_t box[4] = {(uint32_t)kD, 1, (uint32_t)kTileQ, 1};
    if ((rc = make_tmap_bf16(&om, a.out, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE))) return rc;
  }
  kernel<<<(unsigned)grid, kAttnThreads, kAttnSmem, reinterpret_cast<cudaStream_t>(stream)>>>(qm, qt, km, kt, vm, vt, om, p);
  launch_counter()++;
  PXA_CHECK_CUDA(cudaGetLastError());
  return PXA_OK;
}
