// pxa_flash_attn_d72_bwd_bf16: backward of softmax(Q K^T * scale) V for head_dim 72 on tcgen05 tensor cores (sm_100a).
// Replaces the autograd backward of xformers.ops.memory_efficient_attention at PixArt_blocks.py:52-53,153 when
// train_scripts/train.py:197 runs the loss backward through the blocks.
//
// Flash-attention-2 style recomputation from (Q, K, V, dO, lse, delta); nothing N x N ever reaches HBM.  One kernel
// template, two launches -- each is "a 128-row STATIONARY tile against a stream of 128-row tiles":
//   dKV pass (kDKV = true):  stationary K_j, V_j;   stream Q_i, dO_i over all query tiles of the sample
//       S'  = K_j Q_i^T            (keys on TMEM lanes, queries on columns)      P'  = exp2(S' * c - lse[q])
//       dP' = V_j dO_i^T                                                         dS' = P' * (dP' - delta[q])
//       dV_j += P' dO_i,   dK_j += dS' Q_i                                       (dK scaled by `scale` at the end)
//   dQ pass  (kDKV = false): stationary Q_i, dO_i;  stream K_j, V_j over the sample's key tiles
//       S' = Q_i K_j^T,  dP' = dO_i V_j^T,  dS' as above (statistics per lane),  dQ_i += dS' K_j
// The dQ pass recomputes S' and dP' instead of exchanging dS through smem / atomics: 7 instead of 5 tile products per
// (i, j) pair, in exchange for a deterministic kernel built only from the operand forms the forward kernel already
// uses: score MMAs are SS with both operands K-major (d contiguous; 4 K-steps from the 64-wide main tile + 1 from the
// tail tile), gradient MMAs are TS (A = P' / dS' read from TMEM as bf16, written by the softmax threads over the fp32
// columns they came from) with the streamed tile consumed MN-major in its natural [row, d] layout (N = 80: main + tail).
//
// 384 threads: warp 0 TMA producer, warp 1 MMA issuer (one elected thread), warp 2 TMEM allocator, warps 4-11 the
// elementwise stage (8 or 16 warps): thread = (TMEM lane, column slice) -> 64 / kBEw columns of S' and dP' per sub-block.
// The unit of work is a 64-row SUB-BLOCK of the stream (one TMA stage, 4-deep ring).  S' and dP' each have two 64-column
// TMEM buffers used as a double buffer, exactly like the forward kernel: while the elementwise threads turn S'(n), dP'(n)
// into P'(n), dS'(n), the tensor pipe already computes the scores of sub-block n+1 into the other buffer, and as soon as
// P'(n) / dS'(n) are published it runs the gradient MMAs of n and the scores of n+2.  (The first version of this kernel
// used one 128-column buffer: score MMAs, elementwise stage and gradient MMAs were strictly sequential.)
// head_dim 72: every tile is staged as a main part (d 0..63) and a 64-wide tail part (d 64..127, zero-filled past 71 by
// TMA out-of-bounds handling), both 128B-swizzled, so one staging serves the K-major and the MN-major use.
//
// Algorithmic work: dKV pass 8 * Nq * Nk * 72 FLOP, dQ pass 6 * Nq * Nk * 72 FLOP per (sample, head) (model FLOPs of the
// attention backward: 10 * Nq * Nk * 72 -- the difference is the recomputation).  Any Nq / Nk (partial last tiles are masked).
#include <type_traits>

#include "host_common.cuh"
#include "ptx.cuh"

namespace pxa {

// Elementwise warps per TMEM sub-partition (2 or 4): thread = (TMEM lane, 64 / kBEw column slice).  4 warps per
// sub-partition hide the TMEM load / MUFU / TMEM store latencies of the slice-serial elementwise stage much better than 2.
#ifndef PXA_BWD_EW
#define PXA_BWD_EW 4
#endif
// kTmaStats (template parameter of the kernel; PXA_BWD_TMA_STATS=0 forces it off): the TMA producer brings lse / delta of
// every streamed query sub-block into the stage (two 1-D bulk copies on the stage's full barrier) instead of the
// elementwise warps loading them and handing them over through a per-sub-block 512-thread named barrier.  Measured on
// B200 (round 2, 4 x 16 x 4096): 1.249 ms vs 1.354 ms.  Needs Nq % 4 == 0 (16-byte aligned copies); other token counts
// take the kTmaStats = false instantiation.
#ifndef PXA_BWD_TMA_STATS
#define PXA_BWD_TMA_STATS 1
#endif
// (Round 2 also measured split-phase TMEM loads -- S'(n+1) / dP'(n+1) requested behind the TMEM store of n: 1.498 vs 1.249 ms,
// removed.)
#ifndef PXA_BWD_PERSISTENT
#define PXA_BWD_PERSISTENT 0  // 1: one CTA per SM walks the work items (same kernel, grid = SM count).  Measured at 4 x 16 x 4096:
                              // 1.267 vs 1.226 ms, c5 step 122.2 vs 121.7 ms -- unlike the forward kernel, whose short cross-attention
                              // items were half set-up time, every backward item streams >= 64 sub-blocks: one CTA per item stays.
#endif
#ifndef PXA_BWD_BULK_OUT
#define PXA_BWD_BULK_OUT 0    // 1: gradient rows through smem + one TMA bulk copy per ROW -- the forward measured this form slower
                              // than direct stores (profiles/r2_attn_epilogue.txt); default: 256-bit stores
#endif
constexpr int kBEw = PXA_BWD_EW;
constexpr int kBEwThreads = 128 * kBEw;        // elementwise threads
constexpr int kBCols = 64 / kBEw;      // score columns per thread per sub-block
constexpr int kBwdThreads = 128 + kBEwThreads;
constexpr int kBT = 128;                       // rows per tile (both stationary and streamed)
constexpr int kBSub = 64;                      // streamed rows per sub-block (= TMA stage)
constexpr int kBMain = 128 * 128;              // stationary: 128 rows x 64 bf16
constexpr int kBTile = 2 * kBMain;             // main + 64-wide tail
constexpr int kBYMain = kBSub * 128;           // streamed: 64 rows x 64 bf16
constexpr int kBYTile = 2 * kBYMain;           // main + 64-wide tail
constexpr int kBStages = 4;
constexpr int kBOffX1 = 0;
constexpr int kBOffX2 = kBTile;
constexpr int kBOffY = 2 * kBTile;             // stage s: Y1 at kBOffY + s * 2 * kBYTile, Y2 right behind it
constexpr int kBOffStat = kBOffY + kBStages * 2 * kBYTile;  // 2 buffers x (lse[64] | delta[64]) fp32 (one per stage with TMA stats)
constexpr int kBOffBars = kBOffStat + kBStages * 512;
constexpr int kBwdSmem = kBOffBars + 256 + 1024;            // + alignment slack

constexpr uint32_t kBColS = 0;       // S'  two 64-column buffers; each thread's bf16 P' over the first half of its own column slice
constexpr uint32_t kBColDP = 128;    // dP' (same, dS')
constexpr uint32_t kBColAcc2 = 256;  // dK (dKV pass) / dQ (dQ pass): 80 columns
constexpr uint32_t kBColAcc1 = 384;  // dV (dKV pass): 80 columns

// size-overloaded TMEM accessors so that the elementwise stage is written once for 32 or 16 columns per thread
PXA_DEVICE void ld_scores(uint32_t t0, uint32_t (&a)[32], uint32_t t1, uint32_t (&b)[32]) { tmem_ld_32x32b_x32_pair(t0, a, t1, b); }
PXA_DEVICE void ld_scores(uint32_t t0, uint32_t (&a)[16], uint32_t t1, uint32_t (&b)[16]) { tmem_ld_32x32b_x16_pair(t0, a, t1, b); }
PXA_DEVICE void ld_scores_nowait(uint32_t t0, uint32_t (&a)[32], uint32_t t1, uint32_t (&b)[32]) {
  tmem_ld_32x32b_x32_nowait(t0, a);
  tmem_ld_32x32b_x32_nowait(t1, b);
}
PXA_DEVICE void ld_scores_nowait(uint32_t t0, uint32_t (&a)[16], uint32_t t1, uint32_t (&b)[16]) {
  tmem_ld_32x32b_x16_nowait(t0, a);
  tmem_ld_32x32b_x16_nowait(t1, b);
}
PXA_DEVICE void ld_scores_wait(uint32_t (&a)[32], uint32_t (&b)[32]) { tmem_ld_wait_x32(a); tmem_ld_wait_x32(b); }
PXA_DEVICE void ld_scores_wait(uint32_t (&a)[16], uint32_t (&b)[16]) { tmem_ld_wait_x16(a); tmem_ld_wait_x16(b); }
PXA_DEVICE void st_packed(uint32_t t, const uint32_t (&r)[16]) { tmem_st_32x32b_x16(t, r); }
PXA_DEVICE void st_packed(uint32_t t, const uint32_t (&r)[8]) { tmem_st_32x32b_x8(t, r); }

struct AttnBwdParams {
  const float* lse;        // [B, H, Nq] log2-domain log-sum-exp written by the forward kernel
  const float* delta;      // [B, H, Nq] rowsum(dO * O)
  __nv_bfloat16* d2;       // dK (dKV pass) / dQ (dQ pass)
  __nv_bfloat16* d1;       // dV (dKV pass)
  long long d2_sn, d2_sh, d1_sn, d1_sh;
  const int* kv_len;
  const int* kv_off;
  int B, H, Nq, Nk;
  int nx;                  // 128-row stationary tiles per (sample, head): work item w -> (x = w % nx, h, b)
  float scale, scale_log2;
};

template <bool kDKV, bool kTmaStats>
__global__ void __launch_bounds__(kBwdThreads, 1)
flash_attn_d72_bwd_kernel(const __grid_constant__ CUtensorMap tm_x1m, const __grid_constant__ CUtensorMap tm_x1t,
                          const __grid_constant__ CUtensorMap tm_x2m, const __grid_constant__ CUtensorMap tm_x2t,
                          const __grid_constant__ CUtensorMap tm_y1m, const __grid_constant__ CUtensorMap tm_y1t,
                          const __grid_constant__ CUtensorMap tm_y2m, const __grid_constant__ CUtensorMap tm_y2t,
                          const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kBOffBars);
  uint64_t* x_full = bars;                     // [1]
  uint64_t* y_full = bars + 1;                 // [kBStages]  TMA -> MMA
  uint64_t* y_empty = y_full + kBStages;       // [kBStages]  MMA -> TMA
  uint64_t* s_full = y_empty + kBStages;       // [2]  MMA -> elementwise: S' and dP' of a sub-block are in buffer hh
  uint64_t* p_full = s_full + 2;               // [2]  elementwise -> MMA: P' / dS' written over buffer hh (kBEwThreads arrivals)
  uint64_t* acc_full = p_full + 2;             // [1]  MMA -> epilogue
  uint64_t* x_empty = acc_full + 1;            // [1]  MMA -> TMA: the item's last score MMA has read the stationary tiles
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(x_empty + 1);
  float* stat = reinterpret_cast<float*>(smem + kBOffStat);     // [2][128]: -lse[64] | delta[64] of a streamed q sub-block

  const int warp = warp_idx_sync();
  const int lane = threadIdx.x & 31;

  // PERSISTENT grid (like the forward kernel): one CTA per SM walks the work items w = blockIdx.x, + gridDim.x, ...; barrier
  // phases, the stream ring position and the TMEM allocation carry over, the producer reloads the stationary tiles as soon as the
  // item's last score MMA has been issued, and the next item's first scores are computed while the gradient rows are stored.
  const int total_items = p.nx * p.H * p.B;
  struct Item {
    int b, h, t0, kv_len, x_row0, y_row0, n_iter;
    bool skip;
  };
  auto decode = [&](int w) {
    Item im;
    const int x = w % p.nx, yz = w / p.nx;
    im.h = yz % p.H;
    im.b = yz / p.H;
    im.t0 = x * kBT;
    int kv_len = p.kv_len ? p.kv_len[im.b] : p.Nk;
    im.kv_len = min(max(kv_len, 0), p.Nk);
    const int kv_row0 = p.kv_off ? p.kv_off[im.b] : im.b * p.Nk;
    im.skip = kDKV && im.t0 >= im.kv_len;          // this key tile holds no keys of the sample: nothing to compute or write
    im.x_row0 = kDKV ? kv_row0 + im.t0 : im.b * p.Nq + im.t0;
    im.y_row0 = kDKV ? im.b * p.Nq : kv_row0;
    im.n_iter = ((kDKV ? p.Nq : im.kv_len) + kBSub - 1) / kBSub;          // 64-row sub-blocks of the stream
    return im;
  };

  if (threadIdx.x == 0) {
    prefetch_tmap(&tm_x1m); prefetch_tmap(&tm_x1t); prefetch_tmap(&tm_x2m); prefetch_tmap(&tm_x2t);
    prefetch_tmap(&tm_y1m); prefetch_tmap(&tm_y1t); prefetch_tmap(&tm_y2m); prefetch_tmap(&tm_y2t);
    mbar_init(x_full, 1);
    for (int s = 0; s < kBStages; ++s) {
      mbar_init(&y_full[s], 1);
      mbar_init(&y_empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], kBEwThreads);
    }
    mbar_init(acc_full, 1);
    mbar_init(x_empty, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================================================ TMA producer
    if (elect_one()) {
      uint32_t itc = 0, ybase = 0;                 // items with work so far; stream sub-blocks loaded so far (ring position)
      for (int w = blockIdx.x; w < total_items; w += gridDim.x) {
        const Item im = decode(w);
        if (im.skip || im.n_iter == 0) continue;
        const int h = im.h;
        mbar_wait(x_empty, (itc & 1) ^ 1);         // the previous item's score MMAs are done with the stationary tiles
        mbar_arrive_expect_tx(x_full, 2 * kBTile);
        tma_load_3d(smem + kBOffX1, &tm_x1m, x_full, 0, h, im.x_row0, kEvictNormal);
        tma_load_3d(smem + kBOffX1 + kBMain, &tm_x1t, x_full, 64, h, im.x_row0, kEvictNormal);
        tma_load_3d(smem + kBOffX2, &tm_x2m, x_full, 0, h, im.x_row0, kEvictNormal);
        tma_load_3d(smem + kBOffX2 + kBMain, &tm_x2t, x_full, 64, h, im.x_row0, kEvictNormal);
        for (int n = 0; n < im.n_iter; ++n) {
          const uint32_t g = ybase + n;
          const int stage = g % kBStages;
          const uint32_t ph = (g / kBStages) & 1;
          mbar_wait(&y_empty[stage], ph ^ 1);
          uint8_t* y1 = smem + kBOffY + stage * 2 * kBYTile;
          uint8_t* y2 = y1 + kBYTile;
          const int yrow = im.y_row0 + n * kBSub;
          if constexpr (kDKV && kTmaStats) {
            const uint32_t sb = (uint32_t)min(kBSub, p.Nq - n * kBSub) * 4u;            // bytes of lse (and of delta) in this sub-block
            const size_t so = ((size_t)im.b * p.H + h) * p.Nq + (size_t)n * kBSub;
            float* sdst = reinterpret_cast<float*>(smem + kBOffStat) + stage * 128;
            mbar_arrive_expect_tx(&y_full[stage], 2 * kBYTile + 2 * sb);
            tma_load_1d(sdst, p.lse + so, sb, &y_full[stage]);
            tma_load_1d(sdst + 64, p.delta + so, sb, &y_full[stage]);
          } else {
            mbar_arrive_expect_tx(&y_full[stage], 2 * kBYTile);
          }
          tma_load_3d(y1, &tm_y1m, &y_full[stage], 0, h, yrow, kEvictLast);
          tma_load_3d(y1 + kBYMain, &tm_y1t, &y_full[stage], 64, h, yrow, kEvictLast);
          tma_load_3d(y2, &tm_y2m, &y_full[stage], 0, h, yrow, kEvictLast);
          tma_load_3d(y2 + kBYMain, &tm_y2t, &y_full[stage], 64, h, yrow, kEvictLast);
        }
        ybase += im.n_iter;
        ++itc;
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, kBSub, 0, 0);
      constexpr uint32_t idesc_g = make_idesc_bf16(128, 80, 0, 1);     // streamed tile MN-major, N = d 0..79
      const uint32_t sbase = smem_u32(smem);
      const uint32_t t_s = tmem_base + kBColS, t_dp = tmem_base + kBColDP;
      // score product D = X Y^T: both tiles K-major; 4 K-steps in the main parts + 1 in the tails (d 64..79, 72.. zero)
      auto issue_score = [&](uint32_t d, uint32_t x, uint32_t y) {
        const uint64_t xd = make_smem_desc(x, 16, 1024, kLayoutSW128);
        const uint64_t yd = make_smem_desc(y, 16, 1024, kLayoutSW128);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(d, xd + 2 * k, yd + 2 * k, idesc_s, k != 0 ? 1u : 0u);
        const uint64_t xt = make_smem_desc(x + kBMain, 16, 1024, kLayoutSW128);
        const uint64_t yt = make_smem_desc(y + kBYMain, 16, 1024, kLayoutSW128);
        umma_ss(d, xt, yt, idesc_s, 1u);
      };
      // gradient product acc += A Y: A = bf16 P' / dS' in TMEM (K-step k of 16 streamed rows at packed column offset
      // kBCols * (16 k / kBCols) + (16 k % kBCols) / 2: each thread's column slice holds its own packed columns at its
      // start), Y MN-major (main + tail atoms LBO apart)
      auto issue_grad = [&](uint32_t acc, uint32_t a_tmem, uint32_t y, bool first) {
        const uint64_t yd = make_smem_desc(y, kBYMain, 1024, kLayoutSW128);
#pragma unroll
        for (int k = 0; k < kBSub / 16; ++k)
          umma_ts(acc, a_tmem + kBCols * ((16 * k) / kBCols) + ((16 * k) % kBCols) / 2, yd + (uint64_t)(k * (2048 >> 4)), idesc_g,
                  (first && k == 0) ? 0u : 1u);
      };
      // running counts across items: stream sub-blocks so far (ring stage / phase), completed uses of score buffer 0 / 1, items
      uint32_t itc = 0, ybase = 0, hb0 = 0, hb1 = 0;
      for (int w = blockIdx.x; w < total_items; w += gridDim.x) {
        const Item im = decode(w);
        if (im.skip || im.n_iter == 0) continue;
        const int n_iter = im.n_iter;
        // S'(n), dP'(n) into score buffer n & 1
        auto issue_scores = [&](int n) {
          const uint32_t g = ybase + n;
          const int stage = g % kBStages, hh = n & 1;
          mbar_wait(&y_full[stage], (g / kBStages) & 1);
          tc_fence_after();
          const uint32_t y1 = sbase + kBOffY + stage * 2 * kBYTile;
          issue_score(t_s + kBSub * hh, sbase + kBOffX1, y1);
          issue_score(t_dp + kBSub * hh, sbase + kBOffX2, y1 + kBYTile);
          umma_commit(&s_full[hh]);
          if (n + 1 == n_iter) umma_commit(x_empty);   // the item's last read of the stationary tiles
        };
        // The score buffers are free: every P' / dS' of the previous item was consumed by a gradient MMA issued earlier on the
        // (in-order) tensor pipe.  The accumulators are overwritten by the first gradient MMA below, which waits for a P' that the
        // elementwise threads publish only after their epilogue has read the previous item's accumulators.
        mbar_wait(x_full, itc & 1);
        issue_scores(0);
        if (n_iter > 1) issue_scores(1);
        for (int n = 0; n < n_iter; ++n) {
          const int stage = (ybase + n) % kBStages, hh = n & 1;
          const uint32_t y1 = sbase + kBOffY + stage * 2 * kBYTile;
          mbar_wait(&p_full[hh], ((hh ? hb1 : hb0) + (n >> 1)) & 1);
          tc_fence_after();
          if (kDKV) issue_grad(tmem_base + kBColAcc1, t_s + kBSub * hh, y1 + kBYTile, n == 0);    // dV += P' dO
          issue_grad(tmem_base + kBColAcc2, t_dp + kBSub * hh, y1, n == 0);                       // dK += dS' Q  /  dQ += dS' K
          umma_commit(&y_empty[stage]);
          if (n + 2 < n_iter) issue_scores(n + 2);       // into the buffer whose P' / dS' the MMAs above have just consumed
        }
        umma_commit(acc_full);
        ybase += n_iter;
        hb0 += (n_iter + 1) >> 1;
        hb1 += n_iter >> 1;
        ++itc;
      }
    }
  } else if (warp >= 4) {
    // ================================================================ elementwise stage + epilogue
    const int tid = threadIdx.x - 128;             // 0 .. kBEwThreads - 1
    const int half = (warp - 4) >> 2;              // which kBCols-wide slice of a sub-block's 64 score columns
    const int qd = warp & 3;                       // TMEM sub-partition this warp may access
    const int row = qd * 32 + lane;                // stationary row (TMEM lane)
    const uint32_t lane_sel = static_cast<uint32_t>(qd * 32) << 16;
    const float sl2 = p.scale_log2;
    const uint64_t sl2x2 = f32x2(sl2, sl2);
    uint32_t itc = 0, ybase = 0, hb0 = 0, hb1 = 0;   // running counts across items (see the MMA issuer)

    for (int w = blockIdx.x; w < total_items; w += gridDim.x) {
    const Item im = decode(w);
    if (im.skip) continue;
    const int b = im.b, h = im.h, t0 = im.t0, kv_len = im.kv_len, x_row0 = im.x_row0, n_iter = im.n_iter;
    const size_t stat_base = ((size_t)b * p.H + h) * p.Nq;
    uint64_t nlse2 = 0, delta2 = 0;                // dQ pass: (-lse, -lse) and (delta, delta) of this thread's query row
    if (!kDKV) {
      const int qrow = min(t0 + row, p.Nq - 1);
      const float l = -p.lse[stat_base + qrow], d = p.delta[stat_base + qrow];
      nlse2 = f32x2(l, l);
      delta2 = f32x2(d, d);
    }
    else if (!kTmaStats && n_iter > 0) {           // -lse | delta of the first streamed query sub-block -> smem buffer 0
      if (tid < 128) {
        const int qi = min(tid & 63, p.Nq - 1);
        stat[tid] = tid < 64 ? -p.lse[stat_base + qi] : p.delta[stat_base + qi];
      }
      named_bar_sync(1, kBEwThreads);
    }

    for (int n = 0; n < n_iter; ++n) {
      const int hh = n & 1;
      const uint32_t t_s = tmem_base + kBColS + lane_sel + kBSub * hh + kBCols * half;
      const uint32_t t_dp = tmem_base + kBColDP + lane_sel + kBSub * hh + kBCols * half;
      // kTmaStats: +lse | delta of the stage, landed with the tiles; else -lse | delta handed over through smem buffer n & 1
      const float* st = stat + (kTmaStats ? ((ybase + n) % kBStages) : (n & 1)) * 128 + kBCols * half;
      [[maybe_unused]] float nxt = 0.f;
      if constexpr (kTmaStats) {
        if (kDKV) mbar_wait(&y_full[(ybase + n) % kBStages], ((ybase + n) / kBStages) & 1);  // (complete long ago) acquire the TMA-written stats
      } else {
        if (kDKV && n + 1 < n_iter && tid < 128) {   // next sub-block's statistics: global load in flight during this one
          const size_t o = stat_base + min((n + 1) * kBSub + (tid & 63), p.Nq - 1);    // the last sub-block may be partial
          nxt = tid < 64 ? -p.lse[o] : p.delta[o];
        }
      }
      // the last sub-block of the stream (queries in the dKV pass, keys in the dQ pass) may be partial
      const int rem = (kDKV ? p.Nq : kv_len) - n * kBSub - kBCols * half;
      mbar_wait(&s_full[hh], ((hh ? hb1 : hb0) + (n >> 1)) & 1);
      tc_fence_after();
      uint32_t vs[kBCols], vd[kBCols];
      ld_scores(t_s, vs, t_dp, vd);
      uint32_t pp[kBCols / 2], pd[kBCols / 2];
      // packed fp32 pairs (FFMA2 / FADD2 / FMUL2: one issue slot for two elements), exp2 on the MUFU pipe.  The masking
      // selects exist only in the copy of the loop taken by a partial last sub-block of the dQ pass.
      auto compute = [&](auto masked_tag) {
#pragma unroll
        for (int i = 0; i < kBCols; i += 2) {
          uint64_t nl, dl;
          if (kDKV) {
            const float2 l2 = *reinterpret_cast<const float2*>(st + i);          // -lse (+lse with TMA stats) of the two columns
            const float2 d2 = *reinterpret_cast<const float2*>(st + 64 + i);     // delta
            nl = kTmaStats ? f32x2(-l2.x, -l2.y) : f32x2(l2.x, l2.y);
            dl = f32x2(d2.x, d2.y);
          } else {
            nl = nlse2;
            dl = delta2;
          }
          const uint64_t x = fma2(f32x2(__uint_as_float(vs[i]), __uint_as_float(vs[i + 1])), sl2x2, nl);
          float x0, x1;
          f32x2_split(x, x0, x1);
          float e0 = fast_exp2(x0), e1 = fast_exp2(x1);
          if constexpr (decltype(masked_tag)::value) {
            if (i >= rem) e0 = 0.f;
            if (i + 1 >= rem) e1 = 0.f;
          }
          const uint64_t g = mul2(f32x2(e0, e1), sub2(f32x2(__uint_as_float(vd[i]), __uint_as_float(vd[i + 1])), dl));
          float g0, g1;
          f32x2_split(g, g0, g1);
          if constexpr (decltype(masked_tag)::value) {
            // the statistics of columns past the end of the stream are whatever the smem held (the TMA copies stop at the last
            // valid row): 0 * (dP' - NaN) would poison dS' -- select, do not multiply
            if (i >= rem) g0 = 0.f;
            if (i + 1 >= rem) g1 = 0.f;
          }
          pp[i / 2] = pack_bf16x2(e0, e1);
          pd[i / 2] = pack_bf16x2(g0, g1);
        }
      };
      if (rem < kBCols) compute(std::true_type{});
      else compute(std::false_type{});
      // bf16 results over the fp32 columns this thread has just consumed (its own slice: no cross-warp hazard)
      if (kDKV) st_packed(t_s, pp);
      st_packed(t_dp, pd);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[hh]);
      if constexpr (!kTmaStats) {
        if (kDKV) {
          if (n + 1 < n_iter && tid < 128) stat[((n + 1) & 1) * 128 + tid] = nxt;
          named_bar_sync(1, kBEwThreads);
        }
      }
    }

    // ---- epilogue: column slice 0 writes acc2 (dK / dQ, times the softmax scale), slice 1 writes acc1 (dV, dKV pass only)
    if (n_iter > 0) {
      mbar_wait(acc_full, itc & 1);
      tc_fence_after();
    }
    if (half == 0 || (kDKV && half == 1)) {
      const uint32_t t_acc = tmem_base + (half == 0 ? kBColAcc2 : kBColAcc1) + lane_sel;
      const float mul = half == 0 ? p.scale : 1.0f;
      const bool row_ok = kDKV ? (t0 + row < kv_len) : (t0 + row < p.Nq);
      __nv_bfloat16* base = half == 0 ? p.d2 : p.d1;
      const long long sn = half == 0 ? p.d2_sn : p.d1_sn;
      const long long sh = half == 0 ? p.d2_sh : p.d1_sh;
      __nv_bfloat16* dst = base + (size_t)(x_row0 + row) * sn + (size_t)h * sh;
      // the row's 72 gradients (144 B) as 36 packed words; 32 lanes write 32 different rows, so a store instruction costs one
      // LSU sector operation per lane whatever its width: 256-bit stores need 5 per row instead of 9 (attn_sm100.cu)
      uint32_t ow[36];
      if (n_iter > 0) {
        uint32_t oa[32], ob[32], o8[8];
        tmem_ld_32x32b_x32_nowait(t_acc, oa);
        tmem_ld_32x32b_x32_nowait(t_acc + 32, ob);
        tmem_ld_32x32b_x8(t_acc + 64, o8);           // tcgen05.wait::ld covers all three
        tmem_ld_wait_x32(oa);
        tmem_ld_wait_x32(ob);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          ow[i] = pack_bf16x2(__uint_as_float(oa[2 * i]) * mul, __uint_as_float(oa[2 * i + 1]) * mul);
          ow[16 + i] = pack_bf16x2(__uint_as_float(ob[2 * i]) * mul, __uint_as_float(ob[2 * i + 1]) * mul);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) ow[32 + i] = pack_bf16x2(__uint_as_float(o8[2 * i]) * mul, __uint_as_float(o8[2 * i + 1]) * mul);
      } else {
#pragma unroll
        for (int i = 0; i < 36; ++i) ow[i] = 0u;
      }
#if PXA_BWD_BULK_OUT
      // The gradient row leaves as ONE asynchronous TMA bulk copy from smem instead of 5 LSU store instructions that each cost 32
      // transactions per warp (32 lanes = 32 different rows; attn_sm100.cu, profiles/r2_attn_epilogue.txt).  Staging: the stream
      // ring -- acc_full has completed, so every TMA load has landed and every MMA that read a stage has retired.
      {
        uint4* srow = reinterpret_cast<uint4*>(smem + kBOffY + ((warp - 4) * 32 + lane) * 144);
#pragma unroll
        for (int c = 0; c < 9; ++c) srow[c] = make_uint4(ow[4 * c], ow[4 * c + 1], ow[4 * c + 2], ow[4 * c + 3]);
        fence_proxy_async_smem();
        if (row_ok) bulk_store_1d(dst, srow, 144);
        tma_store_commit();
        tma_store_wait_read<0>();                  // the CTA may exit once the copy has read the staging row
      }
#else
      if (row_ok) {
        const int lead = static_cast<int>((reinterpret_cast<uintptr_t>(dst) >> 4) & 1) * 4;   // words before the first 32-byte boundary
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (lead)
            st_global_v8(dst + 8 + 16 * c, ow[4 + 8 * c], ow[5 + 8 * c], ow[6 + 8 * c], ow[7 + 8 * c], ow[8 + 8 * c], ow[9 + 8 * c],
                         ow[10 + 8 * c], ow[11 + 8 * c]);
          else
            st_global_v8(dst + 16 * c, ow[8 * c], ow[8 * c + 1], ow[8 * c + 2], ow[8 * c + 3], ow[8 * c + 4], ow[8 * c + 5],
                         ow[8 * c + 6], ow[8 * c + 7]);
        }
        if (lead) *reinterpret_cast<uint4*>(dst) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        else *reinterpret_cast<uint4*>(dst + 64) = make_uint4(ow[32], ow[33], ow[34], ow[35]);
      }
#endif
    }
    if (n_iter > 0) {
      ybase += n_iter;
      hb0 += (n_iter + 1) >> 1;
      hb1 += n_iter >> 1;
      ++itc;
    }
    }   // items
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// main (d 0..63) and 64-wide tail (d 64..127, zero past 71) maps of a [rows, H, 72] bf16 view, both 128B-swizzled
static int make_bwd_maps(CUtensorMap* main_map, CUtensorMap* tail_map, const void* base, int H, long long rows, long long s_row,
                         long long s_head, int box_rows) {
  uint64_t dims[3] = {72, (uint64_t)H, (uint64_t)rows};
  uint64_t str[2] = {(uint64_t)s_head * 2, (uint64_t)s_row * 2};
  uint32_t box[3] = {64, 1, (uint32_t)box_rows};
  int rc = make_tmap_bf16(main_map, base, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  return make_tmap_bf16(tail_map, base, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

}  // namespace pxa


extern "C" int pxa_flash_attn_d72_bwd_bf16(const PxaAttnBwdArgs* args, void* stream) {
  using namespace pxa;
  if (!args) return fail(PXA_ERR_ARG, "null args");
  const PxaAttnBwdArgs& a = *args;
  if (!a.q || !a.k || !a.v || !a.o || !a.d_o || !a.lse || !a.delta || !a.dq || !a.dk || !a.dv)
    return fail(PXA_ERR_ARG, "null pointer");
  if (a.B <= 0 || a.H <= 0 || a.Nq <= 0 || a.Nk <= 0 || a.kv_rows <= 0) return fail(PXA_ERR_ARG, "bad B/H/Nq/Nk/kv_rows");
  const int64_t strides[] = {a.q_sn, a.q_sh, a.k_sn, a.k_sh, a.v_sn, a.v_sh, a.ldo, a.lddo,
                             a.dq_sn, a.dq_sh, a.dk_sn, a.dk_sh, a.dv_sn, a.dv_sh};
  for (int64_t s : strides)
    if (s & 7) return fail(PXA_ERR_ALIGN, "strides must be multiples of 8 elements");
  if ((reinterpret_cast<uintptr_t>(a.dq) | reinterpret_cast<uintptr_t>(a.dk) | reinterpret_cast<uintptr_t>(a.dv) |
       reinterpret_cast<uintptr_t>(a.o) | reinterpret_cast<uintptr_t>(a.d_o)) & 15)
    return fail(PXA_ERR_ALIGN, "o / dO / dq / dk / dv must be 16-byte aligned");
  const bool tma_stats = PXA_BWD_TMA_STATS && a.Nq % 4 == 0;      // 16-byte aligned bulk copies of lse / delta
  PXA_REQUIRE_SM100();
  int rc = pxa_attn_delta_d72(a.o, a.d_o, a.delta, a.B, a.H, a.Nq, a.ldo, a.lddo, stream);
  if (rc) return rc;
  // *: 128-row boxes (stationary tile), *s: 64-row boxes (streamed sub-blocks)
  CUtensorMap qm, qt, km, kt, vm, vt, gm, gt, qms, qts, kms, kts, vms, vts, gms, gts;
  const long long q_rows = (long long)a.B * a.Nq;
  if ((rc = make_bwd_maps(&qm, &qt, a.q, a.H, q_rows, a.q_sn, a.q_sh, kBT))) return rc;
  if ((rc = make_bwd_maps(&km, &kt, a.k, a.H, a.kv_rows, a.k_sn, a.k_sh, kBT))) return rc;
  if ((rc = make_bwd_maps(&vm, &vt, a.v, a.H, a.kv_rows, a.v_sn, a.v_sh, kBT))) return rc;
  if ((rc = make_bwd_maps(&gm, &gt, a.d_o, a.H, q_rows, a.lddo, 72, kBT))) return rc;
  if ((rc = make_bwd_maps(&qms, &qts, a.q, a.H, q_rows, a.q_sn, a.q_sh, kBSub))) return rc;
  if ((rc = make_bwd_maps(&kms, &kts, a.k, a.H, a.kv_rows, a.k_sn, a.k_sh, kBSub))) return rc;
  if ((rc = make_bwd_maps(&vms, &vts, a.v, a.H, a.kv_rows, a.v_sn, a.v_sh, kBSub))) return rc;
  if ((rc = make_bwd_maps(&gms, &gts, a.d_o, a.H, q_rows, a.lddo, 72, kBSub))) return rc;
  AttnBwdParams p;
  p.lse = a.lse; p.delta = a.delta;
  p.kv_len = a.kv_len; p.kv_off = a.kv_off;
  p.B = a.B; p.H = a.H; p.Nq = a.Nq; p.Nk = a.Nk;
  p.nx = 0;
  p.scale = a.scale;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  {  // dK, dV
    auto kern = tma_stats ? flash_attn_d72_bwd_kernel<true, true> : flash_attn_d72_bwd_kernel<true, false>;
    PXA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem));
    p.d2 = reinterpret_cast<__nv_bfloat16*>(a.dk); p.d2_sn = a.dk_sn; p.d2_sh = a.dk_sh;
    p.d1 = reinterpret_cast<__nv_bfloat16*>(a.dv); p.d1_sn = a.dv_sn; p.d1_sh = a.dv_sh;
    p.nx = (a.Nk + kBT - 1) / kBT;
    const long long items = (long long)p.nx * a.H * a.B;
    const unsigned grid = (unsigned)(PXA_BWD_PERSISTENT && items > device_info().sms ? device_info().sms : items);
    kern<<<grid, kBwdThreads, kBwdSmem, s>>>(km, kt, vm, vt, qms, qts, gms, gts, p);
    launch_counter()++;
    PXA_CHECK_CUDA(cudaGetLastError());
  }
  {  // dQ
    auto kern = flash_attn_d72_bwd_kernel<false, false>;
    PXA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem));
    p.d2 = reinterpret_cast<__nv_bfloat16*>(a.dq); p.d2_sn = a.dq_sn; p.d2_sh = a.dq_sh;
    p.d1 = nullptr; p.d1_sn = p.d1_sh = 0;
    p.nx = (a.Nq + kBT - 1) / kBT;
    const long long items = (long long)p.nx * a.H * a.B;
    const unsigned grid = (unsigned)(PXA_BWD_PERSISTENT && items > device_info().sms ? device_info().sms : items);
    kern<<<grid, kBwdThreads, kBwdSmem, s>>>(qm, qt, gm, gt, kms, kts, vms, vts, p);
    launch_counter()++;
    PXA_CHECK_CUDA(cudaGetLastError());
  }
  return PXA_OK;
}
