// pxa_mlp_fused_bf16: the timm Mlp of PixArtMSBlock (PixArtMS.py:67,77) as ONE persistent kernel
//     x32 += gate[b] * ( gelu_tanh(xn W1^T + b1) W2^T + b2 )
// -- GEMM -> GELU -> GEMM with the hidden activations never leaving the chip's L2 (the north_star's "persistent
// GEMM->GELU->GEMM kernel").  Two launches (pxa_gemm_bf16 with the GELU epilogue, then with the residual epilogue) push the
// [M, 4608] hidden tensor through HBM: 302 MB written + 302 MB read per block at c3.  Round 2 measured that L2 keeps only
// the last ~35 MB of a streaming producer (profiles/r2_l2_chain.md) and that row-grouped launches lose more to wave
// quantisation than they save (profiles/r2_mlp_grouping.txt), so the co-scheduling has to happen INSIDE one launch:
//
//   * the persistent grid of CTA pairs walks ONE tile list that interleaves the two GEMMs in groups of G 256-row panels:
//         fc1(g0) | fc1(g1) fc2(g0) | fc1(g2) fc2(g1) | ... | fc2(g_last)
//     (fc1 tiles 256 x 256 over K = 1152, fc2 tiles 256 x 192 over K = 4608; n fastest, so the clusters working side by
//     side share the panel's rows through L2).  fc2 runs ONE GROUP BEHIND fc1: when a cluster reaches a fc2 tile, the fc1
//     tiles it depends on were handed out a whole segment earlier and are complete in practice -- the wait below is a formality;
//   * the hidden tiles go to a RING of R groups (R x G x 256 rows x 4608 bf16 = 28 MB for G = 4, R = 3) that is reused as the
//     panels advance, so the lines are overwritten in L2 before they are ever evicted: neither the reads nor (mostly) the
//     writes reach HBM;
//   * dependencies are two small counter arrays in global memory: `flags[panel][rank]` counts the finished fc1 tiles of a
//     panel per CTA rank (a CTA only ever reads back hidden rows that CTAs of its own rank wrote), `done2[group]` the fc2
//     tiles of a group whose mainloop has consumed the hidden rows (ring slot reuse).  Producers publish with
//     bar.sync + __threadfence + atomicAdd, consumers spin with ld.acquire and issue fence.proxy.async before the TMA loads.
//     All dependencies point backwards in the tile list and every cluster walks it in order on a resident grid: no deadlock.
//
// Roles per CTA (384 threads, the pair kernel's layout): warp 0 TMA producer (both CTAs), warp 1 MMA issuer (leader; UMMA
// 256 x BN x 16, BN = 256 / 192 per tile kind), warp 2 TMEM allocator (2 x 256 fp32 columns: tile i's epilogue overlaps tile
// i+1's mainloop), warps 4-11 epilogue: fc1 tiles -- two warps per TMEM lane quarter take alternate 32-column chunks, bias +
// GELU(tanh) -> bf16 -> smem transpose -> coalesced stores into the ring; fc2 tiles -- warps 4-7 compute gate * (acc + b2) per
// 32-column chunk into swizzled smem and one elected thread adds it into the fp32 residual stream with TMA reduce-add.
#include "gemm_common.cuh"
#include "pair_common.cuh"

namespace pxa {

constexpr int kMlpThreads = 128 + 32 * 8;
constexpr int kMlpBN1 = 256, kMlpBN2 = 192;
constexpr int kMlpStageA = kBM * kBK * 2;                       // 16 KB: this CTA's 128 rows of A
constexpr int kMlpStageB = (kMlpBN1 / 2) * kBK * 2;             // 16 KB slot (fc2 uses 12 KB of it)
constexpr int kMlpStage = kMlpStageA + kMlpStageB;
constexpr int kMlpRing = 2;                                     // chunk buffers of the reduce-add epilogue
constexpr int kMlpEpiBufs = 8 * 2048 + kMlpRing * kResChunkBytes;   // 8 transpose tiles (bf16) + 2 fp32 chunk buffers
constexpr int kMlpEpiSmem = kMlpEpiBufs + ((kEpiConstBytes + 127) / 128) * 128;
constexpr int kMlpStages = (227 * 1024 - kMlpEpiSmem - 256 - 1024) / kMlpStage;
constexpr int kMlpSmem = kMlpStages * kMlpStage + kMlpEpiSmem + 256 + 1024;
static_assert(kMlpStages >= 4, "smem ring too shallow");

struct MlpParams {
  GemmParams p1;     // fc1: bias = b1, out = hidden ring (bf16, ldo = N1), M = ring rows, N = N1, K = K1
  GemmParams p2;     // fc2: bias = b2, gate, out = residual = x32 (fp32, ldo = N2), M, N = N2, K = N1
  GemmParams p2nb;   // fc2 without the bias: K-splits 1.. of an output tile add only gate * acc
  int panels, group, ring, ngroups, lag;
  int ksplit;        // fc2 tiles cover N1 / ksplit of the reduction each (the reduce-add epilogue sums them)
  int t1, t2;        // tiles per panel (t2 = N2 / 192 * ksplit, K-split fastest)
  int* flags;        // [panels][2]
  int* done2;        // [ngroups]
};

struct MlpTile {
  int kind;          // 0 = fc1, 1 = fc2
  int panel, n, grp;
};

PXA_DEVICE int mlp_group_panels(const MlpParams& q, int g) { return min(q.group, q.panels - g * q.group); }

// tile list position -> tile.  Segment s (0 <= s < ngroups + lag) holds fc1(group s) if s < ngroups, then fc2(group s - lag)
// if s >= lag: for lag = 1   fc1(g0) | fc1(g1) fc2(g0) | ... | fc2(g_last).  Every group but the last is full, so the head
// (fc1 only) and the regular middle segments are found by division and only the <= lag + 1 segments around the ragged last
// group are walked.
PXA_DEVICE MlpTile mlp_decode(const MlpParams& q, int t) {
  MlpTile r;
  const int G1 = q.group * q.t1, G2 = q.group * q.t2;
  const int nh = min(q.lag, q.ngroups - 1);          // full fc1-only segments
  if (t < nh * G1) {
    r.kind = 0; r.grp = t / G1; r.panel = t / q.t1; r.n = t % q.t1;
    return r;
  }
  t -= nh * G1;
  int s = nh;
  if (q.lag < q.ngroups - 1) {
    const int nm = q.ngroups - 1 - q.lag;            // regular segments fc1(s) fc2(s - lag), both groups full
    if (t < nm * (G1 + G2)) {
      const int k = t / (G1 + G2);
      s += k;
      t -= k * (G1 + G2);
      if (t < G1) { r.kind = 0; r.grp = s; r.panel = s * q.group + t / q.t1; r.n = t % q.t1; }
      else { t -= G1; r.kind = 1; r.grp = s - q.lag; r.panel = r.grp * q.group + t / q.t2; r.n = t % q.t2; }
      return r;
    }
    t -= nm * (G1 + G2);
    s += nm;
  }
  for (;; ++s) {
    const int a = s < q.ngroups ? mlp_group_panels(q, s) * q.t1 : 0;
    const int b = s >= q.lag ? mlp_group_panels(q, s - q.lag) * q.t2 : 0;
    if (t < a) { r.kind = 0; r.grp = s; r.panel = s * q.group + t / q.t1; r.n = t % q.t1; return r; }
    t -= a;
    if (t < b || s == q.ngroups + q.lag - 1) { r.kind = 1; r.grp = s - q.lag; r.panel = r.grp * q.group + t / q.t2; r.n = t % q.t2; return r; }
    t -= b;
  }
}
// first row of panel `panel` inside the hidden ring
PXA_DEVICE int mlp_ring_row(const MlpParams& q, int panel) {
  return ((panel / q.group) % q.ring) * q.group * (2 * kBM) + (panel % q.group) * (2 * kBM);
}

PXA_DEVICE int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
PXA_DEVICE void spin_until(const int* p, int target) {
  while (ld_acquire_gpu(p) < target) __nanosleep(40);
}
PXA_DEVICE void fence_proxy_async_all() { asm volatile("fence.proxy.async;\n" ::: "memory"); }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kMlpThreads, 1)
mlp_fused_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w1,
                 const __grid_constant__ CUtensorMap tm_h, const __grid_constant__ CUtensorMap tm_w2,
                 const __grid_constant__ CUtensorMap tm_out, const MlpParams q) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_smem = smem + kMlpStages * kMlpStage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + kMlpEpiSmem);
  uint64_t* full_bar = bars;                          // [kMlpStages] the leader's is live: both CTAs' TMA bytes land there
  uint64_t* empty_bar = bars + kMlpStages;            // [kMlpStages] per CTA (multicast commit)
  uint64_t* tfull_bar = bars + 2 * kMlpStages;        // [2] per CTA: accumulator ready
  uint64_t* tempty_bar = bars + 2 * kMlpStages + 2;   // [2] the leader's: 2 x 256 arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMlpStages + 4);

  const int warp = warp_idx_sync();
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tm_x); prefetch_tmap(&tm_w1); prefetch_tmap(&tm_h); prefetch_tmap(&tm_w2); prefetch_tmap(&tm_out);
    for (int s = 0; s < kMlpStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 2 * 256);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc_pair<512>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_tiles = q.panels * (q.t1 + q.t2);
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int nkb1 = (q.p1.K + kBK - 1) / kBK, nkb2 = q.p2.K / kBK / q.ksplit;      // k-blocks per tile

  if (warp == 0) {
    // ================================================================ TMA producer (both CTAs)
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const MlpTile tl = mlp_decode(q, tile);
        const int m0 = tl.panel * (2 * kBM) + rank * kBM;              // logical row of this CTA's half of the panel
        const int hrow = mlp_ring_row(q, tl.panel) + rank * kBM;        // the same rows inside the hidden ring
        if (tl.kind == 0) {
          // ring slot reuse: the fc2 tiles of the group that used this slot R groups ago have consumed their hidden rows
          if (tl.grp >= q.ring) spin_until(q.done2 + (tl.grp - q.ring), mlp_group_panels(q, tl.grp - q.ring) * q.t2 * 2);
          for (int kb = 0; kb < nkb1; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * kMlpStage;
            if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * (kMlpStageA + (kMlpBN1 / 2) * kBK * 2));
            tma_load_2d_pair(sa, &tm_x, &full_bar[stage], kb * kBK, m0, kEvictNormal);
            tma_load_2d_pair(sa + kMlpStageA, &tm_w1, &full_bar[stage], kb * kBK, tl.n * kMlpBN1 + rank * (kMlpBN1 / 2), kEvictLast);
            if (++stage == kMlpStages) { stage = 0; phase ^= 1; }
          }
        } else {
          // all fc1 tiles of this panel written by CTAs of this rank have landed (generic-proxy stores of other SMs), then
          // order the async-proxy (TMA) reads behind the acquire
          spin_until(q.flags + tl.panel * 2 + rank, q.t1);
          fence_proxy_async_all();
          const int n2 = tl.n / q.ksplit, kb0 = (tl.n % q.ksplit) * nkb2;
          for (int kb = kb0; kb < kb0 + nkb2; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * kMlpStage;
            if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * (kMlpStageA + (kMlpBN2 / 2) * kBK * 2));
            tma_load_2d_pair(sa, &tm_h, &full_bar[stage], kb * kBK, hrow, kEvictNormal);
            tma_load_2d_pair(sa + kMlpStageA, &tm_w2, &full_bar[stage], kb * kBK, n2 * kMlpBN2 + rank * (kMlpBN2 / 2), kEvictLast);
            if (++stage == kMlpStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer (leader CTA only)
    if (leader && elect_one()) {
      constexpr uint32_t idesc1 = make_idesc_bf16(2 * kBM, kMlpBN1, 0, 0);
      constexpr uint32_t idesc2 = make_idesc_bf16(2 * kBM, kMlpBN2, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const MlpTile tl = mlp_decode(q, tile);
        const uint32_t idesc = tl.kind == 0 ? idesc1 : idesc2;
        const int nkb = tl.kind == 0 ? nkb1 : nkb2;
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * 256;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kMlpStage);
          const uint64_t adesc = make_smem_desc(sa, 16, 1024, kLayoutSW128);
          const uint64_t bdesc = make_smem_desc(sa + kMlpStageA, 16, 1024, kLayoutSW128);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) umma2_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma2_commit_mc(&empty_bar[stage], 0x3);
          if (++stage == kMlpStages) { stage = 0; phase ^= 1; }
        }
        umma2_commit_mc(&tfull_bar[as], 0x3);
        as ^= 1;
        if (as == 0) aphase ^= 1;
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ================================================================ epilogue (both CTAs, own 128 rows)
    const int e = warp - kEpiWarp0;                  // 0..7
    const int qd = warp & 3;                         // TMEM lane quarter
    const int half = e >> 2;                         // fc1: chunks half, half + 2, ...; fc2: half 0 works, half 1 only signs off
    const int tid = threadIdx.x - kEpiWarp0 * 32;    // 0..255
    const int r = qd * 32 + lane;                    // row of the 128-row tile (fc2 path, half 0)
    uint8_t* stile = epi_smem + e * 2048;
    uint8_t* rbufs = epi_smem + 8 * 2048;
    EpiConst* consts = reinterpret_cast<EpiConst*>(epi_smem + kMlpEpiBufs);
    const bool issuer_warp = warp == kEpiWarp0;
    int as = 0, titer = 0, g = 0;
    uint32_t aphase = 0;
    int pend_panel = -1, pend_grp = -1;              // completion signals of the previous tile, published after this tile's barrier
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++titer) {
      const MlpTile tl = mlp_decode(q, tile);
      const int m0 = tl.panel * (2 * kBM) + rank * kBM;
      const int hrow = mlp_ring_row(q, tl.panel) + rank * kBM;
      const GemmParams& p = tl.kind == 0 ? q.p1 : (tl.n % q.ksplit == 0 ? q.p2 : q.p2nb);
      const int n0 = tl.kind == 0 ? tl.n * kMlpBN1 : (tl.n / q.ksplit) * kMlpBN2;
      const int nch = (tl.kind == 0 ? kMlpBN1 : kMlpBN2) / 32;
      EpiConst* cb = consts + (titer & 1);
      if (tid < kNumEpiThreads) {
        EpiRegs<256> er;
        load_epi_consts<256>(er, p, tid, m0, n0);
        store_epi_consts<256>(cb, er, tid);
      }
      named_bar_sync(2, 256);                        // consts visible; every warp has finished the previous tile
      if (tid == 0) {                                // ... so its hidden stores / its reads are complete: publish
        if (pend_panel >= 0) { __threadfence(); atomicAdd(q.flags + pend_panel * 2 + rank, 1); }
        if (pend_grp >= 0) atomicAdd(q.done2 + pend_grp, 1);
      }
      pend_panel = pend_grp = -1;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      if (tl.kind == 1) pend_grp = tl.grp;           // accumulator complete => the mainloop has read all its hidden rows
      const uint32_t t_acc = tmem_base + as * 256 + (static_cast<uint32_t>(qd * 32) << 16);
      auto release_acc = [&]() {
        tc_fence_before();
        mbar_arrive_cluster(&tempty_bar[as], 0);
      };
      if (tl.kind == 0) {
        // ---- fc1: bias + GELU -> bf16 -> hidden ring, 8 warps on alternate chunks
        uint32_t va[32], vb[32];
        tmem_ld_32x32b_x32_nowait(t_acc + half * 32, va);
#pragma unroll 1
        for (int cc = half; cc < nch; cc += 4) {
          tmem_ld_wait_x32(va);
          if (cc + 2 < nch) tmem_ld_32x32b_x32_nowait(t_acc + (cc + 2) * 32, vb); else release_acc();
          epilogue_chunk_bf16_c<PXA_EPI_BIAS_GELU>(va, p, cb, stile, lane, hrow + qd * 32, n0 + cc * 32, cc * 32);
          if (cc + 2 < nch) {
            tmem_ld_wait_x32(vb);
            if (cc + 4 < nch) tmem_ld_32x32b_x32_nowait(t_acc + (cc + 4) * 32, va); else release_acc();
            epilogue_chunk_bf16_c<PXA_EPI_BIAS_GELU>(vb, p, cb, stile, lane, hrow + qd * 32, n0 + (cc + 2) * 32, (cc + 2) * 32);
          }
        }
        pend_panel = tl.panel;
      } else if (half == 1) {
        release_acc();                               // these four warps do not read fc2 accumulators
      } else {
        // ---- fc2: gate * (acc + b2) per 32-column chunk -> swizzled smem -> TMA reduce-add into the fp32 stream
        float2 st = make_float2(0.f, 0.f);
        uint32_t va[32], vb[32];
        auto process = [&](uint32_t (&v)[32], int cc) {
          uint8_t* rb = rbufs + (g % kMlpRing) * kResChunkBytes;
          residual_chunk_row_c(v, p, cb, rb, nullptr, r, cc * 32, m0 + r, n0 + cc * 32, true, st);
          fence_proxy_async_smem();
          if (issuer_warp) {
            if (elect_one()) tma_store_wait_read<0>();
          }
          named_bar_sync(1, kNumEpiThreads);
          if (issuer_warp) {
            if (elect_one()) {
              tma_reduce_add_2d(&tm_out, rb, n0 + cc * 32, m0);
              tma_store_commit();
            }
          }
          ++g;
        };
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
      }
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
    named_bar_sync(2, 256);                          // the last tile's stores are complete
    if (tid == 0) {
      if (pend_panel >= 0) { __threadfence(); atomicAdd(q.flags + pend_panel * 2 + rank, 1); }
      if (pend_grp >= 0) atomicAdd(q.done2 + pend_grp, 1);
    }
    if (issuer_warp) {
      if (elect_one()) tma_store_wait_all<0>();
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair<512>(tmem_base);
  }
}

}  // namespace pxa

extern "C" int pxa_mlp_fused_bf16(const PxaMlpArgs* args, void* stream) {
  using namespace pxa;
  if (!args) return fail(PXA_ERR_ARG, "null args");
  const PxaMlpArgs& a = *args;
  if (!a.x || !a.w1 || !a.w2 || !a.x32 || !a.hidden_ws || !a.flags_ws) return fail(PXA_ERR_ARG, "null pointer");
  if (a.M <= 0 || a.K1 <= 0 || a.N1 <= 0 || a.N2 <= 0) return fail(PXA_ERR_ARG, "bad shape");
  if (a.N1 % kMlpBN1 || a.N2 % kMlpBN2 || a.K1 % 8 || a.N1 % 8)
    return fail(PXA_ERR_ARG, "pxa_mlp_fused_bf16 needs N1 %% 256 == 0 and N2 %% 192 == 0 (got N1=%d N2=%d)", a.N1, a.N2);
  if ((a.ldx & 7) || (a.ldw1 & 7) || (a.ldw2 & 7) || (a.ldo & 3)) return fail(PXA_ERR_ALIGN, "row strides must keep rows 16-byte aligned");
  if ((reinterpret_cast<uintptr_t>(a.x32) | reinterpret_cast<uintptr_t>(a.b1) | reinterpret_cast<uintptr_t>(a.b2) |
       reinterpret_cast<uintptr_t>(a.gate) | reinterpret_cast<uintptr_t>(a.hidden_ws) | reinterpret_cast<uintptr_t>(a.flags_ws)) & 15)
    return fail(PXA_ERR_ALIGN, "pointers must be 16-byte aligned");
  if (a.gate && (a.gate_batch_stride & 3)) return fail(PXA_ERR_ALIGN, "gate_batch_stride must be a multiple of 4");
  const int rpb = a.rows_per_batch > 0 ? a.rows_per_batch : a.M;
  if (a.gate && rpb < kBM && a.M > rpb) return fail(PXA_ERR_ARG, "rows_per_batch must be >= 128 (a tile spans two samples at most)");
  PXA_REQUIRE_SM100();
  MlpParams q;
  q.panels = (a.M + 2 * kBM - 1) / (2 * kBM);
  q.group = a.group > 0 ? a.group : 4;
  q.ngroups = (q.panels + q.group - 1) / q.group;
  q.lag = a.lag > 0 ? a.lag : 1;
  if (q.lag > q.ngroups) q.lag = q.ngroups;          // all of fc1, then all of fc2
  q.ring = a.ring > 0 ? a.ring : q.lag + 2;
  // fc1(g) reuses the slot of group g - ring, whose fc2 tiles sit in segment g - ring + lag: that must be an EARLIER segment
  if (q.ring < q.lag + 1) return fail(PXA_ERR_ARG, "ring (%d) must be > lag (%d)", q.ring, q.lag);
  q.t1 = a.N1 / kMlpBN1;
  // K-split of fc2 so that its 256 x 192 tiles cost what a 256 x 256 fc1 tile costs (N1 / ksplit * 192 ~ K1 * 256): a static
  // round-robin over a list that mixes tiles of 1x and 3x duration leaves the clusters badly balanced
  q.ksplit = a.k_splits > 0 ? a.k_splits : (int)((2LL * a.N1 * kMlpBN2 + (long long)a.K1 * kMlpBN1) / (2LL * a.K1 * kMlpBN1));
  if (q.ksplit < 1) q.ksplit = 1;
  while (q.ksplit > 1 && (a.N1 / kBK) % q.ksplit) --q.ksplit;
  if (a.N1 % kBK) return fail(PXA_ERR_ARG, "N1 must be a multiple of 64");
  q.t2 = a.N2 / kMlpBN2 * q.ksplit;
  const long long ring_rows = (long long)q.ring * q.group * 2 * kBM;
  if (a.hidden_ws_bytes < ring_rows * a.N1 * 2) return fail(PXA_ERR_ARG, "hidden_ws too small: %lld bytes needed", ring_rows * a.N1 * 2);
  const long long flag_ints = (long long)q.panels * 2 + q.ngroups;
  if (a.flags_ws_bytes < flag_ints * 4) return fail(PXA_ERR_ARG, "flags_ws too small: %lld bytes needed", flag_ints * 4);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  PXA_CHECK_CUDA(cudaMemsetAsync(a.flags_ws, 0, flag_ints * 4, s));
  q.flags = reinterpret_cast<int*>(a.flags_ws);
  q.done2 = q.flags + q.panels * 2;

  CUtensorMap tx, tw1, th, tw2, to;
  int rc;
  {
    uint64_t dims[2] = {(uint64_t)a.K1, (uint64_t)a.M};
    uint64_t str[1] = {(uint64_t)a.ldx * 2};
    uint32_t box[2] = {kBK, kBM};
    if ((rc = make_tmap_bf16(&tx, a.x, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)a.K1, (uint64_t)a.N1};
    uint64_t str[1] = {(uint64_t)a.ldw1 * 2};
    uint32_t box[2] = {kBK, kMlpBN1 / 2};
    if ((rc = make_tmap_bf16(&tw1, a.w1, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)a.N1, (uint64_t)ring_rows};
    uint64_t str[1] = {(uint64_t)a.N1 * 2};
    uint32_t box[2] = {kBK, kBM};
    if ((rc = make_tmap_bf16(&th, a.hidden_ws, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)a.N1, (uint64_t)a.N2};
    uint64_t str[1] = {(uint64_t)a.ldw2 * 2};
    uint32_t box[2] = {kBK, kMlpBN2 / 2};
    if ((rc = make_tmap_bf16(&tw2, a.w2, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)a.N2, (uint64_t)a.M};
    uint64_t str[1] = {(uint64_t)a.ldo * 4};
    uint32_t box[2] = {32, kBM};
    if ((rc = make_tmap(&to, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, a.x32, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  }
  auto fill = [](GemmParams& p) {
    p.bias = nullptr; p.out = nullptr; p.out_aux = nullptr; p.residual = nullptr; p.gate = nullptr; p.gate_batch_stride = 0;
    p.rows_per_batch = 1; p.M = p.N = p.K = p.ldo = 0; p.num_m_tiles = p.num_n_tiles = 0; p.trace = nullptr; p.k_splits = 1;
    p.aux_branch = 0; p.reverse_tiles = 0; p.aux_scale = nullptr; p.aux_scale_batch_stride = 0; p.row_stats_out = nullptr;
    p.ln_stats = nullptr; p.ln_u = p.ln_v = nullptr; p.ln_uv_batch_stride = 0; p.ln_inv_dim = 0.f; p.ln_eps = 0.f;
    p.conv_H = p.conv_W = p.conv_tile_w = p.conv_tile_h = p.conv_cin_blocks = 0;
  };
  fill(q.p1);
  q.p1.bias = reinterpret_cast<const __nv_bfloat16*>(a.b1);
  q.p1.out = a.hidden_ws;
  q.p1.M = (int)ring_rows; q.p1.N = a.N1; q.p1.K = a.K1; q.p1.ldo = a.N1;
  q.p1.rows_per_batch = 1 << 30;
  fill(q.p2);
  q.p2.bias = reinterpret_cast<const __nv_bfloat16*>(a.b2);
  q.p2.out = a.x32; q.p2.residual = a.x32;
  q.p2.gate = a.gate; q.p2.gate_batch_stride = a.gate_batch_stride;
  q.p2.rows_per_batch = rpb;
  q.p2.M = a.M; q.p2.N = a.N2; q.p2.K = a.N1; q.p2.ldo = a.ldo;
  q.p2nb = q.p2;
  q.p2nb.bias = nullptr;
  PXA_CHECK_CUDA(cudaFuncSetAttribute(mlp_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMlpSmem));
  int clusters = device_info().sms / 2;
  if (a.max_ctas > 0 && a.max_ctas / 2 < clusters) clusters = a.max_ctas / 2 > 0 ? a.max_ctas / 2 : 1;
  const int tiles = q.panels * (q.t1 + q.t2);
  if (tiles < clusters) clusters = tiles;
  mlp_fused_kernel<<<2 * clusters, kMlpThreads, kMlpSmem, s>>>(tx, tw1, th, tw2, to, q);
  launch_counter()++;
  PXA_CHECK_CUDA(cudaGetLastError());
  return PXA_OK;
}
