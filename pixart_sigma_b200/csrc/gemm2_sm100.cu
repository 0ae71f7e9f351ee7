// 2-CTA (cta_group::2) variant of pxa_gemm_bf16: a CTA pair on one TPC computes a 256 x BN tile.
//
// Each CTA stages its own 128 rows of A and its own BN/2 rows of W per K-block, so per-SM operand traffic from L2 and
// through shared memory drops from (128 + BN) to (128 + BN/2) rows per 128xBNx64 of MMA work (-30 % at BN = 192); the
// leader CTA's single elected thread issues tcgen05.mma.cta_group::2 (UMMA 256 x BN x 16) which reads both CTAs'
// shared memory and writes each CTA's half of the accumulator into that CTA's own TMEM.
//
//   rank 0 (leader)                                   rank 1
//   warp 0  TMA: A[m0..+128], W[n0..+BN/2]            TMA: A[m0+128..+128], W[n0+BN/2..+BN/2]
//           both signal complete_tx on the LEADER's full barrier (peer bit of the mbarrier address cleared)
//   warp 1  MMA issue + tcgen05.commit multicast      (idle)
//           -> empty[stage] and tmem_full[] of BOTH CTAs
//   warp 2  tcgen05.alloc.cta_group::2 (both CTAs, same warp id)
//   warps 4-7 epilogue of the CTA's own 128 rows; tmem_empty is the leader's barrier (256 arrivals, peer arrives
//           through mapa / shared::cluster)
#include "gemm_common.cuh"
#include "pair_common.cuh"

namespace pxa {

// kTmaRes: the fp32 residual epilogue streams the residual tile through smem by TMA (gemm_common.cuh tma_res_epilogue);
// its 64 KB of chunk buffers cost two operand stages at BN = 192 (5 instead of 7).
template <int BN, bool kTmaRes = false> struct Gemm2Cfg {
  static constexpr int kStageA = kBM * kBK * 2;                 // 16 KB: this CTA's 128 rows of A
  static constexpr int kStageB = (BN / 2) * kBK * 2;            // this CTA's BN/2 rows of W
  static constexpr int kStage = kStageA + kStageB;
  static constexpr int kEpiBufs = kTmaRes ? kResEpiSmem : 4 * 32 * 32 * 4;
  static constexpr int kEpiSmem = kEpiBufs + ((kEpiConstBytes + 127) / 128) * 128;
  static constexpr int kBarBytes = 256;
  static constexpr int kStages = (227 * 1024 - kEpiSmem - kBarBytes - 1024) / kStage > 8
                                     ? 8 : (227 * 1024 - kEpiSmem - kBarBytes - 1024) / kStage;
  static constexpr int kSmem = kStages * kStage + kEpiSmem + kBarBytes + 1024;
  static constexpr uint32_t kTmemCols = (2 * BN <= 256) ? 256 : 512;
};

// kEpiWarps = 8 (bf16 epilogues): TWO epilogue warps per TMEM lane quarter take alternate 32-column chunks of a tile.
// Measured in round 2 (tools/gemm_bench2.py): with 4 epilogue warps the bias / GELU epilogue of a 256 x 256 tile takes about
// as long as the tile's mainloop under the power cap, so every instruction added to it (e.g. the fused LayerNorm algebra)
// showed up 1:1 in the kernel time; halving it makes these GEMMs mainloop-bound again.
template <int BN, int EPI, typename OutT, bool kTmaRes = false, int kEpiWarps = 4>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128 + 32 * kEpiWarps, 1)
gemm2_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                  const __grid_constant__ CUtensorMap tmap_res, const __grid_constant__ CUtensorMap tmap_out,
                  const __grid_constant__ CUtensorMap tmap_aux, const GemmParams p) {
  static_assert(!kTmaRes || (EPI == PXA_EPI_BIAS_RESIDUAL && sizeof(OutT) == 4), "TMA residual epilogue: fp32 stream only");
  using Cfg = Gemm2Cfg<BN, kTmaRes>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_smem = smem + kStages * Cfg::kStage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + Cfg::kEpiSmem);
  uint64_t* full_bar = bars;                      // [kStages] leader's is the live one: both CTAs' TMA bytes land here
  uint64_t* empty_bar = bars + kStages;           // [kStages] per CTA: leader's MMA commit (multicast) frees the slot
  uint64_t* tfull_bar = bars + 2 * kStages;       // [2] per CTA: accumulator ready (multicast commit)
  uint64_t* tempty_bar = bars + 2 * kStages + 2;  // [2] leader's: 256 arrivals (both CTAs' epilogue threads)
  uint64_t* res_full = bars + 2 * kStages + 4;    // [kResBufs] per CTA: TMA residual chunk landed (kTmaRes only)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4 + kResBufs);

  const int warp = warp_idx_sync();
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_w);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 2 * 32 * kEpiWarps);
    }
    for (int s = 0; s < kResBufs; ++s) mbar_init(&res_full[s], 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc_pair<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();               // barriers of BOTH CTAs are initialised before anybody signals remotely
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_tiles = p.num_m_tiles * p.num_n_tiles;      // tiles of 256 x BN, one per cluster at a time
  const int num_kb = (p.K + kBK - 1) / kBK;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  // tile id -> tile: front to back, or back to front (p.reverse_tiles: start on the rows of A that are still in L2)
  auto eff = [&](int t) { return p.reverse_tiles ? num_tiles - 1 - t : t; };

  if (warp == 0) {
    // ================================================================ TMA producer (both CTAs)
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int m0 = (eff(tile) / p.num_n_tiles) * (2 * kBM) + rank * kBM;
        const int n0 = (eff(tile) % p.num_n_tiles) * BN;
        if constexpr (kTmaRes) {
          // pull this tile's residual block into L2 now: the epilogue reads it one mainloop later
          if (!(p.out_aux == nullptr && p.residual == p.out && p.row_stats_out == nullptr))
            for (int c = 0; c < BN / 32; ++c) tma_prefetch_l2_2d(&tmap_res, n0 + c * 32, m0);
        } else if constexpr (EPI == PXA_EPI_BIAS_RESIDUAL) {
          tma_prefetch_l2_2d(&tmap_res, n0, m0);
        }
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::kStage;
          uint8_t* sb = sa + Cfg::kStageA;
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::kStage);   // bytes of both CTAs
          tma_load_2d_pair(sa, &tmap_a, &full_bar[stage], kb * kBK, m0, kEvictNormal);
          tma_load_2d_pair(sb, &tmap_w, &full_bar[stage], kb * kBK, n0 + rank * (BN / 2), kEvictLast);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer (leader CTA only)
    if (leader && elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * kBM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStage);
          const uint64_t adesc = make_smem_desc(sa, 16, 1024, kLayoutSW128);
          const uint64_t bdesc = make_smem_desc(sa + Cfg::kStageA, 16, 1024, kLayoutSW128);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)
            umma2_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma2_commit_mc(&empty_bar[stage], 0x3);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma2_commit_mc(&tfull_bar[as], 0x3);
        as ^= 1;
        if (as == 0) aphase ^= 1;
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ================================================================ epilogue (both CTAs, own 128 rows)
    EpiConst* consts = reinterpret_cast<EpiConst*>(epi_smem + Cfg::kEpiBufs);
    if constexpr (kTmaRes) {
      TileWalk tw;
      tw.first = cluster_id; tw.stride = num_clusters; tw.count = num_tiles;
      tw.mn_tiles = num_tiles; tw.num_n_tiles = p.num_n_tiles; tw.m_mult = 2 * kBM; tw.m_off = rank * kBM;
      tw.reverse = p.reverse_tiles;
      tma_res_epilogue<BN, kResBufs, true>(p, tw, epi_smem, consts, tfull_bar, tempty_bar, res_full, tmem_base, &tmap_res,
                                           &tmap_out, &tmap_aux, warp, lane);
    } else {
      constexpr bool kLn = (EPI == PXA_EPI_LN_BIAS || EPI == PXA_EPI_LN_BIAS_GELU);
      static_assert(kEpiWarps == 4 || (kEpiWarps == 8 && EPI != PXA_EPI_BIAS_RESIDUAL), "8 epilogue warps: bf16 epilogues only");
      constexpr int kStep = kEpiWarps / 4;                         // this warp takes chunks half, half + kStep, ...
      const int q = warp & 3;                                      // TMEM lane quarter (warp % 4)
      const int half = (warp - kEpiWarp0) >> 2;                    // 0, or 0 / 1 with 8 epilogue warps
      const int tid = threadIdx.x - kEpiWarp0 * 32;                // 0 .. 32 * kEpiWarps - 1; the first 128 stage the constants
      uint8_t* stile = epi_smem + (warp - kEpiWarp0) * (4096 / kStep);   // per-warp transpose tile (bf16: 2 KB is enough)
      int as = 0;
      uint32_t aphase = 0;
      int titer = 0;
      EpiRegs<BN> er;                                              // constants / LN statistics of a tile, loaded one tile ahead
      [[maybe_unused]] LnStatRegs sr;
      if (cluster_id < num_tiles) {
        if (tid < kNumEpiThreads)
          load_epi_consts<BN, kLn>(er, p, tid, (eff(cluster_id) / p.num_n_tiles) * (2 * kBM) + rank * kBM, (eff(cluster_id) % p.num_n_tiles) * BN);
        if constexpr (kLn) load_ln_stats(sr, p, (eff(cluster_id) / p.num_n_tiles) * (2 * kBM) + rank * kBM + q * 32 + lane);
      }
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++titer) {
        const int m0 = (eff(tile) / p.num_n_tiles) * (2 * kBM) + rank * kBM;
        const int n0 = (eff(tile) % p.num_n_tiles) * BN;
        const int nch = chunks_of_tile<BN>(p, n0);
        EpiConst* cb = consts + (titer & 1);
        if (tid < kNumEpiThreads) store_epi_consts<BN>(cb, er, tid);
        [[maybe_unused]] float2 ln = make_float2(1.f, 0.f);
        if constexpr (kLn) ln = ln_row_coeffs(p, sr);
        named_bar_sync(2, 32 * kEpiWarps);
        if (tile + num_clusters < num_tiles) {                     // next tile's loads: in flight during this tile's chunks
          const int nt_ = eff(tile + num_clusters);
          if (tid < kNumEpiThreads)
            load_epi_consts<BN, kLn>(er, p, tid, (nt_ / p.num_n_tiles) * (2 * kBM) + rank * kBM, (nt_ % p.num_n_tiles) * BN);
          if constexpr (kLn) load_ln_stats(sr, p, (nt_ / p.num_n_tiles) * (2 * kBM) + rank * kBM + q * 32 + lane);
        }
        [[maybe_unused]] const bool second = q * 32 + lane >= cb->row_split;
        mbar_wait(&tfull_bar[as], aphase);
        tc_fence_after();
        const uint32_t t_acc = tmem_base + as * BN + (static_cast<uint32_t>(q * 32) << 16);
        auto process = [&](uint32_t (&v)[32], int cc) {
          if constexpr (EPI == PXA_EPI_BIAS_RESIDUAL) {
            ResFrag res;
            load_residual_frag<OutT>(res, p, lane, m0 + q * 32, n0 + cc * 32);
            epilogue_chunk_residual<OutT>(v, res, p, stile, lane, m0 + q * 32, n0 + cc * 32);
          } else if constexpr (kLn) {
            epilogue_chunk_bf16_c<EPI>(v, p, cb, stile, lane, m0 + q * 32, n0 + cc * 32, cc * 32, ln, second);
          } else {
            epilogue_chunk_bf16_c<EPI>(v, p, cb, stile, lane, m0 + q * 32, n0 + cc * 32, cc * 32);
          }
        };
        auto release_acc = [&]() {
          tc_fence_before();
          mbar_arrive_cluster(&tempty_bar[as], 0);       // the leader's MMA thread owns the accumulator hand-off
        };
        // software pipeline: the TMEM load of this warp's next chunk is in flight while the current one is processed
        uint32_t va[32], vb[32];
        if (half < nch) tmem_ld_32x32b_x32_nowait(t_acc + half * 32, va); else release_acc();
#pragma unroll 1
        for (int cc = half; cc < nch; cc += 2 * kStep) {
          tmem_ld_wait_x32(va);
          if (cc + kStep < nch) tmem_ld_32x32b_x32_nowait(t_acc + (cc + kStep) * 32, vb); else release_acc();
          process(va, cc);
          if (cc + kStep < nch) {
            tmem_ld_wait_x32(vb);
            if (cc + 2 * kStep < nch) tmem_ld_32x32b_x32_nowait(t_acc + (cc + 2 * kStep) * 32, va); else release_acc();
            process(vb, cc + kStep);
          }
        }
        as ^= 1;
        if (as == 0) aphase ^= 1;
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();               // nobody may exit (or free TMEM) while the peer can still touch this CTA's smem
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair<Cfg::kTmemCols>(tmem_base);
  }
}

template <int BN, int EPI, typename OutT, bool kTmaRes = false, int kEpiWarps = 4>
static int launch_gemm2(const PxaGemmArgs& a, cudaStream_t stream) {
  using Cfg = Gemm2Cfg<BN, kTmaRes>;
  CUtensorMap ta, tw;
  {
    uint64_t dims[2] = {(uint64_t)a.K, (uint64_t)a.M};
    uint64_t str[1] = {(uint64_t)a.lda * 2};
    uint32_t box[2] = {kBK, kBM};
    int rc = make_tmap_bf16(&ta, a.a, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)a.K, (uint64_t)a.N};
    uint64_t str[1] = {(uint64_t)a.ldw * 2};
    uint32_t box[2] = {kBK, (uint32_t)(BN / 2)};
    int rc = make_tmap_bf16(&tw, a.w, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  CUtensorMap tr = ta, to = ta, tx = ta;
  if constexpr (kTmaRes) {
    uint64_t dims[2] = {(uint64_t)a.N, (uint64_t)a.M};
    uint64_t str[1] = {(uint64_t)a.ldo * 4};
    uint32_t box[2] = {32, kBM};           // 32 fp32 = 128 B rows: one TMA 128B-swizzle atom per row
    int rc = make_tmap(&tr, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, a.residual, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_tmap(&to, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, a.out, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    if (a.out_aux_bf16 != nullptr) {
      uint64_t astr[1] = {(uint64_t)a.ldo * 2};
      rc = make_tmap(&tx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, a.out_aux_bf16, 2, dims, astr, box, CU_TENSOR_MAP_SWIZZLE_64B);
      if (rc) return rc;
    }
  } else if constexpr (EPI == PXA_EPI_BIAS_RESIDUAL) {
    uint64_t dims[2] = {(uint64_t)a.N, (uint64_t)a.M};
    uint64_t str[1] = {(uint64_t)a.ldo * sizeof(OutT)};
    uint32_t box[2] = {(uint32_t)BN, kBM};
    int rc = make_tmap(&tr, sizeof(OutT) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                       a.residual, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
  }
  GemmParams p;
  p.bias = reinterpret_cast<const __nv_bfloat16*>(a.bias);
  p.out = a.out;
  p.out_aux = reinterpret_cast<__nv_bfloat16*>(a.out_aux_bf16);
  p.residual = a.residual;
  p.gate = a.gate;
  p.gate_batch_stride = a.gate_batch_stride;
  p.rows_per_batch = a.rows_per_batch > 0 ? a.rows_per_batch : a.M;
  p.M = a.M; p.N = a.N; p.K = a.K; p.ldo = a.ldo;
  p.num_m_tiles = (a.M + 2 * kBM - 1) / (2 * kBM);
  p.num_n_tiles = (a.N + BN - 1) / BN;
  p.trace = reinterpret_cast<long long*>(a.debug_trace);
  p.k_splits = 1;
  p.aux_branch = a.aux_is_branch;
  fill_ln_params(p, a);
  auto kern = gemm2_bf16_kernel<BN, EPI, OutT, kTmaRes, kEpiWarps>;
  PXA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem));
  int clusters = device_info().sms / 2;
  if (a.max_ctas > 0 && a.max_ctas / 2 < clusters) clusters = a.max_ctas / 2 > 0 ? a.max_ctas / 2 : 1;
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  if (tiles < clusters) clusters = tiles;
  kern<<<2 * clusters, 128 + 32 * kEpiWarps, Cfg::kSmem, stream>>>(ta, tw, tr, to, tx, p);
  launch_counter()++;
  PXA_CHECK_CUDA(cudaGetLastError());
  return PXA_OK;
}

// bf16 epilogues: 8 epilogue warps unless the caller asks for 4 (PxaGemmArgs.epi_warps; A/B measurements)
template <int BN, int EPI>
static int launch_gemm2_bf16(const PxaGemmArgs& a, cudaStream_t s) {
  if (a.epi_warps == 4) return launch_gemm2<BN, EPI, __nv_bfloat16, false, 4>(a, s);
  return launch_gemm2<BN, EPI, __nv_bfloat16, false, 8>(a, s);
}

template <int BN>
static int dispatch_epi2(const PxaGemmArgs& a, cudaStream_t s) {
  switch (a.epilogue) {
    case PXA_EPI_BIAS:
      return launch_gemm2_bf16<BN, PXA_EPI_BIAS>(a, s);
    case PXA_EPI_BIAS_GELU:
      return launch_gemm2_bf16<BN, PXA_EPI_BIAS_GELU>(a, s);
    case PXA_EPI_BIAS_GELU_AUX:
      return launch_gemm2_bf16<BN, PXA_EPI_BIAS_GELU_AUX>(a, s);
    case PXA_EPI_MUL_DGELU:
      return launch_gemm2_bf16<BN, PXA_EPI_MUL_DGELU>(a, s);
    case PXA_EPI_LN_BIAS:
      return launch_gemm2_bf16<BN, PXA_EPI_LN_BIAS>(a, s);
    case PXA_EPI_LN_BIAS_GELU:
      return launch_gemm2_bf16<BN, PXA_EPI_LN_BIAS_GELU>(a, s);
    case PXA_EPI_BIAS_RESIDUAL:
      if (a.out_dtype == PXA_DTYPE_F32) {
        // the TMA-streamed epilogue is required for the fused LayerNorm by-products (scaled aux copy, row statistics)
        const bool tma = a.res_epilogue == 2 || (a.res_epilogue == 0 && (a.aux_scale != nullptr || a.row_stats_out != nullptr));
        if (tma) return launch_gemm2<BN, PXA_EPI_BIAS_RESIDUAL, float, true>(a, s);
        return launch_gemm2<BN, PXA_EPI_BIAS_RESIDUAL, float>(a, s);
      }
      return launch_gemm2<BN, PXA_EPI_BIAS_RESIDUAL, __nv_bfloat16>(a, s);
    default:
      return fail(PXA_ERR_ARG, "unknown epilogue %d", a.epilogue);
  }
}

// Called by pxa_gemm_bf16 after argument validation when the CTA-pair path is selected.
int gemm_pair_dispatch(const PxaGemmArgs& a, int bn, cudaStream_t s) {
  switch (bn) {
    case 128: return dispatch_epi2<128>(a, s);
    case 192: return dispatch_epi2<192>(a, s);
    case 256: return dispatch_epi2<256>(a, s);
    default: return fail(PXA_ERR_ARG, "block_n must be 128, 192 or 256 (got %d)", bn);
  }
}

}  // namespace pxa
