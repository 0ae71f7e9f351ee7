// pxa_gemm_bf16: out = epilogue(A[M,K] . W[N,K]^T + bias) on tcgen05 tensor cores (sm_100a).
//
// Structure (one persistent CTA per SM, 256 threads):
//   warp 0      TMA producer: A tile 128x64 and W tile BNx64 (bf16, 128B-swizzled) into a kStages-deep smem ring
//   warp 1      MMA issuer: one thread issues tcgen05.mma 128xBNx16 (SS), fp32 accumulators in TMEM,
//               double-buffered accumulator (2 x BN columns) so the epilogue of tile i overlaps the MMAs of tile i+1
//   warp 2      TMEM allocator
//   warps 4-7   epilogue: tcgen05.ld (row-per-thread) -> bias / GELU / gate*x + residual -> smem transpose (XOR
//               swizzled, conflict-free) -> 128-bit coalesced global stores
// Tiles are visited n-fastest so the CTAs running concurrently share the same few A row-blocks through L2 and the
// whole weight matrix stays L2-resident.
//
// Algorithmic work: 2*M*N*K FLOP; bytes: M*K*2 (A) + N*K*2 (W) + M*N*out bytes (+ residual read).
#include "gemm_common.cuh"

namespace pxa {

// ------------------------------------------------------------------------------------------------- kernel
// kMn: both operands are given MN-major -- A as [K, M] (M contiguous), W as [K, N] (N contiguous) -- the layout of the
// weight-gradient product dW[N, K] = dY[M, N]^T X[M, K], whose contraction runs over the token dimension.  Each K-block
// is staged as 64-wide swizzle atoms of 64 k-rows (two for the 128 rows of A, BN / 64 for W) and consumed through
// MN-major UMMA descriptors (atoms LBO apart, 16 k-rows = 2048 B per K-step).  With p.k_splits > 1 the K range is split
// across tiles (split-K); legal only with the TMA reduce-add epilogue, which makes the partial sums add up in L2.
template <int BN, int EPI, typename OutT, bool kConv = false, bool kMn = false>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                 const __grid_constant__ CUtensorMap tmap_res, const __grid_constant__ CUtensorMap tmap_out,
                 const __grid_constant__ CUtensorMap tmap_aux, const GemmParams p) {
  constexpr bool kTmaRes = (EPI == PXA_EPI_BIAS_RESIDUAL) && sizeof(OutT) == 4;
  using Cfg = GemmCfg<BN, kTmaRes, kMn>;
  constexpr int kStages = Cfg::kStages;
  constexpr int kRing = Cfg::kResRing;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_smem = smem + kStages * Cfg::kStage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + Cfg::kEpiSmem);
  uint64_t* full_bar = bars;                    // [kStages]  TMA -> MMA
  uint64_t* empty_bar = bars + kStages;         // [kStages]  MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * kStages;     // [2]        MMA -> epilogue
  uint64_t* tempty_bar = bars + 2 * kStages + 2;  // [2]      epilogue -> MMA
  uint64_t* res_full = bars + 2 * kStages + 4;    // [kResBufs] TMA residual chunk landed (kTmaRes only)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4 + kResBufs);

  const int warp = warp_idx_sync();
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_w);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], kNumEpiThreads);
    }
    for (int s = 0; s < kResBufs; ++s) mbar_init(&res_full[s], 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int mn_tiles = p.num_m_tiles * p.num_n_tiles;
  const int num_tiles = mn_tiles * p.k_splits;
  const int num_kb = (p.K + kBK - 1) / kBK;
  // tile -> (k split, m tile, n tile), n fastest; K-blocks [kb_lo, kb_hi) of split ks
  auto kb_lo = [&](int tile) { return (int)((long long)(tile / mn_tiles) * num_kb / p.k_splits); };
  auto kb_hi = [&](int tile) { return (int)((long long)(tile / mn_tiles + 1) * num_kb / p.k_splits); };

  if (warp == 0) {
    // ================================================================ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = ((tile % mn_tiles) / p.num_n_tiles) * kBM;
        const int n0 = ((tile % mn_tiles) % p.num_n_tiles) * BN;
        if constexpr (EPI == PXA_EPI_BIAS_RESIDUAL) {
          // pull this tile's residual block into L2 now: the epilogue reads it one mainloop (~7k cycles) later
          if constexpr (kTmaRes) {
            if (!(p.out_aux == nullptr && p.residual == p.out && p.row_stats_out == nullptr))
              for (int c = 0; c < BN / 32; ++c) tma_prefetch_l2_2d(&tmap_res, n0 + c * 32, m0);
          } else if constexpr (kConv) {
            tma_prefetch_l2_2d(&tmap_res, n0, (m0 / p.conv_Wp) * p.conv_W + m0 % p.conv_Wp);   // real row of the tile's first pixel
          } else {
            tma_prefetch_l2_2d(&tmap_res, n0, m0);
          }
        }
        for (int kb = kb_lo(tile); kb < kb_hi(tile); ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::kStage;
          uint8_t* sb = sa + Cfg::kStageA;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStage);
          if constexpr (kMn) {
#pragma unroll
            for (int i = 0; i < kBM / 64; ++i) tma_load_2d(sa + i * 8192, &tmap_a, &full_bar[stage], m0 + 64 * i, kb * kBK, kEvictNormal);
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_2d(sb + j * 8192, &tmap_w, &full_bar[stage], n0 + 64 * j, kb * kBK, kEvictNormal);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
            continue;
          }
          if constexpr (kConv) {
            // implicit GEMM: K-block kb = (tap, 64-channel slice); the A tile is the tile's pixel block shifted by the
            // tap offset -- zero padding comes for free from TMA out-of-bounds fill on the W / H dimensions
            const int tap = kb / p.conv_cin_blocks;
            const int c0 = (kb - tap * p.conv_cin_blocks) * kBK;
            const int pix = m0;                                     // first pixel of the tile in (b, y, x) order of the virtual image
            const int x0 = pix % p.conv_Wp;
            const int y0 = (pix / p.conv_Wp) % p.conv_H;
            const int b0 = pix / (p.conv_Wp * p.conv_H);
            tma_load_4d(sa, &tmap_a, &full_bar[stage], c0, x0 + tap % 3 - 1, y0 + tap / 3 - 1, b0, kEvictNormal);
          } else {
            tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * kBK, m0, kEvictNormal);
          }
          tma_load_2d(sb, &tmap_w, &full_bar[stage], kb * kBK, n0, kEvictLast);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(kBM, BN, kMn ? 1 : 0, kMn ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        const int kb0 = kb_lo(tile), kb1 = kb_hi(tile);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStage);
          // K-major: 8-row groups 1024 B apart; MN-major: 64-wide atoms 8192 B apart (LBO), 8 k-rows = 1024 B (SBO)
          const uint64_t adesc = make_smem_desc(sa, kMn ? 8192 : 16, 1024, kLayoutSW128);
          const uint64_t bdesc = make_smem_desc(sa + Cfg::kStageA, kMn ? 8192 : 16, 1024, kLayoutSW128);
          // one K-step = 16 elements along K: 32 B inside the swizzle atom (K-major) / 16 rows of 128 B (MN-major)
          constexpr int kStep = kMn ? (2048 >> 4) : 2;
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            umma_ss(d_tmem, adesc + kStep * k, bdesc + kStep * k, idesc, ((kb - kb0) | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[as]);
        as ^= 1;
        if (as == 0) aphase ^= 1;
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ================================================================ epilogue
    EpiConst* consts = reinterpret_cast<EpiConst*>(epi_smem + Cfg::kEpiBufs);
    if constexpr (kTmaRes) {
      // fp32 residual stream: TMA-streamed residual chunks, row per thread (gemm_common.cuh)
      TileWalk tw;
      tw.first = blockIdx.x; tw.stride = gridDim.x; tw.count = num_tiles;
      tw.mn_tiles = mn_tiles; tw.num_n_tiles = p.num_n_tiles; tw.m_mult = kBM; tw.m_off = 0;
      tw.reverse = 0;                 // (reverse_tiles is a CTA-pair kernel option)
      tma_res_epilogue<BN, kRing, false>(p, tw, epi_smem, consts, tfull_bar, tempty_bar, res_full, tmem_base, &tmap_res,
                                         &tmap_out, &tmap_aux, warp, lane);
    } else {
      constexpr bool kLn = (EPI == PXA_EPI_LN_BIAS || EPI == PXA_EPI_LN_BIAS_GELU);
      const int q = warp & 3;                         // TMEM sub-partition of this warp
      const int tid = threadIdx.x - kEpiWarp0 * 32;   // 0..127
      int as = 0;
      uint32_t aphase = 0;
      int titer = 0;
      uint8_t* stile = epi_smem + (warp - kEpiWarp0) * 4096;       // per-warp transpose tile
      EpiRegs<BN> er;                                              // constants / LN statistics of a tile, loaded one tile ahead
      [[maybe_unused]] LnStatRegs sr;
      if ((int)blockIdx.x < num_tiles) {
        const int ft = blockIdx.x;
        load_epi_consts<BN, kLn>(er, p, tid, ((ft % mn_tiles) / p.num_n_tiles) * kBM, ((ft % mn_tiles) % p.num_n_tiles) * BN);
        if constexpr (kLn) load_ln_stats(sr, p, ((ft % mn_tiles) / p.num_n_tiles) * kBM + q * 32 + lane);
      }
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++titer) {
        const int m0 = ((tile % mn_tiles) / p.num_n_tiles) * kBM;
        const int n0 = ((tile % mn_tiles) % p.num_n_tiles) * BN;
        const int nch = chunks_of_tile<BN>(p, n0);
        EpiConst* cb = consts + (titer & 1);
        store_epi_consts<BN>(cb, er, tid);
        [[maybe_unused]] float2 ln = make_float2(1.f, 0.f);
        if constexpr (kLn) ln = ln_row_coeffs(p, sr);
        named_bar_sync(2, kNumEpiThreads);
        if (tile + gridDim.x < num_tiles) {                               // next tile's loads: in flight during this tile's chunks
          const int nt_ = tile + gridDim.x;
          load_epi_consts<BN, kLn>(er, p, tid, ((nt_ % mn_tiles) / p.num_n_tiles) * kBM, ((nt_ % mn_tiles) % p.num_n_tiles) * BN);
          if constexpr (kLn) load_ln_stats(sr, p, ((nt_ % mn_tiles) / p.num_n_tiles) * kBM + q * 32 + lane);
        }
        [[maybe_unused]] const bool second = q * 32 + lane >= cb->row_split;
        mbar_wait(&tfull_bar[as], aphase);
        tc_fence_after();
        const uint32_t t_acc = tmem_base + as * BN + (static_cast<uint32_t>(q * 32) << 16);
        // this warp's 32 output rows: GEMM rows m0 + 32 q ..; a convolution over a virtual width (conv_Wp > W: one image line per
        // tile) stores pixel x of line (b, y) at row (b H + y) W + x and drops the columns x >= W
        int row0 = m0 + q * 32;
        [[maybe_unused]] int m_lim = -1;
        if constexpr (kConv) {
          if (p.conv_Wp != p.conv_W) {
            const int xs = m0 % p.conv_Wp + q * 32;
            const int left = p.conv_W - xs;
            row0 = (m0 / p.conv_Wp) * p.conv_W + xs;
            m_lim = row0 + (left < 0 ? 0 : (left > 32 ? 32 : left));
          }
        }
        auto process = [&](uint32_t (&v)[32], int cc) {
          if constexpr (EPI == PXA_EPI_BIAS_RESIDUAL) {
            ResFrag res;                                            // bf16 residual stream (VAE convolutions): simple path
            load_residual_frag<OutT>(res, p, lane, row0, n0 + cc * 32, m_lim);
            epilogue_chunk_residual<OutT>(v, res, p, stile, lane, row0, n0 + cc * 32, m_lim);
          } else if constexpr (kLn) {
            epilogue_chunk_bf16_c<EPI>(v, p, cb, stile, lane, row0, n0 + cc * 32, cc * 32, ln, second);
          } else {
            epilogue_chunk_bf16_c<EPI>(v, p, cb, stile, lane, row0, n0 + cc * 32, cc * 32, make_float2(1.f, 0.f), false, m_lim);
          }
        };
        auto release_acc = [&]() {                                  // all TMEM reads of this accumulator are done
          tc_fence_before();
          mbar_arrive(&tempty_bar[as]);
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
        as ^= 1;
        if (as == 0) aphase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------- host
struct ConvGeom {
  int B, H, W, Cin, tile_w, tile_h;
  int Wp;          // virtual width of the M index (= W when W tiles; else W rounded up to 128, see GemmParams::conv_Wp)
};

template <int BN, int EPI, typename OutT, bool kConv = false, bool kMn = false>
static int launch_gemm(const PxaGemmArgs& a, cudaStream_t stream, const ConvGeom* cg = nullptr, int k_splits = 1) {
  CUtensorMap ta, tw;
  if constexpr (kMn) {
    // A stored [K, M] (row stride lda), W stored [K, N] (row stride ldw): 64 x 64 boxes, M / N the contiguous dimension
    uint64_t dims_a[2] = {(uint64_t)a.M, (uint64_t)a.K}, dims_w[2] = {(uint64_t)a.N, (uint64_t)a.K};
    uint64_t str_a[1] = {(uint64_t)a.lda * 2}, str_w[1] = {(uint64_t)a.ldw * 2};
    uint32_t box[2] = {64, kBK};
    int rc = make_tmap_bf16(&ta, a.a, 2, dims_a, str_a, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_tmap_bf16(&tw, a.w, 2, dims_w, str_w, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  } else if constexpr (kConv) {
    // NHWC image as a 4-D tensor (c, x, y, b); one A tile = tile_h rows x tile_w pixels x 64 channels
    uint64_t dims[4] = {(uint64_t)cg->Cin, (uint64_t)cg->W, (uint64_t)cg->H, (uint64_t)cg->B};
    uint64_t str[3] = {(uint64_t)cg->Cin * 2, (uint64_t)cg->W * cg->Cin * 2, (uint64_t)cg->H * cg->W * cg->Cin * 2};
    uint32_t box[4] = {kBK, (uint32_t)cg->tile_w, (uint32_t)cg->tile_h, 1};
    int rc = make_tmap_bf16(&ta, a.a, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  } else {
    uint64_t dims[2] = {(uint64_t)a.K, (uint64_t)a.M};
    uint64_t str[1] = {(uint64_t)a.lda * 2};
    uint32_t box[2] = {kBK, kBM};
    int rc = make_tmap_bf16(&ta, a.a, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  if constexpr (!kMn) {
    uint64_t dims[2] = {(uint64_t)a.K, (uint64_t)a.N};
    uint64_t str[1] = {(uint64_t)a.ldw * 2};
    uint32_t box[2] = {kBK, (uint32_t)BN};
    int rc = make_tmap_bf16(&tw, a.w, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  constexpr bool kTmaRes = (EPI == PXA_EPI_BIAS_RESIDUAL) && sizeof(OutT) == 4;
  using Cfg = GemmCfg<BN, kTmaRes, kMn>;
  CUtensorMap tr = ta, to = ta, tx = ta;   // residual / out / aux maps: only meaningful for EPI_BIAS_RESIDUAL
  if constexpr (EPI == PXA_EPI_BIAS_RESIDUAL) {
    uint64_t dims[2] = {(uint64_t)a.N, (uint64_t)a.M};
    if constexpr (kConv) dims[1] = (uint64_t)cg->B * cg->H * cg->W;       // the rows that exist (a.M counts the virtual width)
    uint64_t str[1] = {(uint64_t)a.ldo * sizeof(OutT)};
    if constexpr (kTmaRes) {
      uint32_t box[2] = {32, kBM};         // 32 fp32 = 128 B rows: one TMA 128B-swizzle atom per row
      int rc = make_tmap(&tr, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, a.residual, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc) return rc;
      rc = make_tmap(&to, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, a.out, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc) return rc;
      if (a.out_aux_bf16 != nullptr) {
        uint64_t astr[1] = {(uint64_t)a.ldo * 2};
        rc = make_tmap(&tx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, a.out_aux_bf16, 2, dims, astr, box, CU_TENSOR_MAP_SWIZZLE_64B);
        if (rc) return rc;
      }
    } else {
      uint32_t box[2] = {(uint32_t)BN, kBM};   // L2 prefetch only
      int rc = make_tmap(&tr, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, a.residual, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE);
      if (rc) return rc;
    }
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
  p.num_m_tiles = (a.M + kBM - 1) / kBM;
  p.num_n_tiles = (a.N + BN - 1) / BN;
  p.trace = reinterpret_cast<long long*>(a.debug_trace);
  p.k_splits = k_splits;
  p.aux_branch = a.aux_is_branch;
  fill_ln_params(p, a);
  p.conv_H = p.conv_W = p.conv_tile_w = p.conv_tile_h = p.conv_cin_blocks = 0;
  if constexpr (kConv) {
    p.conv_H = cg->H; p.conv_W = cg->W; p.conv_tile_w = cg->tile_w; p.conv_tile_h = cg->tile_h;
    p.conv_cin_blocks = cg->Cin / kBK;
  }
  p.conv_Wp = kConv ? cg->Wp : 0;
  auto kern = gemm_bf16_kernel<BN, EPI, OutT, kConv, kMn>;
  PXA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem));
  int grid = device_info().sms;
  if (a.max_ctas > 0 && a.max_ctas < grid) grid = a.max_ctas;
  const int tiles = p.num_m_tiles * p.num_n_tiles * p.k_splits;
  if (tiles < grid) grid = tiles;
  kern<<<grid, kGemmThreads, Cfg::kSmem, stream>>>(ta, tw, tr, to, tx, p);
  launch_counter()++;
  PXA_CHECK_CUDA(cudaGetLastError());
  return PXA_OK;
}

template <int BN>
static int dispatch_epi(const PxaGemmArgs& a, cudaStream_t s) {
  switch (a.epilogue) {
    case PXA_EPI_BIAS:
      if (a.out_dtype != PXA_DTYPE_BF16) return fail(PXA_ERR_ARG, "EPI_BIAS writes bf16 only");
      return launch_gemm<BN, PXA_EPI_BIAS, __nv_bfloat16>(a, s);
    case PXA_EPI_BIAS_GELU:
      if (a.out_dtype != PXA_DTYPE_BF16) return fail(PXA_ERR_ARG, "EPI_BIAS_GELU writes bf16 only");
      return launch_gemm<BN, PXA_EPI_BIAS_GELU, __nv_bfloat16>(a, s);
    case PXA_EPI_BIAS_GELU_AUX:
      return launch_gemm<BN, PXA_EPI_BIAS_GELU_AUX, __nv_bfloat16>(a, s);
    case PXA_EPI_MUL_DGELU:
      return launch_gemm<BN, PXA_EPI_MUL_DGELU, __nv_bfloat16>(a, s);
    case PXA_EPI_LN_BIAS:
      return launch_gemm<BN, PXA_EPI_LN_BIAS, __nv_bfloat16>(a, s);
    case PXA_EPI_LN_BIAS_GELU:
      return launch_gemm<BN, PXA_EPI_LN_BIAS_GELU, __nv_bfloat16>(a, s);
    case PXA_EPI_BIAS_RESIDUAL:
      if (a.residual == nullptr) return fail(PXA_ERR_ARG, "EPI_BIAS_RESIDUAL needs residual");
      if (a.out_dtype == PXA_DTYPE_F32) return launch_gemm<BN, PXA_EPI_BIAS_RESIDUAL, float>(a, s);
      return launch_gemm<BN, PXA_EPI_BIAS_RESIDUAL, __nv_bfloat16>(a, s);
    default:
      return fail(PXA_ERR_ARG, "unknown epilogue %d", a.epilogue);
  }
}

int gemm_pair_dispatch(const PxaGemmArgs& a, int bn, cudaStream_t s);   // gemm2_sm100.cu

// Weight-gradient form: out[M, N] += A^T W with A [K, M], W [K, N] (both MN-major), fp32 out accumulated in place through
// the TMA reduce-add epilogue, K split over tiles so that the persistent grid is evenly loaded (the layer shapes give
// only 54 .. 216 output tiles against 148 SMs; each tile's K loop is thousands of blocks long).
template <int BN>
static int launch_wgrad(const PxaGemmArgs& a, cudaStream_t s) {
  const int tiles = ((a.M + kBM - 1) / kBM) * ((a.N + BN - 1) / BN);
  const int num_kb = (a.K + kBK - 1) / kBK;
  int sms = device_info().sms;
  if (a.max_ctas > 0 && a.max_ctas < sms) sms = a.max_ctas;
  int best = 1;
  double best_eff = 0.0;
  for (int ks = 1; ks <= 16 && num_kb / ks >= 8; ++ks) {
    const int t = tiles * ks;
    const double eff = (double)t / (double)(((t + sms - 1) / sms) * sms);
    if (eff > best_eff + 1e-9) { best_eff = eff; best = ks; }
  }
  if (a.k_splits > 0) best = a.k_splits;
  if (best > num_kb) best = num_kb;
  return launch_gemm<BN, PXA_EPI_BIAS_RESIDUAL, float, false, true>(a, s, nullptr, best);
}

template <int BN>
static int dispatch_conv(const PxaGemmArgs& a, cudaStream_t s, const ConvGeom& cg) {
  if (a.epilogue == PXA_EPI_BIAS_RESIDUAL) return launch_gemm<BN, PXA_EPI_BIAS_RESIDUAL, __nv_bfloat16, true>(a, s, &cg);
  return launch_gemm<BN, PXA_EPI_BIAS, __nv_bfloat16, true>(a, s, &cg);
}

}  // namespace pxa

extern "C" int pxa_conv3x3_nhwc_bf16(const PxaConv3x3Args* args, void* stream) {
  using namespace pxa;
  if (!args) return fail(PXA_ERR_ARG, "null args");
  const PxaConv3x3Args& c = *args;
  if (!c.x || !c.w || !c.out) return fail(PXA_ERR_ARG, "null x / w / out");
  if (c.B <= 0 || c.H <= 0 || c.W <= 0 || c.Cin <= 0 || c.Cout <= 0) return fail(PXA_ERR_ARG, "bad shape");
  if (c.Cin % 64) return fail(PXA_ERR_ARG, "Cin must be a multiple of 64 (got %d)", c.Cin);
  if (c.Cout % 8) return fail(PXA_ERR_ARG, "Cout must be a multiple of 8 (got %d)", c.Cout);
  ConvGeom cg;
  cg.B = c.B; cg.H = c.H; cg.W = c.W; cg.Cin = c.Cin;
  // M tile = 128 consecutive pixels = ONE TMA box: tile_h full lines when W is a power of two <= 128 (and H a multiple of tile_h),
  // 128 pixels of one line when W is a multiple of 128.  Every other geometry (e.g. the 1024-MS aspect buckets: latent 104 x 152 ->
  // 832 x 1216) runs over a virtual width Wp = W rounded up to 128, one line segment per tile: columns >= W are zero fill going in
  // (the right-hand padding of the convolution, exactly) and are dropped going out; the tail tile of a line is partly idle (W / Wp of
  // the MMA work is useful: 0.81 at W = 104, 0.95 at W = 1216).
  cg.tile_w = c.W < 128 ? c.W : 128;
  cg.tile_h = (128 % cg.tile_w) ? 1 : 128 / cg.tile_w;
  cg.Wp = c.W;
  if ((128 % cg.tile_w) || c.W % cg.tile_w || c.H % cg.tile_h) {
    cg.tile_w = 128; cg.tile_h = 1;
    cg.Wp = (c.W + 127) / 128 * 128;
  }
  if ((reinterpret_cast<uintptr_t>(c.x) | reinterpret_cast<uintptr_t>(c.w) | reinterpret_cast<uintptr_t>(c.out) |
       reinterpret_cast<uintptr_t>(c.bias) | reinterpret_cast<uintptr_t>(c.residual)) & 15)
    return fail(PXA_ERR_ALIGN, "pointers must be 16-byte aligned");
  PXA_REQUIRE_SM100();
  PxaGemmArgs a = {};
  a.a = c.x; a.w = c.w; a.bias = c.bias; a.out = c.out; a.residual = c.residual;
  if ((long long)c.B * c.H * cg.Wp >= (1ll << 31)) return fail(PXA_ERR_ARG, "B x H x W too large");
  a.M = c.B * c.H * cg.Wp; a.N = c.Cout; a.K = 9 * c.Cin;
  a.lda = c.Cin; a.ldw = 9 * c.Cin; a.ldo = c.Cout;
  a.rows_per_batch = a.M;
  a.epilogue = c.residual ? PXA_EPI_BIAS_RESIDUAL : PXA_EPI_BIAS;
  a.out_dtype = PXA_DTYPE_BF16;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (c.Cout % 256 == 0) return dispatch_conv<256>(a, s, cg);
  if (c.Cout % 192 == 0) return dispatch_conv<192>(a, s, cg);
  return dispatch_conv<128>(a, s, cg);
}

extern "C" int pxa_gemm_bf16(const PxaGemmArgs* args, void* stream) {
  using namespace pxa;
  if (!args) return fail(PXA_ERR_ARG, "null args");
  const PxaGemmArgs& a = *args;
  if (!a.a || !a.w || !a.out) return fail(PXA_ERR_ARG, "null a / w / out");
  if (a.M <= 0 || a.N <= 0 || a.K <= 0) return fail(PXA_ERR_ARG, "bad shape M=%d N=%d K=%d", a.M, a.N, a.K);
  if ((a.N & 7) || (!a.operands_mn_major && (a.K & 7)) || (a.operands_mn_major && (a.M & 7)) || (a.lda & 7) ||
      (a.ldw & 7) || (a.ldo & 7))
    return fail(PXA_ERR_ALIGN, "N, the contiguous operand dimension (K; M for operands_mn_major), lda, ldw, ldo must be "
                               "multiples of 8 (M=%d N=%d K=%d lda=%d ldw=%d ldo=%d)", a.M, a.N, a.K, a.lda, a.ldw, a.ldo);
  if (a.k_splits < 0 || (a.k_splits > 1 && !a.operands_mn_major)) return fail(PXA_ERR_ARG, "k_splits needs operands_mn_major");
  if ((reinterpret_cast<uintptr_t>(a.out) | reinterpret_cast<uintptr_t>(a.bias) |
       reinterpret_cast<uintptr_t>(a.residual) | reinterpret_cast<uintptr_t>(a.gate) |
       reinterpret_cast<uintptr_t>(a.out_aux_bf16)) & 15)
    return fail(PXA_ERR_ALIGN, "out / bias / residual / gate / aux must be 16-byte aligned");
  if (a.gate && (a.gate_batch_stride & 3)) return fail(PXA_ERR_ALIGN, "gate_batch_stride must be a multiple of 4");
  PXA_REQUIRE_SM100();
  int bn = a.block_n;
  if (bn == 0) bn = (a.N % 192 == 0) ? 192 : ((a.N % 256 == 0) ? 256 : (a.N >= 192 ? 192 : 128));
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (a.operands_mn_major) {
    if (a.epilogue != PXA_EPI_BIAS_RESIDUAL || a.out_dtype != PXA_DTYPE_F32 || a.residual != a.out || a.bias || a.gate ||
        a.out_aux_bf16)
      return fail(PXA_ERR_ARG, "operands_mn_major is the weight-gradient form: EPI_BIAS_RESIDUAL, fp32 out, residual == out, "
                               "no bias / gate / aux");
    switch (bn) {
      case 128: return launch_wgrad<128>(a, s);
      case 192: return launch_wgrad<192>(a, s);
      case 256: return launch_wgrad<256>(a, s);
      default: return fail(PXA_ERR_ARG, "block_n must be 0, 128, 192 or 256 (got %d)", bn);
    }
  }
  if (a.epilogue == PXA_EPI_BIAS_RESIDUAL && a.residual == nullptr) return fail(PXA_ERR_ARG, "EPI_BIAS_RESIDUAL needs residual");
  if (a.epilogue != PXA_EPI_BIAS_RESIDUAL && a.out_dtype != PXA_DTYPE_BF16)
    return fail(PXA_ERR_ARG, "EPI_BIAS / EPI_BIAS_GELU write bf16 only");
  if (a.epilogue == PXA_EPI_BIAS_GELU_AUX && a.out_aux_bf16 == nullptr) return fail(PXA_ERR_ARG, "EPI_BIAS_GELU_AUX needs out_aux_bf16");
  if (a.epilogue == PXA_EPI_MUL_DGELU && (a.residual == nullptr || a.bias != nullptr))
    return fail(PXA_ERR_ARG, "EPI_MUL_DGELU needs the pre-activation in `residual` and no bias");
  if (a.cta_pair < 0 || a.cta_pair > 2) return fail(PXA_ERR_ARG, "cta_pair must be 0, 1 or 2");
  // fused LayerNorm-modulate: consumer epilogues and the producer by-products of the fp32 residual epilogue
  const bool ln_epi = a.epilogue == PXA_EPI_LN_BIAS || a.epilogue == PXA_EPI_LN_BIAS_GELU;
  const int rpb = a.rows_per_batch > 0 ? a.rows_per_batch : a.M;
  if (ln_epi) {
    if (!a.ln_stats || !a.ln_u || !a.ln_v || a.ln_dim <= 0 || a.bias)
      return fail(PXA_ERR_ARG, "EPI_LN_*: ln_stats / ln_u / ln_v / ln_dim required and bias must be NULL (it is part of ln_v)");
    if ((reinterpret_cast<uintptr_t>(a.ln_stats) | reinterpret_cast<uintptr_t>(a.ln_u) | reinterpret_cast<uintptr_t>(a.ln_v)) & 15)
      return fail(PXA_ERR_ALIGN, "ln_stats / ln_u / ln_v must be 16-byte aligned");
  }
  if (a.aux_scale || a.row_stats_out) {
    if (a.epilogue != PXA_EPI_BIAS_RESIDUAL || a.out_dtype != PXA_DTYPE_F32)
      return fail(PXA_ERR_ARG, "aux_scale / row_stats_out belong to the fp32 residual epilogue");
    if (a.aux_scale && !a.out_aux_bf16) return fail(PXA_ERR_ARG, "aux_scale needs out_aux_bf16");
    if (a.res_epilogue == 1) return fail(PXA_ERR_ARG, "aux_scale / row_stats_out need the TMA-streamed residual epilogue");
    if ((reinterpret_cast<uintptr_t>(a.aux_scale) | reinterpret_cast<uintptr_t>(a.row_stats_out)) & 15)
      return fail(PXA_ERR_ALIGN, "aux_scale / row_stats_out must be 16-byte aligned");
  }
  if ((ln_epi || a.aux_scale) && rpb < kBM && a.M > rpb)
    return fail(PXA_ERR_ARG, "the fused LayerNorm-modulate needs rows_per_batch >= %d (a tile may span two samples at most)", kBM);
  if (a.res_epilogue < 0 || a.res_epilogue > 2) return fail(PXA_ERR_ARG, "res_epilogue must be 0, 1 or 2");
  // auto (measured at M = 32768, tools/gemm_bench.py): the CTA pair with 256 x 256 tiles wins for every bias / GELU
  // GEMM; the fp32 residual epilogue is HBM-bound at K = 1152 and streams through TMA chunk buffers on the single-CTA
  // kernel, while at K >= 2304 it is MMA-bound and the pair kernel's deeper smem ring (7 stages) wins.
  bool pair = a.cta_pair == 2;
  PxaGemmArgs b = a;                    // `res_epilogue` may be filled in by the auto choice below
  if (a.cta_pair == 0 && a.block_n == 0 && a.M >= 1024) {
    if (a.epilogue != PXA_EPI_BIAS_RESIDUAL) {
      pair = true;
      bn = 256;
    } else if (a.out_dtype == PXA_DTYPE_F32 && a.residual == a.out && !a.out_aux_bf16 && !a.row_stats_out && a.res_epilogue != 1) {
      // in-place update of the fp32 stream with no by-product: CTA pair + TMA reduce-add epilogue (round 2, M = 32768:
      // 94 us vs 103 us single-CTA at K = 1152; 255 us vs 268 us with the register-staged pair epilogue at K = 4608)
      pair = true;
      bn = a.K >= 2304 ? 192 : 256;
      b.res_epilogue = 2;
    } else if (a.K >= 2304) {
      pair = true;
      bn = 192;
    } else {
      bn = 256;      // read-modify-write with a bf16 copy: single CTA, TMA-streamed residual chunks (123 us vs 129 us pair)
    }
  }
  if (a.row_stats_out && (a.N + bn - 1) / bn > PXA_LN_STAT_PARTS)
    return fail(PXA_ERR_ARG, "row_stats_out: N / block_n = %d column tiles exceed PXA_LN_STAT_PARTS", (a.N + bn - 1) / bn);
  if (pair) return gemm_pair_dispatch(b, bn, s);
  switch (bn) {
    case 128: return dispatch_epi<128>(a, s);
    case 192: return dispatch_epi<192>(a, s);
    case 256: return dispatch_epi<256>(a, s);
    default: return fail(PXA_ERR_ARG, "block_n must be 0, 128, 192 or 256 (got %d)", bn);
  }
}
