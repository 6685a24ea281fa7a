// CTA-pair (cta_group::2) tcgen05 GEMM kernel for the rank-2 MLP GEMMs; see the banner below and DESIGN.md 3.1.
#include <stdlib.h>
#include "gemm_epilogue.cuh"

namespace {

// ================================================================================================================
// CTA-pair (cta_group::2) kernel for the rank-2 MLP GEMMs: a cluster of two CTAs computes a 256 x 256 output tile.
// Each CTA loads ITS 128 rows of A and ITS 128 of the 256 weight rows (32 KB per k-chunk instead of 48 KB: the
// cta_group::1 mainloops of stages 2/3 are bound by the L2 -> SM operand stream), the leader CTA's
// elected thread issues tcgen05.mma.cta_group::2 (M = 256: every CTA accumulates its own 128 x 256 block in its own
// TMEM), and both CTAs run the TMA-store epilogue on their rows.  Barriers:
//   full[s]   (leader)   <- expect_tx by the leader's producer + complete_tx of BOTH CTAs' TMA loads
//   empty[s]  (each CTA) <- tcgen05.commit multicast to both CTAs when the MMAs that read slot s retire
//   tfull[a]  (each CTA) <- tcgen05.commit multicast after the last k-chunk of a tile
//   tempty[a] (leader)   <- the 8 epilogue warps of BOTH CTAs (remote mbarrier arrive from the peer)
// ================================================================================================================
// NEW = epilogue warps per CTA: 8 (6-stage ring) or 16 (4-stage ring; four warps per scheduler hide the latency of the
// GELU epilogue, which with two warps per scheduler runs at ~0.5 IPC and bounds the short-K fc1 GEMMs).
constexpr int P2_B_BYTES = 128 * BLOCK_K * 2;                 // this CTA's half of the 256 weight rows
constexpr int P2_STAGE_BYTES = A_STAGE_BYTES + P2_B_BYTES;    // 32 KB
template <int NEW> struct P2Cfg {
  static constexpr int STAGES = NEW == 16 ? 4 : 6;
  static constexpr int THREADS = 128 + 32 * NEW;
  static constexpr int SMEM_BYTES = STAGES * P2_STAGE_BYTES + NEW * EPI_STAGE_BYTES + 256 + 1024;
};

template <int EPI, int NEW>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(P2Cfg<NEW>::THREADS, 1) gemm_pair_kernel(const __grid_constant__ GemmPlan p) {
  constexpr int BLOCK_N = 256;
  constexpr int P2_STAGES = P2Cfg<NEW>::STAGES;
  constexpr int NUM_EPI_WARPS = NEW;   // shadows the file-scope constant inside this kernel
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + P2_STAGES * P2_STAGE_BYTES + NUM_EPI_WARPS * EPI_STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (P2_STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * P2_STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * P2_STAGES + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * P2_STAGES + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(
      smem_gen + P2_STAGES * P2_STAGE_BYTES + NUM_EPI_WARPS * EPI_STAGE_BYTES + 8 * (2 * P2_STAGES + 4));
  uint8_t* stage_base = smem_gen + P2_STAGES * P2_STAGE_BYTES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();   // 0 = leader
  ptx::griddep_launch();
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&p.tmap_a);
    ptx::prefetch_tmap(&p.tmap_b);
    ptx::prefetch_tmap(&p.tmap_out);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < P2_STAGES; ++s) { ptx::mbar_init(full_bar(s), 1); ptx::mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { ptx::mbar_init(tfull_bar(s), 1); ptx::mbar_init(tempty_bar(s), 2 * NUM_EPI_WARPS); }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc_pair(tmem_slot, 512);
    ptx::tmem_relinquish_pair();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();   // both CTAs' barriers and TMEM exist before anything crosses the pair
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  ptx::griddep_wait();

  const int m_pairs = (p.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
  const int total_tiles = m_pairs * p.n_tiles;
  const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int k_iters = p.k_chunks;

  if (warp == 0 && lane == 0) {
    // ================= TMA producer (both CTAs) =================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = pair_id; tile < total_tiles; tile += num_pairs) {
      const int n_tile = tile % p.n_tiles, m_pair = tile / p.n_tiles;
      const int arow = (m_pair * 2 + (int)rank) * BLOCK_M;
      const int brow = n_tile * BLOCK_N + (int)rank * 128;
      for (int kc = 0; kc < k_iters; ++kc) {
        ptx::mbar_wait(empty_bar(stage), phase ^ 1);
        const uint32_t sa = smem_base + stage * P2_STAGE_BYTES;
        const uint32_t sb = sa + A_STAGE_BYTES;
        const uint32_t lead_full = ptx::mapa_shared(full_bar(stage), 0);
        if (rank == 0) ptx::mbar_arrive_expect_tx(full_bar(stage), 2 * P2_STAGE_BYTES);
        ptx::tma_load_2d_pair(sa, &p.tmap_a, lead_full, kc * BLOCK_K, arow);
        ptx::tma_load_2d_pair(sb, &p.tmap_b, lead_full, kc * BLOCK_K, brow);
        if (++stage == P2_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && lane == 0 && rank == 0) {
    // ================= MMA issuer (leader CTA only) =================
    constexpr uint32_t idesc = ptx::make_idesc_bf16(2 * BLOCK_M, BLOCK_N);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = pair_id; tile < total_tiles; tile += num_pairs, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      ptx::mbar_wait(tempty_bar(as), aphase ^ 1);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * 256;
      for (int k = 0; k < k_iters; ++k) {
        ptx::mbar_wait(full_bar(stage), phase);
        ptx::tc_fence_after();
        const uint32_t sa = smem_base + stage * P2_STAGE_BYTES;
        const uint64_t adesc = ptx::make_sw128_kmajor_desc(sa);
        const uint64_t bdesc = ptx::make_sw128_kmajor_desc(sa + A_STAGE_BYTES);
#pragma unroll
        for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk)
          ptx::tc_mma_bf16_pair(d_tmem, adesc + 2u * kk, bdesc + 2u * kk, idesc, (k | kk) != 0 ? 1u : 0u);
        ptx::tc_commit_pair(empty_bar(stage), 0x3);   // slot reusable in BOTH CTAs
        if (++stage == P2_STAGES) { stage = 0; phase ^= 1; }
      }
      ptx::tc_commit_pair(tfull_bar(as), 0x3);         // accumulators complete in both CTAs
    }
  } else if (warp >= 4) {
    // ================= epilogue (both CTAs) =================
    const int ew = warp - 4;
    int it = 0;
    for (int tile = pair_id; tile < total_tiles; tile += num_pairs, ++it) {
      const int n_tile = tile % p.n_tiles, m_pair = tile / p.n_tiles;
      const int m_tile = m_pair * 2 + (int)rank;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      ptx::mbar_wait(tfull_bar(as), aphase);
      ptx::tc_fence_after();
      constexpr bool kF32 = (EPI == EPI_RESID);
      if constexpr (EPI == EPI_STORE) {
        if (p.out_f32) epilogue_tile_tma<BLOCK_N, EPI, true, NEW>(p, m_tile, n_tile, tmem_base + as * 256, ew, lane, stage_base + ew * EPI_STAGE_BYTES);
        else epilogue_tile_tma<BLOCK_N, EPI, false, NEW>(p, m_tile, n_tile, tmem_base + as * 256, ew, lane, stage_base + ew * EPI_STAGE_BYTES);
      } else {
        epilogue_tile_tma<BLOCK_N, EPI, kF32, NEW>(p, m_tile, n_tile, tmem_base + as * 256, ew, lane, stage_base + ew * EPI_STAGE_BYTES);
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa_shared(tempty_bar(as), 0));
    }
    if (lane == 0) ptx::bulk_wait0();
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();   // nobody frees TMEM or exits while the peer can still touch this CTA
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_pair(tmem_base, 512);
  }
}

template <int EPI, int NEW>
int launch_pair(const GemmPlan& plan, cudaStream_t stream) {
  auto kfn = gemm_pair_kernel<EPI, NEW>;
  constexpr int P2_SMEM_BYTES = P2Cfg<NEW>::SMEM_BYTES;
  constexpr int NUM_THREADS = P2Cfg<NEW>::THREADS;   // shadows the file-scope constant
  GDRN_OPT_IN_SMEM(kfn, P2_SMEM_BYTES);
  const int m_pairs = (plan.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
  const int total = m_pairs * plan.n_tiles;
  if (total <= 0) return GDRN_OK;
  int pairs = gdrn_num_sms() / 2;
  if (pairs > total) pairs = total;
  GDRN_CHECK_CUDA(gdrn_launch_dep(kfn, dim3(2 * pairs), dim3(NUM_THREADS), P2_SMEM_BYTES, stream, plan));
  gdrn_count_launch(1);
  return GDRN_OK;
}

}  // namespace

// epi in {EPI_STORE, EPI_GELU, EPI_RESID}; epi_warps = 8 or 16 (16 only for EPI_GELU)
int gemm_pair_launch(const GemmPlan& plan, int epi_warps, cudaStream_t stream) {
  if (plan.epi == EPI_GELU) return epi_warps == 16 ? launch_pair<EPI_GELU, 16>(plan, stream) : launch_pair<EPI_GELU, 8>(plan, stream);
  if (plan.epi == EPI_RESID) return launch_pair<EPI_RESID, 8>(plan, stream);
  if (plan.epi == EPI_STORE) return launch_pair<EPI_STORE, 8>(plan, stream);
  gdrn_set_last_error(__FILE__, __LINE__, "gemm_pair: unsupported epilogue");
  return GDRN_ERR_INVALID;
}
