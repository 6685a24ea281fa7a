// CTA-pair (cta_group::2) split-bf16 ("bf16x3") implicit-GEMM kernel: the tensor-core path of the fp32-parity precision
// mode (BASELINE.json north_star: R within 1e-4 rad / t within 1e-3 of the reference's fp32 forward,
// core/gdrn_modeling/models/GDRN_double_mask.py:96-160 run with AMP_TEST=False, configs/_base_/common_base.py:219).
//
//   D[m, n] = sum_tap sum_k  A_lo*W_hi + A_hi*W_lo + A_hi*W_hi          (bf16 x bf16 -> fp32 in TMEM)
//
// where every operand v is the pair hi = bf16(v), lo = bf16(v - hi).  The general kernel (gemm_tc.cu) runs the three
// products as a 3x longer tap list, i.e. it streams SIX operand tiles from L2 per (tap, k-chunk).  B200's L2 -> SM
// fabric sustains ~6.3 KB/clk chip-wide = ~42 B/clk per SM, and a cta_group::1 128x256x64 stage needs 96 B/clk at full
// tensor rate: the operand stream, not the tensor pipe, bounds that design.  Here one shared-memory stage holds the FOUR
// distinct tiles {A_hi, A_lo, W_hi, W_lo} of a (tap, k-chunk) and feeds all three products (12 UMMAs), and the CTA pair
// splits the weight rows, so a CTA pulls 64 KB per 1536 tensor-clk = 41.7 B/clk -- at the fabric's rate, not above it.
//
//   cluster of 2 CTAs -> one 256 x BLOCK_N output tile; each CTA: its 128 rows of A (hi + lo) and its BLOCK_N/2 rows of
//   W (hi + lo) per stage; leader's elected thread issues tcgen05.mma.cta_group::2 (M = 256); accumulators in each CTA's
//   TMEM, two accumulator stages; 8 epilogue warps per CTA.
// A operand: rank 2 (plain rows), rank 4 (NHWC pixel boxes, conv taps = shifted boxes, zero padding = TMA OOB fill) or
// rank 5 (stride-2 view), exactly like gemm_tc.cu.  Epilogues: GELU -> split bf16 (TMA stores), RESID (fp32 TMA
// reduce-add into the residual stream), STORE fp32, GNSTATS (fp32 raw conv output + GroupNorm statistics).
// Barrier protocol = gemm_pair.cu (full on the leader, empty / tfull multicast to both CTAs, tempty on the leader).
#include <stdlib.h>
#include "gemm_epilogue.cuh"

namespace {

// NEWARPS = epilogue warps per CTA: 8 (3-stage ring at BLOCK_N 256) or 16 (2-stage ring; four warps per scheduler for the
// GELU epilogue of the short-K fc1 GEMMs, which is issue bound at ~0.5 IPC with two warps per scheduler)
template <int BLOCK_N, int NEWARPS>
struct X3Cfg {
  static constexpr int B_BYTES = (BLOCK_N / 2) * BLOCK_K * 2;             // this CTA's half of the W rows, hi OR lo
  static constexpr int STAGE_BYTES = 2 * A_STAGE_BYTES + 2 * B_BYTES;     // 64 KB (BLOCK_N 256) / 48 KB (128)
  static constexpr int NEW = NEWARPS;
  static constexpr int STAGES = BLOCK_N == 256 ? (NEWARPS == 16 ? 2 : 3) : (NEWARPS == 16 ? 3 : 4);
  static constexpr int THREADS = 128 + 32 * NEW;
  static constexpr int TMEM_COLS = 2 * BLOCK_N;                           // two accumulator stages (512 / 256)
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + NEW * EPI_STAGE_BYTES + 256 + 1024;
};

// Work sequence of one CTA pair.  Default: whole tiles, strided over the pairs.  Balanced k-split (p.streamk, in-place
// EPI_RESID only): the total_tiles * k_iters (tile, k-iteration) units are cut into num_pairs contiguous ranges of equal
// length (+-1), so a launch whose tile count is not a multiple of the pair count (stage-2 fc2: 128 tiles on 74 pairs =
// 0.86 wave efficiency) keeps every pair busy to the end.  A pair's first and last work items may be PART of a tile's K
// range; the parts meet in the L2 reduce-add (x += gamma * partial; the bias goes with the k = 0 part).
struct X3Work { int tile, kb, ke; };
struct X3Sched {
  int streamk, k_iters, total_tiles, num_pairs, tile, u, u_end;
  __device__ X3Sched(int streamk_, int k_iters_, int total_tiles_, int pair_id, int num_pairs_)
      : streamk(streamk_), k_iters(k_iters_), total_tiles(total_tiles_), num_pairs(num_pairs_), tile(pair_id) {
    const long long units = (long long)total_tiles * k_iters;
    u = (int)(units * pair_id / num_pairs);
    u_end = (int)(units * (pair_id + 1) / num_pairs);
  }
  __device__ bool next(X3Work& w) {
    if (!streamk) {
      if (tile >= total_tiles) return false;
      w.tile = tile; w.kb = 0; w.ke = k_iters;
      tile += num_pairs;
      return true;
    }
    if (u >= u_end) return false;
    w.tile = u / k_iters;
    w.kb = u - w.tile * k_iters;
    w.ke = min(k_iters, w.kb + (u_end - u));
    u += w.ke - w.kb;
    return true;
  }
  // k-split: how many parts of this tile are written BEFORE the part [kb, ke) in the fixed order "highest k range first"
  // (= the parts owned by the pairs after this one, up to the owner of the tile's last unit)
  __device__ int parts_before(const X3Work& w, int pair_id) const {
    const long long units = (long long)total_tiles * k_iters;
    const int last = (w.tile + 1) * k_iters - 1;     // the tile's last unit
    int q = pair_id;
    while (q + 1 < num_pairs && (int)(units * (q + 1) / num_pairs) <= last) ++q;
    return q - pair_id;
  }
};

// Bounded acquire spin on a global word (k-split ordering of partial reduce-adds); a protocol bug traps instead of hanging.
__device__ __forceinline__ void sk_wait_flag(const unsigned* f, unsigned want) {
  long long t0 = 0;
  int spins = 0;
  while (true) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
    if (v == want) break;
    if (++spins == 64) t0 = clock64();
    if (spins > 64 && (clock64() - t0) > 4000000000LL) {
      printf("gdrn: k-split flag wait timeout (block %d thread %d want %u have %u)\n", blockIdx.x, threadIdx.x, want, v);
      __trap();
    }
  }
}

template <int BLOCK_N, int EPI, int NEWARPS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(X3Cfg<BLOCK_N, NEWARPS>::THREADS, 1)
gemm_pair_x3_kernel(const __grid_constant__ GemmPlan p) {
  using C = X3Cfg<BLOCK_N, NEWARPS>;
  constexpr int STAGES = C::STAGES;
  constexpr int NEW = C::NEW;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + STAGES * C::STAGE_BYTES + NEW * EPI_STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(
      smem_gen + STAGES * C::STAGE_BYTES + NEW * EPI_STAGE_BYTES + 8 * (2 * STAGES + 4));
  uint8_t* stage_base = smem_gen + STAGES * C::STAGE_BYTES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();   // 0 = leader
  ptx::griddep_launch();
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&p.tmap_a);
    ptx::prefetch_tmap(&p.tmap_b);
    if (p.use_tma_store) ptx::prefetch_tmap(&p.tmap_out);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { ptx::mbar_init(full_bar(s), 1); ptx::mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { ptx::mbar_init(tfull_bar(s), 1); ptx::mbar_init(tempty_bar(s), 2 * NEW); }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc_pair(tmem_slot, C::TMEM_COLS);
    ptx::tmem_relinquish_pair();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();   // both CTAs' barriers and TMEM exist before anything crosses the pair
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  ptx::griddep_wait();       // the prologue above ran under the previous kernel's tail; its results are needed from here on

  const int m_pairs = (p.m_tiles + 1) >> 1;
  const int total_tiles = m_pairs * p.n_tiles;
  const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int k_iters = p.num_taps * p.k_chunks;
  const bool tr = (p.trace != nullptr) && blockIdx.x == 0;   // GDRN_GEMM_TRACE: cycle accounting of the leader of pair 0
  long long tr_acc0 = 0, tr_acc1 = 0;
  const long long tr_start = tr ? clock64() : 0;

  if (warp == 0 && lane == 0) {
    // ================= TMA producer (both CTAs): {A hi, A lo, W hi, W lo} per (tap, k-chunk) =================
    int stage = 0;
    uint32_t phase = 0;
    X3Sched sched(p.streamk, k_iters, total_tiles, pair_id, num_pairs);
    X3Work wk;
    while (sched.next(wk)) {
      const int tile = wk.tile;
      const int n_tile = tile % p.n_tiles, m_pair = tile / p.n_tiles;
      const int m_tile = m_pair * 2 + (int)rank;
      int x0 = 0, y0 = 0, b0 = 0;
      if (p.a_rank != 2) {
        const int tx = m_tile % p.tiles_x;
        const int t2 = m_tile / p.tiles_x;
        x0 = tx << p.lg_bw;
        y0 = (t2 % p.tiles_y) << p.lg_bh;
        b0 = (t2 / p.tiles_y) << p.lg_bb;     // an odd tile count leaves the last peer tile out of bounds: zero fill
      }
      const int brow = n_tile * BLOCK_N + (int)rank * (BLOCK_N / 2);
      int tap = wk.kb / p.k_chunks, kc = wk.kb - tap * p.k_chunks;
      GemmTap tp = p.taps[tap];
      for (int k = wk.kb; k < wk.ke; ++k) {   // k = tap * k_chunks + kc
        { const long long t0 = tr ? clock64() : 0; ptx::mbar_wait(empty_bar(stage), phase ^ 1); if (tr) tr_acc0 += clock64() - t0; }
        const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
        const uint32_t sb = sa + 2 * A_STAGE_BYTES;
        const uint32_t lead_full = ptx::mapa_shared(full_bar(stage), 0);
        if (rank == 0) ptx::mbar_arrive_expect_tx(full_bar(stage), 2 * C::STAGE_BYTES);
        const int k0 = kc * BLOCK_K;
#pragma unroll
        for (int h = 0; h < 2; ++h) {      // h = 0: hi halves, 1: lo halves
          const int ca = k0 + tp.c0 + h * p.x3_a_lo;
          if (p.a_rank == 2) ptx::tma_load_2d_pair(sa + h * A_STAGE_BYTES, &p.tmap_a, lead_full, ca, m_tile * BLOCK_M + tp.d1);
          else if (p.a_rank == 4) ptx::tma_load_4d_pair(sa + h * A_STAGE_BYTES, &p.tmap_a, lead_full, ca, x0 + tp.d1, y0 + tp.d2, b0);
          else ptx::tma_load_5d_pair(sa + h * A_STAGE_BYTES, &p.tmap_a, lead_full, ca, x0 + tp.d1, tp.d2, y0 + tp.d3, b0);
          ptx::tma_load_2d_pair(sb + h * C::B_BYTES, &p.tmap_b, lead_full, tp.b_off + k0 + h * p.x3_b_lo, brow);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
        if (++kc == p.k_chunks) { kc = 0; if (++tap < p.num_taps) tp = p.taps[tap]; }
      }
    }
    if (tr) { p.trace[0] = tr_acc0; p.trace[6] = clock64() - tr_start; p.trace[7] = (total_tiles - 1 - pair_id) / num_pairs + 1; }
  } else if (warp == 1 && lane == 0 && rank == 0) {
    // ================= MMA issuer (leader CTA only): three products per stage =================
    constexpr uint32_t idesc = ptx::make_idesc_bf16(2 * BLOCK_M, BLOCK_N);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    X3Sched sched(p.streamk, k_iters, total_tiles, pair_id, num_pairs);
    X3Work wk;
    for (; sched.next(wk); ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      { const long long t0 = tr ? clock64() : 0; ptx::mbar_wait(tempty_bar(as), aphase ^ 1); if (tr) tr_acc1 += clock64() - t0; }
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * BLOCK_N;
      for (int k = 0; k < wk.ke - wk.kb; ++k) {   // k counts from the start of this part: the first MMA overwrites
        { const long long t0 = tr ? clock64() : 0; ptx::mbar_wait(full_bar(stage), phase); if (tr) tr_acc0 += clock64() - t0; }
        ptx::tc_fence_after();
        const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
        const uint64_t a_hi = ptx::make_sw128_kmajor_desc(sa);
        const uint64_t a_lo = ptx::make_sw128_kmajor_desc(sa + A_STAGE_BYTES);
        const uint64_t b_hi = ptx::make_sw128_kmajor_desc(sa + 2 * A_STAGE_BYTES);
        const uint64_t b_lo = ptx::make_sw128_kmajor_desc(sa + 2 * A_STAGE_BYTES + C::B_BYTES);
        if (p.x3_collect) {
          // per 16-wide k-step: A_lo.W_hi, then A_hi.W_lo keeping A_hi in the collector, then A_hi.W_hi re-using it -- one
          // shared-memory read of A less per k-step (the short-K shapes run at ~98 % of the 128 B/clk shared-memory port)
#pragma unroll
          for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk) {
            ptx::tc_mma_bf16_pair(d_tmem, a_lo + 2u * kk, b_hi + 2u * kk, idesc, (k | kk) != 0 ? 1u : 0u);
            ptx::tc_mma_bf16_pair_a_keep(d_tmem, a_hi + 2u * kk, b_lo + 2u * kk, idesc, 1u);
            ptx::tc_mma_bf16_pair_a_reuse(d_tmem, a_hi + 2u * kk, b_hi + 2u * kk, idesc, 1u);
          }
        } else {
          // small terms first within the stage
#pragma unroll
          for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk)
            ptx::tc_mma_bf16_pair(d_tmem, a_lo + 2u * kk, b_hi + 2u * kk, idesc, (k | kk) != 0 ? 1u : 0u);
#pragma unroll
          for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk) ptx::tc_mma_bf16_pair(d_tmem, a_hi + 2u * kk, b_lo + 2u * kk, idesc, 1u);
#pragma unroll
          for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk) ptx::tc_mma_bf16_pair(d_tmem, a_hi + 2u * kk, b_hi + 2u * kk, idesc, 1u);
        }
        ptx::tc_commit_pair(empty_bar(stage), 0x3);   // slot reusable in BOTH CTAs
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      ptx::tc_commit_pair(tfull_bar(as), 0x3);         // accumulators complete in both CTAs
    }
    if (tr) { p.trace[1] = tr_acc0; p.trace[2] = tr_acc1; }
  } else if (warp >= 4) {
    // ================= epilogue (both CTAs) =================
    const int ew = warp - 4;
    uint8_t* stg = stage_base + ew * EPI_STAGE_BYTES;
    int it = 0;
    X3Sched sched(p.streamk, k_iters, total_tiles, pair_id, num_pairs);
    X3Work wk;
    for (; sched.next(wk); ++it) {
      const int tile = wk.tile;
      const int n_tile = tile % p.n_tiles, m_pair = tile / p.n_tiles;
      const int m_tile = m_pair * 2 + (int)rank;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      long long t0 = tr ? clock64() : 0;
      ptx::mbar_wait(tfull_bar(as), aphase);
      if (tr) { const long long t1 = clock64(); tr_acc0 += t1 - t0; t0 = t1; }
      ptx::tc_fence_after();
      const uint32_t acc = tmem_base + as * BLOCK_N;
      if (m_tile < p.m_tiles) {   // the peer of an odd tile count has nothing to write
        if constexpr (EPI == EPI_GELU) {
          epilogue_tile_tma_split<BLOCK_N, EPI_GELU, NEW>(p, m_tile, n_tile, acc, ew, lane, stg);
        } else if constexpr (EPI == EPI_GNSTATS) {
          epilogue_tile_staged_t<BLOCK_N, EPI_GNSTATS, true>(p, m_tile, n_tile, acc, ew, lane, stg);
        } else if constexpr (EPI == EPI_RESID) {
          // k-split: a PART of the tile's K range adds gamma * partial (bias with the k = 0 part).  The parts of one tile
          // are reduce-added in a fixed order (highest k range first; flag word per (tile, CTA, warp) = parts written so
          // far), so x does not depend on which pair happened to finish first.
          const bool part = p.streamk && (wk.kb != 0 || wk.ke != k_iters);
          unsigned* flag = nullptr;
          if (part) {
            flag = p.sk_flags + ((size_t)tile * 2 + rank) * NEW + ew;
            const int before = sched.parts_before(wk, pair_id);
            if (before > 0) {
              if (lane == 0) sk_wait_flag(flag, (unsigned)before);
              __syncwarp();
            }
          }
          epilogue_tile_tma<BLOCK_N, EPI, true, NEW>(p, m_tile, n_tile, acc, ew, lane, stg, 0, 1 << 30, wk.kb == 0);
          if (part && lane == 0) {
            if (wk.kb == 0) {
              *reinterpret_cast<volatile unsigned*>(flag) = 0u;   // last part of the tile: leave the word clean for the next launch
            } else {
              ptx::bulk_wait0();   // this part's reduce-adds have been performed (not only read from the staging tile)
              asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(flag) : "memory");
            }
          }
        } else {   // EPI_STORE, fp32 out through TMA stores
          epilogue_tile_tma<BLOCK_N, EPI, true, NEW>(p, m_tile, n_tile, acc, ew, lane, stg);
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa_shared(tempty_bar(as), 0));
      if (tr) tr_acc1 += clock64() - t0;
    }
    if (lane == 0) ptx::bulk_wait0();
    if (tr && ew == 0 && lane == 0) { p.trace[3] = tr_acc0; p.trace[4] = tr_acc1; }
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();   // nobody frees TMEM or exits while the peer can still touch this CTA
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_pair(tmem_base, C::TMEM_COLS);
  }
}

template <int BLOCK_N, int EPI, int NEWARPS = 8>
int launch_x3(const GemmPlan& plan, cudaStream_t stream) {
  using C = X3Cfg<BLOCK_N, NEWARPS>;
  auto kfn = gemm_pair_x3_kernel<BLOCK_N, EPI, NEWARPS>;
  GDRN_OPT_IN_SMEM(kfn, C::SMEM_BYTES);
  const int m_pairs = (plan.m_tiles + 1) / 2;
  const int total = m_pairs * plan.n_tiles;
  if (total <= 0) return GDRN_OK;
  int pairs = gdrn_num_sms() / 2;
  if (pairs > total && !plan.streamk) pairs = total;   // the k-split schedule feeds every pair even with fewer tiles than pairs
  GDRN_CHECK_CUDA(gdrn_launch_dep(kfn, dim3(2 * pairs), dim3(C::THREADS), C::SMEM_BYTES, stream, plan));
  gdrn_count_launch(1);
  return GDRN_OK;
}

}  // namespace

// Shapes the pair-x3 kernel takes (the caller falls back to the general kernel with an expanded tap list otherwise):
// whole BLOCK_N tiles, an epilogue it implements, and enough tiles to occupy the 74 CTA pairs.
int gemm_pair_x3_supported(const GemmPlan& plan, int block_n) {
  if (!plan.split || plan.x3_expanded || plan.b_ptr == nullptr) return 0;
  if (block_n != 128 && block_n != 256) return 0;
  if (plan.N % block_n != 0) return 0;
  if (plan.a_rank != 2 && plan.a_rank != 4 && plan.a_rank != 5) return 0;
  const bool f32 = plan.epi == EPI_RESID || (plan.epi == EPI_STORE && plan.out_f32) || (plan.epi == EPI_GNSTATS && plan.out_f32);
  if (!(f32 || plan.epi == EPI_GELU)) return 0;
  if (plan.epi != EPI_GNSTATS && plan.a_rank != 2) return 0;   // TMA-store epilogues write plain [M, ldo] rows
  static int min_tiles = -1;   // GDRN_X3_MIN_TILES: pair tiles below which the general kernel (narrower tiles) is used
  if (min_tiles < 0) { const char* e = getenv("GDRN_X3_MIN_TILES"); min_tiles = e ? atoi(e) : 48; }
  const int total = ((plan.m_tiles + 1) / 2) * plan.n_tiles;
  return total >= min_tiles;
}

int gemm_pair_x3_launch(const GemmPlan& plan_in, int block_n, cudaStream_t stream) {
  GemmPlan plan = plan_in;
  {
    static int collect = -1;   // GDRN_X3_COLLECT=1: A_hi through the A collector (measured: no change, 74.1 vs 74.3 us on stage-2 fc1)
    if (collect < 0) { const char* e = getenv("GDRN_X3_COLLECT"); collect = e ? atoi(e) : 0; }
    plan.x3_collect = collect;
  }
  plan.streamk = 0;
  if (plan.epi == EPI_RESID && plan.resid_reduce && plan.sk_flags != nullptr) {
    // balanced k-split for the in-place residual GEMMs whose tile count leaves the last wave partly empty (stage-1/2/3 fc2
    // at B = 64: 256 / 128 / 64 tiles on 74 pairs = 0.86 wave efficiency)
    static int sk = -1;   // GDRN_X3_KSPLIT=0: whole tiles only (A/B experiments)
    if (sk < 0) { const char* e = getenv("GDRN_X3_KSPLIT"); sk = e ? atoi(e) : 0; }
    const int pairs = gdrn_num_sms() / 2;
    const long long total = (long long)((plan.m_tiles + 1) / 2) * plan.n_tiles;
    const long long k_iters = (long long)plan.num_taps * plan.k_chunks;
    const long long waves = (total + pairs - 1) / pairs;
    if (sk && total * 100 < waves * pairs * 93 && total * k_iters >= 8LL * pairs && total * 2 * 8 <= plan.sk_flag_words)
      plan.streamk = 1;
  }
  if (plan.epi == EPI_GELU) {
    // 16 epilogue warps for the short-K GELU GEMMs (stages 0 / 1: the epilogue, not the mainloop, sets their time)
    static int gelu16_max_k = -1;   // GDRN_X3_GELU16_MAX_KITERS: use the 16-warp variant up to this many k-iterations (0 = never)
    if (gelu16_max_k < 0) { const char* e = getenv("GDRN_X3_GELU16_MAX_KITERS"); gelu16_max_k = e ? atoi(e) : 4; }
    if (plan.num_taps * plan.k_chunks <= gelu16_max_k) {
      if (block_n == 256) return launch_x3<256, EPI_GELU, 16>(plan, stream);
      if (block_n == 128) return launch_x3<128, EPI_GELU, 16>(plan, stream);
    }
  }
#define GDRN_X3_CASE(BN, E) if (block_n == BN && plan.epi == E) return launch_x3<BN, E>(plan, stream);
  GDRN_X3_CASE(256, EPI_GELU)
  GDRN_X3_CASE(256, EPI_RESID)
  GDRN_X3_CASE(256, EPI_STORE)
  GDRN_X3_CASE(256, EPI_GNSTATS)
  GDRN_X3_CASE(128, EPI_GELU)
  GDRN_X3_CASE(128, EPI_RESID)
  GDRN_X3_CASE(128, EPI_STORE)
  GDRN_X3_CASE(128, EPI_GNSTATS)
#undef GDRN_X3_CASE
  gdrn_set_last_error(__FILE__, __LINE__, "gemm_pair_x3: unsupported (block_n, epilogue) combination");
  return GDRN_ERR_INVALID;
}
