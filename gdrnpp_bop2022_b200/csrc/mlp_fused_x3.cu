// Fused ConvNeXt MLP block kernel for the split-bf16 (parity) mode: fc1 -> GELU(erf) -> fc2 -> residual on chip, every
// contraction as three bf16 tensor-core products of hi/lo operand halves (see gemm_pair_x3.cu for the arithmetic).
// Reference op: timm ConvNeXtBlock.mlp (Linear C->4C, GELU, Linear 4C->C) * gamma + shortcut (timm 0.6.7, un-vendored),
// called from models/GDRN_double_mask.py:102 through the backbone.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "gemm_epilogue.cuh"

namespace {

// ================================================================================================================
// Stage 0 (C = 128, 262 144 rows at B = 64): unfused, fc1 writes and fc2 re-reads a 537 MB split hidden activation per
// block (fc1 HBM-write / epilogue bound, fc2 HBM-read bound: 149 + 137 us for 206 GFLOP executed).  Here the hidden
// activation never leaves the SM.  One persistent CTA per SM; per 128-row tile the 512 hidden columns are processed
// in eight rounds of 64, software-pipelined over a flat sequence of global rounds G:
//   MMA-1(G)  H[G&1][128x64]  = A_lo.W1hi + A_hi.W1lo + A_hi.W1hi  (rows r*64.. of W1)   TMEM columns [(G&1)*64, +64)
//   E1(G)     8 warps: tcgen05.ld -> +b1 -> gelu_erf2 -> hi/lo bf16 -> A'hi[G&1], A'lo[G&1] in shared memory, written
//             directly in the 128B-swizzled K-major UMMA operand layout (one 64-wide k-chunk = one 128-byte row)
//   MMA-2(G)  O[ob][128x128] += A'lo.W2hi + A'hi.W2lo + A'hi.W2hi  (hidden k-chunk r of W2)   TMEM columns [128+ob*128, +128)
//   E2        after round 7: gamma*(O+b2) -> staging -> TMA reduce-add into x (epilogue_tile_tma), one tile late
// Shared memory: A {hi,lo} 64 KB | A' 2 x {hi,lo} 64 KB | weight ring 2 x 32 KB | E2 staging 32 KB  (= 224 KB).
// Per round the tensor core has 1536 clk of work, the weight ring streams 64 KB (at the ~42 B/clk/SM L2 fabric limit) and
// E1 runs ~1.5 k clk: the three are balanced by construction; MMA-1(G+1) is issued before the wait for A'(G).
// ================================================================================================================
struct MlpX3Plan {
  GemmPlan g;            // fields used by E2: tmap_out, M, N (= C), bias (= b2), gamma, resid/out (= x), ldo, resid_reduce
  CUtensorMap tmap_a;    // A   [M, 2C]    bf16 (hi | lo), box {64, 128}
  CUtensorMap tmap_w1;   // W1  [4C, 2C]   bf16 (hi | lo), box {64, 64}
  CUtensorMap tmap_w2;   // W2  [C, 2*4C]  bf16 (hi | lo), box {64, 128}
  const float* b1;       // [4C]
  int m_tiles;
};

constexpr int X3F_C = 128;
constexpr int X3F_HC = 64;                                   // hidden columns per round
constexpr int X3F_ROUNDS = 4 * X3F_C / X3F_HC;               // 8
constexpr int X3F_KC1 = X3F_C / BLOCK_K;                     // k-chunks of fc1 (2)
constexpr int X3F_A_BYTES = 2 * X3F_KC1 * A_STAGE_BYTES;     // hi + lo: 64 KB
constexpr int X3F_AP_BYTES = 2 * A_STAGE_BYTES;              // one A' buffer: hi + lo of 128 rows x 64 hidden columns: 32 KB
constexpr int X3F_SLOT_BYTES = 32768;                        // W1 round: {hi kc0, hi kc1, lo kc0, lo kc1} x 8 KB; W2 round: {hi, lo} x 16 KB
constexpr int X3F_SLOTS = 2;
constexpr int X3F_W1_CHUNK = X3F_HC * BLOCK_K * 2;           // 64 rows x 64 k: 8 KB
constexpr int X3F_O_COL = 2 * X3F_HC;                        // 128
constexpr int X3F_SMEM = X3F_A_BYTES + 2 * X3F_AP_BYTES + X3F_SLOTS * X3F_SLOT_BYTES + NUM_EPI_WARPS * EPI_STAGE_BYTES + 256 + 1024;
static_assert(X3F_SMEM <= 227 * 1024, "fused x3 MLP: shared memory budget");

__global__ void __launch_bounds__(NUM_THREADS, 1) mlp_fused_x3_kernel(const __grid_constant__ MlpX3Plan fp) {
  constexpr int C = X3F_C, ROUNDS = X3F_ROUNDS, KC1 = X3F_KC1;
  const GemmPlan& p = fp.g;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  // layout: [A hi kc0,kc1 | A lo kc0,kc1][A' buf0 {hi,lo} | buf1 {hi,lo}][ring][E2 staging 8 x 4 KB][barriers]
  const uint32_t a_smem = smem_base;
  const uint32_t ap_smem = a_smem + X3F_A_BYTES;
  const uint32_t ring_smem = ap_smem + 2 * X3F_AP_BYTES;
  const uint32_t stg_off = X3F_A_BYTES + 2 * X3F_AP_BYTES + X3F_SLOTS * X3F_SLOT_BYTES;
  const uint32_t bar_base = smem_base + stg_off + NUM_EPI_WARPS * EPI_STAGE_BYTES;
  enum { B_AFULL = 0, B_AEMPTY = 1, B_HFULL = 2, B_HEMPTY = 4, B_APFULL = 6, B_APEMPTY = 8, B_OFULL = 10, B_OEMPTY = 12,
         B_RFULL = 14, B_REMPTY = B_RFULL + X3F_SLOTS, B_END = B_REMPTY + X3F_SLOTS };
  auto bar = [&](int i) { return bar_base + 8u * i; };
  const uint32_t tmem_slot = bar_base + 8u * B_END;
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_gen + stg_off + NUM_EPI_WARPS * EPI_STAGE_BYTES + 8 * B_END);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  ptx::griddep_launch();
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&fp.tmap_a); ptx::prefetch_tmap(&fp.tmap_w1); ptx::prefetch_tmap(&fp.tmap_w2);
    ptx::prefetch_tmap(&p.tmap_out);
  }
  if (warp == 1 && lane == 0) {
    ptx::mbar_init(bar(B_AFULL), 1); ptx::mbar_init(bar(B_AEMPTY), 1);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(bar(B_HFULL + i), 1); ptx::mbar_init(bar(B_HEMPTY + i), NUM_EPI_WARPS);
      ptx::mbar_init(bar(B_APFULL + i), NUM_EPI_WARPS); ptx::mbar_init(bar(B_APEMPTY + i), 1);
      ptx::mbar_init(bar(B_OFULL + i), 1); ptx::mbar_init(bar(B_OEMPTY + i), NUM_EPI_WARPS);
    }
    for (int s = 0; s < X3F_SLOTS; ++s) { ptx::mbar_init(bar(B_RFULL + s), 1); ptx::mbar_init(bar(B_REMPTY + s), 1); }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  ptx::griddep_wait();
  const int my_tiles = (int)blockIdx.x < fp.m_tiles ? (fp.m_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const uint32_t total_rounds = (uint32_t)my_tiles * ROUNDS;
  const bool tr = p.trace != nullptr && blockIdx.x == 0;   // GDRN_MLP_TRACE: cycle accounting of CTA 0
  long long tw[6] = {0, 0, 0, 0, 0, 0};
  const long long tr_start = tr ? clock64() : 0;
#define X3F_TIMED_WAIT(slot_, b_, par_) do { const long long t0_ = tr ? clock64() : 0; ptx::mbar_wait((b_), (par_)); \
                                             if (tr) tw[slot_] += clock64() - t0_; } while (0)

  if (warp == 0 && lane == 0) {
    // ================= TMA producer: A(it), then ring slots in the MMA thread's consumption order =================
    int slot = 0;
    uint32_t sphase = 0;
    auto load_w1 = [&](uint32_t G) {       // W1 rows [r*64, +64): hi k-chunks 0,1 then lo k-chunks 0,1
      const int r = (int)(G % ROUNDS);
      ptx::mbar_wait(bar(B_REMPTY + slot), sphase ^ 1);
      ptx::mbar_arrive_expect_tx(bar(B_RFULL + slot), X3F_SLOT_BYTES);
      const uint32_t s0 = ring_smem + slot * X3F_SLOT_BYTES;
      for (int h = 0; h < 2; ++h)
        for (int kc = 0; kc < KC1; ++kc)
          ptx::tma_load_2d(s0 + (h * KC1 + kc) * X3F_W1_CHUNK, &fp.tmap_w1, bar(B_RFULL + slot), h * C + kc * BLOCK_K, r * X3F_HC);
      if (++slot == X3F_SLOTS) { slot = 0; sphase ^= 1; }
    };
    auto load_w2 = [&](uint32_t G) {       // W2 all C rows, hidden k-chunk r: hi then lo
      const int r = (int)(G % ROUNDS);
      ptx::mbar_wait(bar(B_REMPTY + slot), sphase ^ 1);
      ptx::mbar_arrive_expect_tx(bar(B_RFULL + slot), X3F_SLOT_BYTES);
      const uint32_t s0 = ring_smem + slot * X3F_SLOT_BYTES;
      for (int h = 0; h < 2; ++h)
        ptx::tma_load_2d(s0 + h * A_STAGE_BYTES, &fp.tmap_w2, bar(B_RFULL + slot), h * 4 * C + r * BLOCK_K, 0);
      if (++slot == X3F_SLOTS) { slot = 0; sphase ^= 1; }
    };
    auto load_a = [&](uint32_t it) {
      const int tile = (int)blockIdx.x + (int)it * (int)gridDim.x;
      ptx::mbar_wait(bar(B_AEMPTY), (it & 1u) ^ 1u);
      ptx::mbar_arrive_expect_tx(bar(B_AFULL), X3F_A_BYTES);
      for (int h = 0; h < 2; ++h)
        for (int kc = 0; kc < KC1; ++kc)
          ptx::tma_load_2d(a_smem + (h * KC1 + kc) * A_STAGE_BYTES, &fp.tmap_a, bar(B_AFULL), h * C + kc * BLOCK_K, tile * BLOCK_M);
    };
    if (total_rounds > 0) {
      load_a(0);
      load_w1(0);
      for (uint32_t G = 0; G < total_rounds; ++G) {
        if (G + 1 < total_rounds) {
          if ((G + 1) % ROUNDS == 0) load_a((G + 1) / ROUNDS);
          load_w1(G + 1);
        }
        load_w2(G);
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc1 = ptx::make_idesc_bf16(BLOCK_M, X3F_HC);
    constexpr uint32_t idesc2 = ptx::make_idesc_bf16(BLOCK_M, C);
    int slot = 0;
    uint32_t sphase = 0;
    auto mma1 = [&](uint32_t G) {
      const uint32_t hb = G & 1u, n = G >> 1;
      const int r = (int)(G % ROUNDS);
      if (r == 0) X3F_TIMED_WAIT(0, bar(B_AFULL), (G / ROUNDS) & 1u);
      X3F_TIMED_WAIT(1, bar(B_HEMPTY + hb), (n & 1u) ^ 1u);
      X3F_TIMED_WAIT(2, bar(B_RFULL + slot), sphase);
      ptx::tc_fence_after();
      const uint32_t d = tmem_base + hb * X3F_HC;
      const uint32_t w0 = ring_smem + slot * X3F_SLOT_BYTES;
      // per k-chunk, small terms first: A_lo.W1hi, A_hi.W1lo, A_hi.W1hi (the accumulation order of gemm_pair_x3_kernel, so
      // that the fused and the unfused paths produce the same bits)
#pragma unroll
      for (int kc = 0; kc < KC1; ++kc) {
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          const int ah = (t == 0) ? 1 : 0, wh = (t == 1) ? 1 : 0;     // 1 = lo half
          const uint64_t adesc = ptx::make_sw128_kmajor_desc(a_smem + (ah * KC1 + kc) * A_STAGE_BYTES);
          const uint64_t bdesc = ptx::make_sw128_kmajor_desc(w0 + (wh * KC1 + kc) * X3F_W1_CHUNK);
#pragma unroll
          for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk)
            ptx::tc_mma_bf16(d, adesc + 2u * kk, bdesc + 2u * kk, idesc1, (t | kc | kk) != 0 ? 1u : 0u);
        }
      }
      ptx::tc_commit(bar(B_REMPTY + slot));
      if (++slot == X3F_SLOTS) { slot = 0; sphase ^= 1; }
      if (r == ROUNDS - 1) ptx::tc_commit(bar(B_AEMPTY));   // A tile consumed
      ptx::tc_commit(bar(B_HFULL + hb));
    };
    auto mma2 = [&](uint32_t G) {
      const uint32_t hb = G & 1u, n = G >> 1;
      const int r = (int)(G % ROUNDS);
      const uint32_t it = G / ROUNDS, ob = it & 1u;
      X3F_TIMED_WAIT(3, bar(B_APFULL + hb), n & 1u);
      if (r == 0) X3F_TIMED_WAIT(4, bar(B_OEMPTY + ob), ((it >> 1) & 1u) ^ 1u);
      X3F_TIMED_WAIT(2, bar(B_RFULL + slot), sphase);
      ptx::tc_fence_after();
      const uint32_t d = tmem_base + X3F_O_COL + ob * C;
      const uint32_t ap0 = ap_smem + hb * X3F_AP_BYTES;
      const uint32_t w0 = ring_smem + slot * X3F_SLOT_BYTES;
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        const int ah = (t == 0) ? 1 : 0, wh = (t == 1) ? 1 : 0;
        const uint64_t adesc = ptx::make_sw128_kmajor_desc(ap0 + ah * A_STAGE_BYTES);
        const uint64_t bdesc = ptx::make_sw128_kmajor_desc(w0 + wh * A_STAGE_BYTES);
#pragma unroll
        for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk)
          ptx::tc_mma_bf16(d, adesc + 2u * kk, bdesc + 2u * kk, idesc2, (r | t | kk) != 0 ? 1u : 0u);
      }
      ptx::tc_commit(bar(B_REMPTY + slot));
      if (++slot == X3F_SLOTS) { slot = 0; sphase ^= 1; }
      ptx::tc_commit(bar(B_APEMPTY + hb));
      if (r == ROUNDS - 1) ptx::tc_commit(bar(B_OFULL + ob));
    };
    if (total_rounds > 0) {
      mma1(0);
      for (uint32_t G = 0; G < total_rounds; ++G) {
        if (G + 1 < total_rounds) mma1(G + 1);
        mma2(G);
      }
    }
    if (tr) { for (int i = 0; i < 5; ++i) p.trace[i] = tw[i]; p.trace[5] = clock64() - tr_start; p.trace[6] = my_tiles; }
  } else if (warp >= 4) {
    // ================= epilogue warps: E1 per round, E2 per tile =================
    const int ew = warp - 4;
    const int q = ew & 3, half = ew >> 2;      // TMEM lane quarter / 32-column half of the round
    const int row = q * 32 + lane;
    const int sw = row & 7;
    uint8_t* ap_gen = smem_gen + X3F_A_BYTES;
    uint32_t G = 0;
    // E2 (x += gamma * (O + b2)) of a tile is issued one tile LATE, as two 128-byte-column groups interleaved with the E1
    // rounds of the next tile (see mlp_fused.cu): O is long complete by then and the staging buffer has drained.
    auto e2_group = [&](int jt, int grp) {
      if (jt < 0) return;
      const uint32_t job = (uint32_t)jt & 1u;
      const long long te1 = tr ? clock64() : 0;
      if (grp == 0) {
        ptx::mbar_wait(bar(B_OFULL + job), ((uint32_t)jt >> 1) & 1u);
        ptx::tc_fence_after();
      }
      const int jtile = (int)blockIdx.x + jt * (int)gridDim.x;
      epilogue_tile_tma<C, EPI_RESID, true, NUM_EPI_WARPS>(p, jtile, 0, tmem_base + X3F_O_COL + job * C, ew, lane,
                                                           smem_gen + stg_off + ew * EPI_STAGE_BYTES, grp * 32, grp * 32 + 32);
      if (grp == 1) {
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(bar(B_OEMPTY + job));
      }
      if (tr) tw[4] += clock64() - te1;
    };
    for (int it = 0; it < my_tiles; ++it) {
      for (int r = 0; r < ROUNDS; ++r, ++G) {
        const uint32_t hb = G & 1u, n = G >> 1;
        X3F_TIMED_WAIT(0, bar(B_HFULL + hb), n & 1u);
        X3F_TIMED_WAIT(1, bar(B_APEMPTY + hb), (n & 1u) ^ 1u);   // MMA-2 that last read this A' buffer has retired
        ptx::tc_fence_after();
        const long long te0 = tr ? clock64() : 0;
        const uint32_t tmem_row = tmem_base + hb * X3F_HC + ((uint32_t)(q * 32) << 16) + half * 32;
        uint4* dst_hi = reinterpret_cast<uint4*>(ap_gen + hb * X3F_AP_BYTES + row * 128);
        uint4* dst_lo = reinterpret_cast<uint4*>(ap_gen + hb * X3F_AP_BYTES + A_STAGE_BYTES + row * 128);
        const float* b1 = fp.b1 + r * X3F_HC + half * 32;
        float v[32];
        tmem_load_chunk<32>(tmem_row, v);
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          const float4 ba = __ldg(reinterpret_cast<const float4*>(b1 + j));
          const float4 bb = __ldg(reinterpret_cast<const float4*>(b1 + j + 4));
          const float bj[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const f32x2_t g = gelu_erf2(f2_add(f2_pack(v[j + 2 * e], v[j + 2 * e + 1]), f2_pack(bj[2 * e], bj[2 * e + 1])));
            const float2 gf = f2_unpack(g);
            const uint32_t hb16 = pack_bf16(gf.x, gf.y);
            hi[e] = hb16;
            const float2 d = f2_unpack(f2_sub(g, f2_pack(__uint_as_float(hb16 << 16), __uint_as_float(hb16 & 0xffff0000u))));
            lo[e] = pack_bf16(d.x, d.y);
          }
          const int piece = (half * 4 + (j >> 3)) ^ sw;      // 16-byte piece of the 128-byte row, 128B swizzle
          dst_hi[piece] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          dst_lo[piece] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        ptx::tc_fence_before();
        ptx::fence_proxy_async();   // A' was written through the generic proxy, tcgen05.mma reads it through the async proxy
        __syncwarp();
        if (lane == 0) { ptx::mbar_arrive(bar(B_HEMPTY + hb)); ptx::mbar_arrive(bar(B_APFULL + hb)); }
        if (tr) tw[2] += clock64() - te0;
        if (r == 0) e2_group(it - 1, 0);   // four rounds apart: each 32 KB batch of reduce-adds drains before the next
        if (r == 4) e2_group(it - 1, 1);
      }
    }
    e2_group(my_tiles - 1, 0);
    e2_group(my_tiles - 1, 1);
    if (lane == 0) ptx::bulk_wait0();
    if (tr && ew == 0 && lane == 0) for (int i = 0; i < 5; ++i) p.trace[8 + i] = tw[i];
  }
#undef X3F_TIMED_WAIT
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

int mlp_fused_x3_supported(int C, long long M) { return C == X3F_C && M % 128 == 0 && M >= 128 * 148; }

// A [M, 2C] (hi | lo), W1 [4C, 2C] (hi | lo), W2 [C, 8C] (hi | lo) bf16; x [M, C] fp32 updated in place
int mlp_fused_x3_launch(const void* A, const void* W1, const float* b1, const void* W2, const float* b2, const float* gamma,
                        float* x, long long M, int C, cudaStream_t stream) {
  GDRN_REQUIRE(mlp_fused_x3_supported(C, M), "mlp_fused_x3: unsupported shape");
  MlpX3Plan fp;
  memset(&fp, 0, sizeof(fp));
  {
    const uint64_t d[2] = {(uint64_t)2 * C, (uint64_t)M}; const uint64_t st[1] = {(uint64_t)2 * C * 2}; const uint32_t bx[2] = {64, 128};
    int rc = make_tmap_bf16(&fp.tmap_a, A, 2, d, st, bx);
    if (rc != GDRN_OK) return rc;
  }
  {
    const uint64_t d[2] = {(uint64_t)2 * C, (uint64_t)4 * C}; const uint64_t st[1] = {(uint64_t)2 * C * 2}; const uint32_t bx[2] = {64, (uint32_t)X3F_HC};
    int rc = make_tmap_bf16(&fp.tmap_w1, W1, 2, d, st, bx);
    if (rc != GDRN_OK) return rc;
  }
  {
    const uint64_t d[2] = {(uint64_t)8 * C, (uint64_t)C}; const uint64_t st[1] = {(uint64_t)8 * C * 2}; const uint32_t bx[2] = {64, (uint32_t)C};
    int rc = make_tmap_bf16(&fp.tmap_w2, W2, 2, d, st, bx);
    if (rc != GDRN_OK) return rc;
  }
  {
    const uint64_t d[2] = {(uint64_t)C, (uint64_t)M}; const uint64_t st[1] = {(uint64_t)C * 4}; const uint32_t bx[2] = {32, 32};
    int rc = make_tmap_store(&fp.g.tmap_out, x, 1, d, st, bx);
    if (rc != GDRN_OK) return rc;
  }
  fp.g.a_rank = 2; fp.g.M = (int)M; fp.g.N = C; fp.g.epi = EPI_RESID; fp.g.out_f32 = 1; fp.g.out = x; fp.g.resid = x;
  fp.g.ldo = C; fp.g.bias = b2; fp.g.gamma = gamma; fp.g.use_tma_store = 1; fp.g.resid_reduce = 1;
  fp.b1 = b1;
  fp.m_tiles = (int)(M / 128);
  auto kfn = mlp_fused_x3_kernel;
  GDRN_OPT_IN_SMEM(kfn, X3F_SMEM);
  const int grid = fp.m_tiles < gdrn_num_sms() ? fp.m_tiles : gdrn_num_sms();
  static int trace_on = -1;
  if (trace_on < 0) trace_on = getenv("GDRN_MLP_TRACE") ? 1 : 0;
  static long long* d_trace = nullptr;
  if (trace_on) {
    if (!d_trace) GDRN_CHECK_CUDA(cudaMalloc(&d_trace, 16 * sizeof(long long)));
    GDRN_CHECK_CUDA(cudaMemsetAsync(d_trace, 0, 16 * sizeof(long long), stream));
    fp.g.trace = d_trace;
  }
  GDRN_CHECK_CUDA(gdrn_launch_dep(kfn, dim3(grid), dim3(NUM_THREADS), X3F_SMEM, stream, fp));
  gdrn_count_launch(1);
  if (trace_on) {
    long long h[16];
    GDRN_CHECK_CUDA(cudaMemcpyAsync(h, d_trace, sizeof(h), cudaMemcpyDeviceToHost, stream));
    GDRN_CHECK_CUDA(cudaStreamSynchronize(stream));
    fprintf(stderr, "[mlp fused x3 trace] cta0 cycles=%lld tiles=%lld | mma waits: A=%lld H-empty=%lld ring=%lld A'-full=%lld O-empty=%lld | "
                    "epi0: H-full wait=%lld A'-empty wait=%lld E1 busy=%lld O-full wait=%lld E2 busy=%lld\n",
            h[5], h[6], h[0], h[1], h[2], h[3], h[4], h[8], h[9], h[10], h[11], h[12]);
  }
  return GDRN_OK;
}
