// tcgen05 + TMA implicit-GEMM kernel for the GDRNPP dense path (sm_100a).
//
// One persistent CTA per SM, warp-specialised:
//   warp 0 lane 0 : TMA producer  (A pixel boxes + W tiles -> 128B-swizzled smem ring)
//   warp 1 lane 0 : tcgen05.mma issuer (UMMA 128 x BLOCK_N x 16, bf16 -> fp32 accumulators in TMEM)
//   warp 2        : TMEM allocate / free
//   warps 4..7    : epilogue (tcgen05.ld -> registers -> fused math -> global), double-buffered
//                   against the next tile's MMAs through two TMEM accumulator stages.
//
// It replaces, for the reference forward (core/gdrn_modeling/models/GDRN_double_mask.py:102-160), every
// cuDNN/cuBLAS call made by timm ConvNeXt (Linear fc1/fc2, stem 4x4s4, 2x2s2 downsample), by the geometry
// head (heads/top_down_doublemask_xyz_region_head.py:177-211: ConvTranspose2d, six 3x3 convs, 1x1 out
// conv) and by ConvPnPNet (heads/conv_pnp_net.py:120-183: three 3x3 s2 convs and four Linear layers).
#include "common.cuh"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "gemm_tc.h"

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KB
constexpr int NUM_THREADS = 384;              // 4 control warps + 8 epilogue warps
constexpr int NUM_EPI_WARPS = 8;
constexpr int EPI_STAGE_PITCH = 80;           // 64-byte row segment + 16 B pad (conflict-free 16-byte accesses)
// per epilogue warp: either the TMA-store staging tile (32 rows x 128 B, 128B-swizzled, 1024-aligned) or, for the
// gather-store path, 32 rows x 80 B + 32 x int64 row map + bias/gamma (2 x 128 f32) = 3840 B
constexpr int EPI_STAGE_BYTES = 4096;
constexpr int SMEM_BUDGET = 227 * 1024 - NUM_EPI_WARPS * EPI_STAGE_BYTES - 1024 - 256;

template <int BLOCK_N>
struct Cfg {
  static constexpr int B_STAGE_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES = (SMEM_BUDGET / STAGE_BYTES) > 8 ? 8 : (SMEM_BUDGET / STAGE_BYTES);
  static constexpr int ACC_COLS = BLOCK_N <= 128 ? 128 : 256;  // TMEM columns per accumulator stage
  static constexpr int TMEM_COLS = 2 * ACC_COLS;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + NUM_EPI_WARPS * EPI_STAGE_BYTES;
};

struct RowInfo {
  long long orow;  // output row index
  int b;           // image index (rank 4/5) or row / rows_per_roi
  bool valid;
};

__device__ __forceinline__ RowInfo map_row(const GemmPlan& p, int m_tile, int r) {
  RowInfo ri;
  if (p.a_rank == 2) {
    long long grow = (long long)m_tile * BLOCK_M + r;
    ri.orow = grow;
    ri.valid = grow < p.M;
    ri.b = p.rows_per_roi > 0 ? (int)(grow / p.rows_per_roi) : 0;
  } else {
    int tx = m_tile % p.tiles_x;
    int t2 = m_tile / p.tiles_x;
    int ty = t2 % p.tiles_y;
    int tb = t2 / p.tiles_y;
    int ix = r & ((1 << p.lg_bw) - 1);
    int iy = (r >> p.lg_bw) & ((1 << p.lg_bh) - 1);
    int ib = r >> (p.lg_bw + p.lg_bh);
    int b = (tb << p.lg_bb) + ib;
    int y = (ty << p.lg_bh) + iy;
    int x = (tx << p.lg_bw) + ix;
    ri.b = b;
    ri.valid = b < p.M;
    ri.orow = ((long long)b * p.OH + (y * p.osy + p.ooy)) * p.OW + (x * p.osx + p.oox);
  }
  return ri;
}

constexpr int PREFETCH_AHEAD = 6;  // k-iterations of A/B requested into L2 ahead of the shared-memory ring


template <int CH>
__device__ __forceinline__ void store_row_chunk(const GemmPlan& p, const RowInfo& ri, int col, const float (&v)[CH],
                                                bool f32) {
  // col is a multiple of CH; columns >= N are dropped
  int nvalid = p.N - col;
  if (nvalid <= 0) return;
  if (f32) {
    float* o = reinterpret_cast<float*>(p.out) + ri.orow * p.ldo + col;
    if (nvalid >= CH) {
#pragma unroll
      for (int j = 0; j < CH; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < CH; ++j)
        if (j < nvalid) o[j] = v[j];
    }
  } else {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + ri.orow * p.ldo + col;
    if (nvalid >= CH) {
#pragma unroll
      for (int j = 0; j < CH; j += 8) {
        uint4 u;
        u.x = pack_bf16(v[j], v[j + 1]);
        u.y = pack_bf16(v[j + 2], v[j + 3]);
        u.z = pack_bf16(v[j + 4], v[j + 5]);
        u.w = pack_bf16(v[j + 6], v[j + 7]);
        *reinterpret_cast<uint4*>(o + j) = u;
      }
    } else {
#pragma unroll
      for (int j = 0; j < CH; ++j)
        if (j < nvalid) o[j] = __float2bfloat16(v[j]);
    }
  }
}

template <int CH>
__device__ __forceinline__ void load_vec(const float* __restrict__ src, int col, int N, float (&v)[CH]) {
  if (src == nullptr) {
#pragma unroll
    for (int j = 0; j < CH; ++j) v[j] = 0.f;
    return;
  }
  if (col + CH <= N) {
#pragma unroll
    for (int j = 0; j < CH; j += 4) {
      float4 t = __ldg(reinterpret_cast<const float4*>(src + col + j));
      v[j] = t.x; v[j + 1] = t.y; v[j + 2] = t.z; v[j + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < CH; ++j) v[j] = (col + j < N) ? __ldg(src + col + j) : 0.f;
  }
}

template <int CH>
__device__ __forceinline__ void tmem_load_chunk(uint32_t taddr, float (&v)[CH]) {
  if constexpr (CH == 32) {
    uint32_t r[32];
    ptx::tmem_ld32(taddr, r);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
  } else {
    uint32_t r[16];
    ptx::tmem_ld16(taddr, r);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
  }
}

// ------------------------------------------------------------------------------------------------
// Epilogues.  Each of the 128 epilogue threads owns one accumulator row (TMEM lane).
// ------------------------------------------------------------------------------------------------
template <int BLOCK_N, int EPI>
__device__ __forceinline__ void epilogue_tile(const GemmPlan& p, int m_tile, int n_tile, uint32_t tmem_row, int lane,
                                              uint8_t* stg) {
  constexpr int CH = BLOCK_N >= 32 ? 32 : 16;
  const int r = ((threadIdx.x >> 5) & 3) * 32 + lane;
  const RowInfo ri = map_row(p, m_tile, r);
  const int n0 = n_tile * BLOCK_N;

  if constexpr (EPI == EPI_STORE || EPI == EPI_GELU || EPI == EPI_RESID) {
#pragma unroll 1
    for (int c = 0; c < BLOCK_N; c += CH) {
      float v[CH];
      tmem_load_chunk<CH>(tmem_row + c, v);
      const int col = n0 + c;
      if (!ri.valid || col >= p.N) continue;
      float bias[CH];
      load_vec<CH>(p.bias, col, p.N, bias);
      if constexpr (EPI == EPI_STORE) {
#pragma unroll
        for (int j = 0; j < CH; ++j) v[j] += bias[j];
        store_row_chunk<CH>(p, ri, col, v, p.out_f32 != 0);
      } else if constexpr (EPI == EPI_GELU) {
#pragma unroll
        for (int j = 0; j < CH; ++j) v[j] = gelu_fast(v[j] + bias[j]);
        store_row_chunk<CH>(p, ri, col, v, false);
      } else {  // EPI_RESID
        float g[CH], x[CH];
        load_vec<CH>(p.gamma, col, p.N, g);
        const float* rs = p.resid + ri.orow * p.ldo + col;
#pragma unroll
        for (int j = 0; j < CH; j += 4) {
          float4 t = *reinterpret_cast<const float4*>(rs + j);
          x[j] = t.x; x[j + 1] = t.y; x[j + 2] = t.z; x[j + 3] = t.w;
        }
#pragma unroll
        for (int j = 0; j < CH; ++j) v[j] = fmaf(g[j], v[j] + bias[j], x[j]);
        store_row_chunk<CH>(p, ri, col, v, true);
      }
    }
  } else if constexpr (EPI == EPI_GNSTATS) {
    const int cpg = p.gn_cpg;  // 4 or 8
#pragma unroll 1
    for (int c = 0; c < BLOCK_N; c += CH) {
      float v[CH];
      tmem_load_chunk<CH>(tmem_row + c, v);
      const int col = n0 + c;
      if (col >= p.N) continue;
      if (ri.valid) store_row_chunk<CH>(p, ri, col, v, p.out_f32 != 0);
      // per-group partial sums over this thread's row (stats use the values as stored)
      float s[8], ss[8];
#pragma unroll
      for (int g = 0; g < 8; ++g) { s[g] = 0.f; ss[g] = 0.f; }
      if (!p.out_f32) {
#pragma unroll
        for (int j = 0; j < CH; ++j) v[j] = __bfloat162float(__float2bfloat16(v[j]));
      }
      if (cpg == 8) {
#pragma unroll
        for (int j = 0; j < CH; ++j) { s[j >> 3] += v[j]; ss[j >> 3] = fmaf(v[j], v[j], ss[j >> 3]); }
      } else {
#pragma unroll
        for (int j = 0; j < CH; ++j) { s[j >> 2] += v[j]; ss[j >> 2] = fmaf(v[j], v[j], ss[j >> 2]); }
      }
      const int ng = CH / cpg;  // groups in this chunk (4 or 8)
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        if (g < ng) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            s[g] += __shfl_xor_sync(0xffffffffu, s[g], o);
            ss[g] += __shfl_xor_sync(0xffffffffu, ss[g], o);
          }
        }
      }
      if (ri.valid) {  // warp-uniform: a warp's 32 rows lie in one image
        double* st = p.gn_stats + ((long long)ri.b * p.gn_groups + col / cpg) * 2;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          if (g < ng && lane == g) {
            atomicAdd(st + 2 * g, (double)s[g]);
            atomicAdd(st + 2 * g + 1, (double)ss[g]);
          }
        }
      }
    }
  } else if constexpr (EPI == EPI_BIAS_LN) {
    // N <= BLOCK_N (N % 32 == 0): the thread sees the whole channel vector of its pixel. Three TMEM passes.
    float sum = 0.f;
#pragma unroll 1
    for (int c = 0; c < p.N; c += CH) {
      float v[CH], bias[CH];
      tmem_load_chunk<CH>(tmem_row + c, v);
      load_vec<CH>(p.bias, c, p.N, bias);
#pragma unroll
      for (int j = 0; j < CH; ++j) sum += v[j] + bias[j];
    }
    const float mean = sum / (float)p.N;
    float sq = 0.f;
#pragma unroll 1
    for (int c = 0; c < p.N; c += CH) {
      float v[CH], bias[CH];
      tmem_load_chunk<CH>(tmem_row + c, v);
      load_vec<CH>(p.bias, c, p.N, bias);
#pragma unroll
      for (int j = 0; j < CH; ++j) { float d = v[j] + bias[j] - mean; sq = fmaf(d, d, sq); }
    }
    const float rstd = rsqrtf(sq / (float)p.N + p.ln_eps);
#pragma unroll 1
    for (int c = 0; c < p.N; c += CH) {
      float v[CH], bias[CH], w[CH], bb[CH];
      tmem_load_chunk<CH>(tmem_row + c, v);
      load_vec<CH>(p.bias, c, p.N, bias);
      load_vec<CH>(p.ln_w, c, p.N, w);
      load_vec<CH>(p.ln_b, c, p.N, bb);
#pragma unroll
      for (int j = 0; j < CH; ++j) v[j] = fmaf((v[j] + bias[j] - mean) * rstd, w[j], bb[j]);
      if constexpr (CH == 32) {
        if (p.use_tma_store) {  // 32 rows x 128 B staging tile (SWIZZLE_128B) -> one TMA store, rows >= M clipped
          if (lane == 0) ptx::bulk_wait_read0();
          __syncwarp();
          uint4* dst = reinterpret_cast<uint4*>(stg + lane * 128);
          const int sw = lane & 7;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            dst[j ^ sw] = make_uint4(__float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]), __float_as_uint(v[4 * j + 2]),
                                     __float_as_uint(v[4 * j + 3]));
          ptx::fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            ptx::tma_store_2d(&p.tmap_out, ptx::smem_u32(stg), c, m_tile * BLOCK_M + (int)((threadIdx.x >> 5) & 3) * 32);
            ptx::bulk_commit();
          }
          continue;
        }
      }
      if (ri.valid) store_row_chunk<CH>(p, ri, c, v, true);
    }
  } else if constexpr (EPI == EPI_OUTCONV) {
    // BLOCK_N == 80: [vis, full, x, y, z, region_0..64, pad x10] of the ROI's own class
    // (GDRN_double_mask.py:107-126 gather + :131-148 feature assembly + conv_pnp_net.py:130-136).
    float v[80];
    {
      float t[32];
      tmem_load_chunk<32>(tmem_row, t);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = t[j];
      tmem_load_chunk<32>(tmem_row + 32, t);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[32 + j] = t[j];
      float t16[16];
      tmem_load_chunk<16>(tmem_row + 64, t16);
#pragma unroll
      for (int j = 0; j < 16; ++j) v[64 + j] = t16[j];
    }
    if (ri.valid) {
      const int b = ri.b;
      const int pix = (int)(ri.orow - (long long)b * p.rows_per_roi);
      int cls = (int)p.roi_classes[b];
      cls = cls < 0 ? 0 : (cls >= p.num_classes ? p.num_classes - 1 : cls);
      const float* ob = p.oc_bias + cls * 80;
#pragma unroll
      for (int j = 0; j < 72; j += 4) {
        float4 t = __ldg(reinterpret_cast<const float4*>(ob + j));
        v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
      }
      const long long hw = p.rows_per_roi;
      if (p.map_mask) {
        p.map_mask[(long long)b * hw + pix] = v[0];
        p.map_full[(long long)b * hw + pix] = v[1];
        p.map_x[(long long)b * hw + pix] = v[2];
        p.map_y[(long long)b * hw + pix] = v[3];
        p.map_z[(long long)b * hw + pix] = v[4];
#pragma unroll
        for (int j = 0; j < 65; ++j) p.map_region[((long long)b * 65 + j) * hw + pix] = v[5 + j];
      }
      // softmax over region[1:65] = v[6..69]
      float mx = v[6];
#pragma unroll
      for (int j = 7; j < 70; ++j) mx = fmaxf(mx, v[j]);
      float den = 0.f;
#pragma unroll
      if (p.split) {  // precise mode: full-accuracy exp
#pragma unroll
        for (int j = 6; j < 70; ++j) { v[j] = expf(v[j] - mx); den += v[j]; }
      } else {
#pragma unroll
        for (int j = 6; j < 70; ++j) { v[j] = __expf(v[j] - mx); den += v[j]; }
      }
      const float inv = 1.0f / den;
      const float ex = __ldg(p.roi_extents + b * 3 + 0), ey = __ldg(p.roi_extents + b * 3 + 1),
                  ez = __ldg(p.roi_extents + b * 3 + 2);
      float f[72];
      f[0] = (v[2] - 0.5f) * ex;
      f[1] = (v[3] - 0.5f) * ey;
      f[2] = (v[4] - 0.5f) * ez;
      f[3] = __ldg(p.roi_coord_2d + ((long long)b * 2 + 0) * hw + pix);
      f[4] = __ldg(p.roi_coord_2d + ((long long)b * 2 + 1) * hw + pix);
#pragma unroll
      for (int j = 0; j < 64; ++j) f[5 + j] = v[6 + j] * inv;
      f[69] = 0.f; f[70] = 0.f; f[71] = 0.f;
      const int pw = p.split ? 256 : 128;  // Patch-PnP input row width (split mode: [hi 128 | lo 128])
      uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.pnp_in) + ri.orow * pw);
#pragma unroll
      for (int j = 0; j < 72; j += 8) {
        uint4 u;
        u.x = pack_bf16(f[j], f[j + 1]);
        u.y = pack_bf16(f[j + 2], f[j + 3]);
        u.z = pack_bf16(f[j + 4], f[j + 5]);
        u.w = pack_bf16(f[j + 6], f[j + 7]);
        dst[j >> 3] = u;
      }
#pragma unroll
      for (int j = 9; j < 16; ++j) dst[j] = make_uint4(0, 0, 0, 0);
      if (p.split) {
#pragma unroll
        for (int j = 0; j < 72; ++j) f[j] = f[j] - __bfloat162float(__float2bfloat16(f[j]));
#pragma unroll
        for (int j = 0; j < 72; j += 8) {
          uint4 u;
          u.x = pack_bf16(f[j], f[j + 1]);
          u.y = pack_bf16(f[j + 2], f[j + 3]);
          u.z = pack_bf16(f[j + 4], f[j + 5]);
          u.w = pack_bf16(f[j + 6], f[j + 7]);
          dst[16 + (j >> 3)] = u;
        }
#pragma unroll
        for (int j = 25; j < 32; ++j) dst[j] = make_uint4(0, 0, 0, 0);
      }
    }
  }
}


// GELU mode 2: fp32 tanh form with the hardware tanh.approx.f32 (1 MUFU / element):
// 0.5*x*(1 + tanh(x*(c0 + c1*x^2 + c2*x^4))) with (c0,c1,c2) fitted to the erf form (tools/fit_gelu.py).
__device__ __forceinline__ float gelu_tanh_f32(float x) {
  const float xc = fminf(fmaxf(x, -8.f), 8.f);  // the fitted polynomial changes sign beyond |x| ~ 11
  const float x2 = xc * xc;
  float p = fmaf(x2, GELU_T2, GELU_T1);
  p = fmaf(p, x2, GELU_T0);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(xc * p));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}

// Epilogue for EPI_STORE / EPI_GELU / EPI_RESID / EPI_GNSTATS with coalesced global traffic.
// Eight warps: warp ew owns TMEM lanes [32*(ew&3), +32) and the column half (ew>>2) of the tile.  A thread owns
// one accumulator row; 64-byte row segments (32 bf16 or 16 fp32 columns) are staged in the warp's private
// shared-memory buffer (pitch 80 B, conflict-free 16-byte accesses) and then written (and, for the residual,
// first read) with 16-byte accesses in which 4 lanes cover one row segment: every warp-wide access touches
// eight full 64-byte runs (16 whole sectors) instead of 32 scattered 16-byte pieces.
template <int BLOCK_N, int EPI, bool F32>
__device__ __forceinline__ void epilogue_tile_staged_t(const GemmPlan& p, int m_tile, int n_tile, uint32_t tmem_acc,
                                                       int ew, int lane, uint8_t* stg) {
  static_assert(BLOCK_N >= 64, "staged epilogue needs BLOCK_N >= 64");
  constexpr int CPW = BLOCK_N / 2;      // columns per warp
  constexpr int CH = F32 ? 16 : 32;     // columns per staged row segment (64 bytes)
  constexpr int ESZ = F32 ? 4 : 2;
  const int q = ew & 3, half = ew >> 2;
  const int r = q * 32 + lane;
  const RowInfo ri = map_row(p, m_tile, r);
  long long* s_orow = reinterpret_cast<long long*>(stg + 32 * EPI_STAGE_PITCH);
  const float* s_bias = reinterpret_cast<const float*>(stg + 32 * EPI_STAGE_PITCH + 256);  // [128] bias, [128] gamma
  s_orow[lane] = ri.valid ? ri.orow : -1;
  __syncwarp();
  const int n0 = n_tile * BLOCK_N + half * CPW;
  const uint32_t tmem_row = tmem_acc + ((uint32_t)(q * 32) << 16) + half * CPW;
  uint8_t* my_row = stg + lane * EPI_STAGE_PITCH;
  const int cpg = p.gn_cpg;
  const int fl_row = lane >> 2, fl_piece = lane & 3;  // flush / prefetch mapping: 8 rows x 4 pieces per pass
  const bool trc = (p.trace != nullptr) && blockIdx.x == 0 && ew == 0;
  long long tq_tmem = 0, tq_comp = 0, tq_flush = 0;
  uint4 rres[4];
  if constexpr (EPI == EPI_RESID) {
    if (n0 < p.N) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const long long orow = s_orow[k * 8 + fl_row];
        rres[k] = (orow >= 0) ? *reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(p.resid + orow * p.ldo + n0) + fl_piece * 16)
                              : make_uint4(0, 0, 0, 0);
      }
    }
  }

#pragma unroll 1
  for (int c = 0; c < CPW; c += CH) {
    const int col = n0 + c;
    if (col >= p.N) break;  // warp-uniform
    if constexpr (EPI == EPI_RESID) {
      // residual segment rows: registers (fetched one chunk ahead, see below) -> staging buffer
#pragma unroll
      for (int k = 0; k < 4; ++k) *reinterpret_cast<uint4*>(stg + (k * 8 + fl_row) * EPI_STAGE_PITCH + fl_piece * 16) = rres[k];
      // issue the coalesced loads of the NEXT chunk now; they complete while this chunk is processed
      if (c + CH < CPW && col + CH < p.N) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const long long orow = s_orow[k * 8 + fl_row];
          rres[k] = (orow >= 0) ? *reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(p.resid + orow * p.ldo + col + CH) + fl_piece * 16)
                                : make_uint4(0, 0, 0, 0);
        }
      }
      __syncwarp();
    }
    float v[CH];
    long long tq0 = trc ? clock64() : 0;
    tmem_load_chunk<CH>(tmem_row + c, v);
    if (trc) { const long long t = clock64(); tq_tmem += t - tq0; tq0 = t; }
    if constexpr (EPI == EPI_STORE || EPI == EPI_GELU) {
      const float4* sb = reinterpret_cast<const float4*>(s_bias + c);  // warp-wide broadcast reads
#pragma unroll
      for (int j = 0; j < CH; j += 4) {
        const float4 b4 = sb[j >> 2];
        v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
      }
    } else if constexpr (EPI == EPI_RESID) {
      float bias[CH], g[CH];
      {
        const float4* sb = reinterpret_cast<const float4*>(s_bias + c);
        const float4* sg = reinterpret_cast<const float4*>(s_bias + 128 + c);
#pragma unroll
        for (int j = 0; j < CH; j += 4) {
          const float4 b4 = sb[j >> 2], g4 = sg[j >> 2];
          bias[j] = b4.x; bias[j + 1] = b4.y; bias[j + 2] = b4.z; bias[j + 3] = b4.w;
          g[j] = g4.x; g[j + 1] = g4.y; g[j + 2] = g4.z; g[j + 3] = g4.w;
        }
      }
      const float4* xr = reinterpret_cast<const float4*>(my_row);
#pragma unroll
      for (int j = 0; j < CH; j += 4) {
        const float4 x = xr[j >> 2];
        v[j] = fmaf(g[j], v[j] + bias[j], x.x);
        v[j + 1] = fmaf(g[j + 1], v[j + 1] + bias[j + 1], x.y);
        v[j + 2] = fmaf(g[j + 2], v[j + 2] + bias[j + 2], x.z);
        v[j + 3] = fmaf(g[j + 3], v[j + 3] + bias[j + 3], x.w);
      }
    }
    // ---- registers -> staging row ----
    float lo[F32 ? 1 : CH];  // split mode: residual of the bf16 rounding, written in a second round
    if constexpr (F32) {
      float4* dst = reinterpret_cast<float4*>(my_row);
#pragma unroll
      for (int j = 0; j < CH; j += 4) dst[j >> 2] = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
      uint4* dst = reinterpret_cast<uint4*>(my_row);
      if (EPI == EPI_GELU && p.gelu_mode == 1) {
#pragma unroll
        for (int j = 0; j < CH; j += 8) {
          uint4 w;
          w.x = gelu_pack2_f16(v[j], v[j + 1]); w.y = gelu_pack2_f16(v[j + 2], v[j + 3]);
          w.z = gelu_pack2_f16(v[j + 4], v[j + 5]); w.w = gelu_pack2_f16(v[j + 6], v[j + 7]);
          dst[j >> 3] = w;
        }
      } else {
        if constexpr (EPI == EPI_GELU) {
          if (p.gelu_mode == 3) {
#pragma unroll
            for (int j = 0; j < CH; ++j) v[j] = gelu_erf(v[j]);
          } else if (p.gelu_mode == 2) {
#pragma unroll
            for (int j = 0; j < CH; ++j) v[j] = gelu_tanh_f32(v[j]);
          } else {
#pragma unroll
            for (int j = 0; j < CH; ++j) v[j] = gelu_fast(v[j]);
          }
        }
        if (p.split) {  // hi must be exactly the cvt.rn value the lo halves are computed against
#pragma unroll
          for (int j = 0; j < CH; j += 8) {
            uint4 w;
            w.x = pack_bf16(v[j], v[j + 1]); w.y = pack_bf16(v[j + 2], v[j + 3]);
            w.z = pack_bf16(v[j + 4], v[j + 5]); w.w = pack_bf16(v[j + 6], v[j + 7]);
            dst[j >> 3] = w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < CH; j += 8) {
            uint4 w;
            w.x = pack_bf16(v[j], v[j + 1]); w.y = pack_bf16(v[j + 2], v[j + 3]);
            w.z = pack_bf16(v[j + 4], v[j + 5]); w.w = pack_bf16(v[j + 6], v[j + 7]);
            dst[j >> 3] = w;
          }
        }
        if (p.split) {
#pragma unroll
          for (int j = 0; j < CH; ++j) lo[j] = v[j] - __bfloat162float(__float2bfloat16(v[j]));
        }
      }
    }
    if constexpr (EPI == EPI_GNSTATS) {
      // per-(image, group) sum / sum of squares of the values as stored
      constexpr int NG_MAX = CH / 4;
      float s[NG_MAX], ss[NG_MAX];
#pragma unroll
      for (int g = 0; g < NG_MAX; ++g) { s[g] = 0.f; ss[g] = 0.f; }
      if (!F32) {
#pragma unroll
        for (int j = 0; j < CH; ++j) v[j] = __bfloat162float(__float2bfloat16(v[j]));
      }
      if (cpg == 8) {
#pragma unroll
        for (int j = 0; j < CH; ++j) { s[j >> 3] += v[j]; ss[j >> 3] = fmaf(v[j], v[j], ss[j >> 3]); }
      } else {
#pragma unroll
        for (int j = 0; j < CH; ++j) { s[j >> 2] += v[j]; ss[j >> 2] = fmaf(v[j], v[j], ss[j >> 2]); }
      }
      const int ng = CH / cpg;
#pragma unroll
      for (int g = 0; g < NG_MAX; ++g) {
        if (g < ng) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            s[g] += __shfl_xor_sync(0xffffffffu, s[g], o);
            ss[g] += __shfl_xor_sync(0xffffffffu, ss[g], o);
          }
        }
      }
      if (ri.valid) {  // warp-uniform: a warp's 32 rows lie in one image
        double* st = p.gn_stats + ((long long)ri.b * p.gn_groups + col / cpg) * 2;
#pragma unroll
        for (int g = 0; g < NG_MAX; ++g) {
          if (g < ng && lane == g) {
            atomicAdd(st + 2 * g, (double)s[g]);
            atomicAdd(st + 2 * g + 1, (double)ss[g]);
          }
        }
      }
    }
    __syncwarp();
    if (trc) { const long long t = clock64(); tq_comp += t - tq0; tq0 = t; }
    // ---- coalesced flush of the 32 row segments ----
#pragma unroll
    for (int pass = 0; pass < 32; pass += 8) {
      const int rr = pass + fl_row;
      const long long orow = s_orow[rr];
      if (orow >= 0)
        *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(p.out) + (orow * p.ldo + col) * ESZ + fl_piece * 16) =
            *reinterpret_cast<const uint4*>(stg + rr * EPI_STAGE_PITCH + fl_piece * 16);
    }
    __syncwarp();
    if (trc) { const long long t = clock64(); tq_flush += t - tq0; tq0 = t; }
    if constexpr (!F32) {
      if (p.split) {  // second round: the lo halves go to columns [N + col, ...)
        uint4* dst = reinterpret_cast<uint4*>(my_row);
#pragma unroll
        for (int j = 0; j < CH; j += 8) {
          uint4 w;
          w.x = pack_bf16(lo[j], lo[j + 1]); w.y = pack_bf16(lo[j + 2], lo[j + 3]);
          w.z = pack_bf16(lo[j + 4], lo[j + 5]); w.w = pack_bf16(lo[j + 6], lo[j + 7]);
          dst[j >> 3] = w;
        }
        __syncwarp();
#pragma unroll
        for (int pass = 0; pass < 32; pass += 8) {
          const int rr = pass + fl_row;
          const long long orow = s_orow[rr];
          if (orow >= 0)
            *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(p.out) + (orow * p.ldo + p.N + col) * ESZ + fl_piece * 16) =
                *reinterpret_cast<const uint4*>(stg + rr * EPI_STAGE_PITCH + fl_piece * 16);
        }
        __syncwarp();
      }
    }
  }
  if (trc && lane == 0) { p.trace[8] += tq_tmem; p.trace[9] += tq_comp; p.trace[10] += tq_flush; }
}


// Epilogue for rank-2 outputs (EPI_STORE / EPI_GELU / EPI_RESID) with TMA stores.
// Eight warps as above.  A thread owns one accumulator row; it writes 128-byte row segments (64 bf16 or 32 fp32
// columns, produced as two 64-byte halves = two tcgen05.ld) into the warp's 32 x 128 B staging tile in the
// SWIZZLE_128B layout (16-byte piece j of row r at piece j ^ (r & 7): conflict-free row-wise STS.128), and one lane
// issues a single cp.async.bulk.tensor store per tile.  No address arithmetic, LDS or STG in the flush; rows beyond
// M are clipped by the tensor map.  The residual of EPI_RESID is read by its owner thread (64 contiguous bytes per
// half, register-prefetched one half ahead).
template <int BLOCK_N, int EPI, bool F32, int NEW = NUM_EPI_WARPS>
__device__ __forceinline__ void epilogue_tile_tma(const GemmPlan& p, int m_tile, int n_tile, uint32_t tmem_acc, int ew,
                                                  int lane, uint8_t* stg, int g_begin = 0, int g_end = 1 << 30) {
  // [g_begin, g_end): sub-range of this warp's columns (multiples of GW) -- lets a caller interleave the 128-byte groups
  // of a tile with other work so that the wait for the previous TMA store never blocks (fused MLP kernel)
  static_assert(BLOCK_N >= 128, "TMA-store epilogue needs BLOCK_N >= 128");
  constexpr int CPW = BLOCK_N / (NEW / 4);  // columns per warp (NEW / 4 warps share one TMEM lane quarter)
  static_assert(CPW * (F32 ? 4 : 2) >= 128, "a warp must own at least one 128-byte row segment");
  constexpr int CH = F32 ? 16 : 32;     // columns per 64-byte half row (one tcgen05.ld)
  constexpr int GW = 2 * CH;            // columns per staged 128-byte row (one TMA store)
  const int q = ew & 3, half = ew >> 2;
  const long long grow = (long long)m_tile * BLOCK_M + q * 32 + lane;
  const bool rvalid = grow < p.M;
  const int n0 = n_tile * BLOCK_N + half * CPW;
  const uint32_t tmem_row = tmem_acc + ((uint32_t)(q * 32) << 16) + half * CPW;
  const uint32_t stg_u32 = ptx::smem_u32(stg);
  uint8_t* my_row = stg + lane * 128;
  const int sw = lane & 7;
  const bool trc = (p.trace != nullptr) && blockIdx.x == 0 && ew == 0;
  long long tq_tmem = 0, tq_comp = 0, tq_flush = 0;

  uint4 rres[4];
  const float* rrow = nullptr;
  const bool red = (EPI == EPI_RESID) && p.resid_reduce;  // launch-uniform: out += gamma*(acc+bias) by TMA reduce-add
  if constexpr (EPI == EPI_RESID) {
    rrow = p.resid + grow * p.ldo + n0;
    if (rvalid && n0 + g_begin < p.N && !red) {
#pragma unroll
      for (int k = 0; k < 4; ++k) rres[k] = *reinterpret_cast<const uint4*>(rrow + g_begin + 4 * k);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) rres[k] = make_uint4(0, 0, 0, 0);
    }
  }

#pragma unroll 1
  for (int g = g_begin; g < CPW && g < g_end; g += GW) {
    const int gcol = n0 + g;
    if (gcol >= p.N) break;  // warp-uniform
    long long tq0 = trc ? clock64() : 0;
    // the previous store of this warp must have finished reading the staging tile
    if (lane == 0) ptx::bulk_wait_read0();
    __syncwarp();
    if (trc) { const long long t = clock64(); tq_flush += t - tq0; tq0 = t; }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = g + h * CH;
      const int col = n0 + c;
      // bias is read four columns at a time right where it is consumed (keeps the live register set small
      // enough for the 16-warp epilogue variant); the loads are warp-uniform L1 hits
      auto bias4 = [&](int j) {
        return p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + col + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
      };
      float x[EPI == EPI_RESID ? CH : 1], gm[EPI == EPI_RESID ? CH : 1];
      if constexpr (EPI == EPI_RESID) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          x[4 * k] = __uint_as_float(rres[k].x); x[4 * k + 1] = __uint_as_float(rres[k].y);
          x[4 * k + 2] = __uint_as_float(rres[k].z); x[4 * k + 3] = __uint_as_float(rres[k].w);
          const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.gamma + col + 4 * k));
          gm[4 * k] = g4.x; gm[4 * k + 1] = g4.y; gm[4 * k + 2] = g4.z; gm[4 * k + 3] = g4.w;
        }
        // next half's residual: in flight while this half is processed
        if (rvalid && !red && c + CH < CPW && col + CH < p.N) {
#pragma unroll
          for (int k = 0; k < 4; ++k) rres[k] = *reinterpret_cast<const uint4*>(rrow + c + CH + 4 * k);
        }
      }
      float v[CH];
      long long tq1 = trc ? clock64() : 0;
      tmem_load_chunk<CH>(tmem_row + c, v);
      if (trc) { const long long t = clock64(); tq_tmem += t - tq1; }
      uint4* dst = reinterpret_cast<uint4*>(my_row);
      if constexpr (F32) {
        if constexpr (EPI == EPI_RESID) {
#pragma unroll
          for (int j = 0; j < CH; j += 4) {
            const float4 b4 = bias4(j);
            v[j] = fmaf(gm[j], v[j] + b4.x, x[j]); v[j + 1] = fmaf(gm[j + 1], v[j + 1] + b4.y, x[j + 1]);
            v[j + 2] = fmaf(gm[j + 2], v[j + 2] + b4.z, x[j + 2]); v[j + 3] = fmaf(gm[j + 3], v[j + 3] + b4.w, x[j + 3]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < CH; j += 4) {
            const float4 b4 = bias4(j);
            v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
          }
        }
#pragma unroll
        for (int j = 0; j < CH; j += 4)
          dst[(h * 4 + (j >> 2)) ^ sw] = make_uint4(__float_as_uint(v[j]), __float_as_uint(v[j + 1]),
                                                    __float_as_uint(v[j + 2]), __float_as_uint(v[j + 3]));
      } else {
        if (EPI == EPI_GELU && p.gelu_mode == 1) {
#pragma unroll
          for (int j = 0; j < CH; j += 8) {
            const float4 ba = bias4(j), bb = bias4(j + 4);
            v[j] += ba.x; v[j + 1] += ba.y; v[j + 2] += ba.z; v[j + 3] += ba.w;
            v[j + 4] += bb.x; v[j + 5] += bb.y; v[j + 6] += bb.z; v[j + 7] += bb.w;
            uint4 w;
            w.x = gelu_pack2_f16(v[j], v[j + 1]); w.y = gelu_pack2_f16(v[j + 2], v[j + 3]);
            w.z = gelu_pack2_f16(v[j + 4], v[j + 5]); w.w = gelu_pack2_f16(v[j + 6], v[j + 7]);
            dst[(h * 4 + (j >> 3)) ^ sw] = w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < CH; j += 4) {
            const float4 b4 = bias4(j);
            v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
          }
          if constexpr (EPI == EPI_GELU) {
            if (p.gelu_mode == 3) {
#pragma unroll
              for (int j = 0; j < CH; ++j) v[j] = gelu_erf(v[j]);
            } else if (p.gelu_mode == 2) {
#pragma unroll
              for (int j = 0; j < CH; ++j) v[j] = gelu_tanh_f32(v[j]);
            } else {
#pragma unroll
              for (int j = 0; j < CH; ++j) v[j] = gelu_fast(v[j]);
            }
          }
#pragma unroll
          for (int j = 0; j < CH; j += 8) {
            uint4 w;
            w.x = pack_bf16(v[j], v[j + 1]); w.y = pack_bf16(v[j + 2], v[j + 3]);
            w.z = pack_bf16(v[j + 4], v[j + 5]); w.w = pack_bf16(v[j + 6], v[j + 7]);
            dst[(h * 4 + (j >> 3)) ^ sw] = w;
          }
        }
      }
    }
    if (trc) { const long long t = clock64(); tq_comp += t - tq0; tq0 = t; }
    ptx::fence_proxy_async();  // generic-proxy smem writes -> visible to the TMA (async proxy)
    __syncwarp();
    if (lane == 0) {
      if (red) ptx::tma_reduce_add_2d(&p.tmap_out, stg_u32, gcol, m_tile * BLOCK_M + q * 32);
      else ptx::tma_store_2d(&p.tmap_out, stg_u32, gcol, m_tile * BLOCK_M + q * 32);
      ptx::bulk_commit();
    }
    if (trc) { const long long t = clock64(); tq_flush += t - tq0; }
  }
  if (trc && lane == 0) { p.trace[8] += tq_tmem; p.trace[9] += tq_comp; p.trace[10] += tq_flush; }
}

template <int BLOCK_N, int EPI>
__device__ __forceinline__ void epilogue_tile_staged(const GemmPlan& p, int m_tile, int n_tile, uint32_t tmem_acc,
                                                     int ew, int lane, uint8_t* stg) {
  if constexpr (BLOCK_N >= 128 && (EPI == EPI_STORE || EPI == EPI_GELU || EPI == EPI_RESID)) {
    if (p.use_tma_store) {  // launch-uniform
      if constexpr (EPI == EPI_RESID) {
        epilogue_tile_tma<BLOCK_N, EPI, true>(p, m_tile, n_tile, tmem_acc, ew, lane, stg);
      } else if constexpr (EPI == EPI_GELU) {
        epilogue_tile_tma<BLOCK_N, EPI, false>(p, m_tile, n_tile, tmem_acc, ew, lane, stg);
      } else {
        if (p.out_f32) epilogue_tile_tma<BLOCK_N, EPI, true>(p, m_tile, n_tile, tmem_acc, ew, lane, stg);
        else epilogue_tile_tma<BLOCK_N, EPI, false>(p, m_tile, n_tile, tmem_acc, ew, lane, stg);
      }
      return;
    }
  }
  if constexpr (EPI == EPI_RESID) {
    epilogue_tile_staged_t<BLOCK_N, EPI, true>(p, m_tile, n_tile, tmem_acc, ew, lane, stg);
  } else if constexpr (EPI == EPI_GELU) {
    epilogue_tile_staged_t<BLOCK_N, EPI, false>(p, m_tile, n_tile, tmem_acc, ew, lane, stg);
  } else {
    if (p.out_f32) epilogue_tile_staged_t<BLOCK_N, EPI, true>(p, m_tile, n_tile, tmem_acc, ew, lane, stg);
    else epilogue_tile_staged_t<BLOCK_N, EPI, false>(p, m_tile, n_tile, tmem_acc, ew, lane, stg);
  }
}

template <int BLOCK_N, int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1) gemm_tc_kernel(const __grid_constant__ GemmPlan p) {
  using C = Cfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment required by SWIZZLE_128B
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + C::STAGES * C::STAGE_BYTES + NUM_EPI_WARPS * EPI_STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * C::STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * C::STAGES + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * C::STAGES + 4);
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(
      smem_gen + C::STAGES * C::STAGE_BYTES + NUM_EPI_WARPS * EPI_STAGE_BYTES + 8 * (2 * C::STAGES + 4));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // epilogues that need the whole row in one thread (LayerNorm, out-conv) or tiny N keep 4 warps + direct stores
  constexpr bool kStaged = (BLOCK_N >= 64) && (EPI == EPI_STORE || EPI == EPI_GELU || EPI == EPI_RESID || EPI == EPI_GNSTATS);
  uint8_t* stage_base = smem_gen + C::STAGES * C::STAGE_BYTES;  // 1024-aligned (TMA-store tiles are 128B-swizzled)

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&p.tmap_a);
    ptx::prefetch_tmap(&p.tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(tfull_bar(s), 1);
      ptx::mbar_init(tempty_bar(s), kStaged ? NUM_EPI_WARPS : 4);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, C::TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int total_tiles = p.m_tiles * p.n_tiles;
  const int k_iters = p.num_taps * p.k_chunks;
  // optional cycle accounting of CTA 0 (GDRN_GEMM_TRACE, api_core.cu): who waits for whom
  const bool tr = (p.trace != nullptr) && blockIdx.x == 0;
  long long tr_acc0 = 0, tr_acc1 = 0;
  const long long tr_start = tr ? clock64() : 0;

  if (warp == 0 && lane == 0) {
    // ================= TMA producer =================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int n_tile = p.n_major ? tile / p.m_tiles : tile % p.n_tiles;
      const int m_tile = p.n_major ? tile % p.m_tiles : tile / p.n_tiles;
      int x0 = 0, y0 = 0, b0 = 0;
      if (p.a_rank != 2) {
        int tx = m_tile % p.tiles_x;
        int t2 = m_tile / p.tiles_x;
        x0 = tx << p.lg_bw;
        y0 = (t2 % p.tiles_y) << p.lg_bh;
        b0 = (t2 / p.tiles_y) << p.lg_bb;
      }
      int brow = n_tile * BLOCK_N;
      if (EPI == EPI_OUTCONV) {
        long long grow = (long long)m_tile * BLOCK_M;
        int b = (int)(grow / p.rows_per_roi);
        int cls = (int)p.roi_classes[b];
        cls = cls < 0 ? 0 : (cls >= p.num_classes ? p.num_classes - 1 : cls);  // never index outside the weights
        brow += cls * p.b_rows_per_class;
      }
      for (int tap = 0; tap < p.num_taps; ++tap) {
        const GemmTap tp = p.taps[tap];
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          { const long long t0 = tr ? clock64() : 0; ptx::mbar_wait(empty_bar(stage), phase ^ 1); if (tr) tr_acc0 += clock64() - t0; }
          const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
          const uint32_t sb = sa + A_STAGE_BYTES;
          ptx::mbar_arrive_expect_tx(full_bar(stage), C::STAGE_BYTES);
          const int k0 = kc * BLOCK_K;
          if (p.a_rank == 2) {
            ptx::tma_load_2d(sa, &p.tmap_a, full_bar(stage), k0 + tp.c0, m_tile * BLOCK_M + tp.d1);
            // A is streamed from HBM (it was written by the previous kernel): request it into L2 several
            // k-iterations ahead so the shared-memory ring only has to cover L2 latency, not DRAM latency.
            int pk = kc + PREFETCH_AHEAD, pm = m_tile;
            if (pk >= p.k_chunks) {
              pk -= p.k_chunks;
              const int nt = tile + gridDim.x;
              const int nm = p.n_major ? nt % p.m_tiles : nt / p.n_tiles;
              pm = (nt < total_tiles && nm != m_tile) ? nm : -1;
            }
            if (pm >= 0 && pk < p.k_chunks && p.num_taps == 1) ptx::tma_prefetch_2d(&p.tmap_a, pk * BLOCK_K, pm * BLOCK_M);
          } else if (p.a_rank == 4) {
            ptx::tma_load_4d(sa, &p.tmap_a, full_bar(stage), k0 + tp.c0, x0 + tp.d1, y0 + tp.d2, b0);
          } else {
            ptx::tma_load_5d(sa, &p.tmap_a, full_bar(stage), k0 + tp.c0, x0 + tp.d1, tp.d2, y0 + tp.d3, b0);
          }
          ptx::tma_load_2d(sb, &p.tmap_b, full_bar(stage), tp.b_off + k0, brow);
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc = ptx::make_idesc_bf16(BLOCK_M, BLOCK_N);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      { const long long t0 = tr ? clock64() : 0; ptx::mbar_wait(tempty_bar(as), aphase ^ 1); if (tr) tr_acc1 += clock64() - t0; }
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * C::ACC_COLS;
      for (int k = 0; k < k_iters; ++k) {
        { const long long t0 = tr ? clock64() : 0; ptx::mbar_wait(full_bar(stage), phase); if (tr) tr_acc0 += clock64() - t0; }
        ptx::tc_fence_after();
        const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
        const uint32_t sb = sa + A_STAGE_BYTES;
        const uint64_t adesc = ptx::make_sw128_kmajor_desc(sa);
        const uint64_t bdesc = ptx::make_sw128_kmajor_desc(sb);
#pragma unroll
        for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk) {
          // advance 16 bf16 = 32 B inside the 128B swizzle row: +2 in the (addr >> 4) field
          ptx::tc_mma_bf16(d_tmem, adesc + 2u * kk, bdesc + 2u * kk, idesc, (k | kk) != 0 ? 1u : 0u);
        }
        ptx::tc_commit(empty_bar(stage));  // smem slot reusable once these MMAs retire
        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
      }
      ptx::tc_commit(tfull_bar(as));  // accumulator complete
    }
  } else if (warp >= 4 && (kStaged || warp < 8)) {
    // ================= epilogue =================
    const int ew = warp - 4;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int n_tile = p.n_major ? tile / p.m_tiles : tile % p.n_tiles;
      const int m_tile = p.n_major ? tile % p.m_tiles : tile / p.n_tiles;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      if (kStaged && EPI != EPI_GNSTATS && !p.use_tma_store) {
        // this warp's bias (and layer-scale gamma) segment -> its shared-memory slot, before the accumulator wait
        float* sbw = reinterpret_cast<float*>(stage_base + ew * EPI_STAGE_BYTES + 32 * EPI_STAGE_PITCH + 256);
        const int cb = n_tile * BLOCK_N + (ew >> 2) * (BLOCK_N / 2) + lane * 4;
        if (lane * 4 < BLOCK_N / 2) {
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f), g4 = b4;
          if (cb + 3 < p.N) {
            if (p.bias) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + cb));
            if (EPI == EPI_RESID) g4 = __ldg(reinterpret_cast<const float4*>(p.gamma + cb));
          }
          *reinterpret_cast<float4*>(sbw + lane * 4) = b4;
          if (EPI == EPI_RESID) *reinterpret_cast<float4*>(sbw + 128 + lane * 4) = g4;
        }
        __syncwarp();
      }
      if (kStaged && EPI == EPI_RESID && !p.resid_reduce) {
        // the residual rows of this tile are not produced by the MMAs: pull them into L2 while the mainloop runs
        const RowInfo pri = map_row(p, m_tile, (ew & 3) * 32 + lane);
        if (pri.valid) {
          const float* rp = p.resid + pri.orow * p.ldo + n_tile * BLOCK_N + (ew >> 2) * (BLOCK_N / 2);
#pragma unroll
          for (int b = 0; b < (BLOCK_N / 2) * 4; b += 128)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const uint8_t*>(rp) + b));
        }
      }
      long long t0 = tr ? clock64() : 0;
      ptx::mbar_wait(tfull_bar(as), aphase);
      if (tr) { const long long t1 = clock64(); tr_acc0 += t1 - t0; t0 = t1; }
      ptx::tc_fence_after();
      if constexpr (kStaged) {
        epilogue_tile_staged<BLOCK_N, EPI>(p, m_tile, n_tile, tmem_base + as * C::ACC_COLS, ew, lane,
                                           stage_base + ew * EPI_STAGE_BYTES);
      } else {
        const uint32_t tmem_row = tmem_base + ((uint32_t)((ew & 3) * 32) << 16) + as * C::ACC_COLS;
        epilogue_tile<BLOCK_N, EPI>(p, m_tile, n_tile, tmem_row, lane, stage_base + (ew & 3) * EPI_STAGE_BYTES);
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(tempty_bar(as));
      if (tr) tr_acc1 += clock64() - t0;
    }
    if (tr && ew == 0 && lane == 0) { p.trace[3] = tr_acc0; p.trace[4] = tr_acc1; }
    if (lane == 0) ptx::bulk_wait0();  // outstanding TMA stores (no-op when none were issued)
  }
  if (tr && lane == 0) {
    if (warp == 0) { p.trace[0] = tr_acc0; p.trace[6] = clock64() - tr_start; p.trace[7] = (total_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1; }
    if (warp == 1) { p.trace[1] = tr_acc0; p.trace[2] = tr_acc1; }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

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
  static bool configured = false;
  auto kfn = gemm_pair_kernel<EPI, NEW>;
  constexpr int P2_SMEM_BYTES = P2Cfg<NEW>::SMEM_BYTES;
  constexpr int NUM_THREADS = P2Cfg<NEW>::THREADS;   // shadows the file-scope constant
  if (!configured) {
    GDRN_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, P2_SMEM_BYTES));
    configured = true;
  }
  const int m_pairs = (plan.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
  const int total = m_pairs * plan.n_tiles;
  if (total <= 0) return GDRN_OK;
  int pairs = gdrn_num_sms() / 2;
  if (pairs > total) pairs = total;
  kfn<<<2 * pairs, NUM_THREADS, P2_SMEM_BYTES, stream>>>(plan);
  GDRN_CHECK_CUDA(cudaGetLastError());
  gdrn_count_launch(1);
  return GDRN_OK;
}

// ================================================================================================================
// Fused ConvNeXt MLP for the wide-and-shallow stage (C = 128):  x += gamma * (W2 . gelu(W1 . a + b1) + b2)
// in ONE kernel per block; the 4C-wide hidden activation never leaves the SM.  Unfused, stage 0 writes and re-reads a
// 268 MB bf16 `Hb` per block (fc1 HBM-write bound, fc2 HBM-read bound: 190 us for 69 GFLOP).
// One persistent CTA per SM; per 128-row tile the hidden dimension is processed in four rounds of 128 columns, software
// pipelined over a flat sequence of global rounds G (tile boundaries included):
//   MMA-1(G)  H[G&1][128x128] = A[128xC] . W1[r]^T                   TMEM columns [ (G&1)*128, +128 )
//   E1(G)     8 warps: tcgen05.ld -> +b1 -> GELU -> bf16 -> A'[G&1] in shared memory, written directly in the
//             128B-swizzled K-major UMMA operand layout (16-byte piece j of row r at j ^ (r & 7))
//   MMA-2(G)  O[ob][128xC]   += A'[G&1][128x128] . W2[:, r]^T        TMEM columns [ 256 + ob*128, +128 )
//   E2        after round 3: gamma*(O+b2) -> staging -> TMA reduce-add into x  (epilogue_tile_tma)
// The MMA thread issues MMA-1(G+1) BEFORE it waits for A'(G): the tensor core, the weight stream (ring of 32 KB slots
// in consumption order W1(0), {W1(G+1), W2(G)}...) and the A-tile load of the next row block all run underneath E1,
// which is the critical path (GELU at ~0.5 IPC on 8 warps).  H, A' and O are double-buffered.
// ================================================================================================================
struct MlpFusedPlan {
  GemmPlan g;            // fields used by E2: tmap_out, M, N (= C), bias (= b2), gamma, resid/out (= x), ldo, resid_reduce
  CUtensorMap tmap_a;    // A   [M, C]   bf16, box {64, 128}
  CUtensorMap tmap_w1;   // W1  [4C, C]  bf16, box {64, 128}
  CUtensorMap tmap_w2;   // W2  [C, 4C]  bf16, box {64, 128}
  const float* b1;       // [4C]
  int m_tiles;
};

constexpr int MF_SLOT_BYTES = 32768;
constexpr int MF_SLOTS = 3;

template <int C>
__global__ void __launch_bounds__(NUM_THREADS, 1) mlp_fused_kernel(const __grid_constant__ MlpFusedPlan fp) {
  static_assert(C == 128, "fused MLP: only C = 128 is instantiated");
  constexpr int KC1 = C / 64;                 // k-chunks of fc1 (2)
  constexpr int ROUNDS = 4 * C / 128;         // hidden rounds of 128 columns (4)
  constexpr int A_BYTES = KC1 * A_STAGE_BYTES;            // 32 KB
  constexpr int AP_BYTES = 2 * A_STAGE_BYTES;             // one A' buffer: 128 rows x 128 hidden columns
  constexpr int CHUNK_BYTES = 128 * BLOCK_K * 2;          // a 128-row x 64-k weight chunk (16 KB); two per slot
  constexpr int O_COL = 256;
  const GemmPlan& p = fp.g;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  // layout: [A][A' x2][ring][E2 staging 8 x 4 KB][barriers]
  const uint32_t a_smem = smem_base;
  const uint32_t ap_smem = a_smem + A_BYTES;
  const uint32_t ring_smem = ap_smem + 2 * AP_BYTES;
  const uint32_t stg_off = A_BYTES + 2 * AP_BYTES + MF_SLOTS * MF_SLOT_BYTES;
  const uint32_t bar_base = smem_base + stg_off + NUM_EPI_WARPS * EPI_STAGE_BYTES;
  enum { B_AFULL = 0, B_AEMPTY = 1, B_HFULL = 2, B_HEMPTY = 4, B_APFULL = 6, B_APEMPTY = 8, B_OFULL = 10, B_OEMPTY = 12,
         B_RFULL = 14, B_REMPTY = B_RFULL + MF_SLOTS, B_END = B_REMPTY + MF_SLOTS };
  auto bar = [&](int i) { return bar_base + 8u * i; };
  const uint32_t tmem_slot = bar_base + 8u * B_END;
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_gen + stg_off + NUM_EPI_WARPS * EPI_STAGE_BYTES + 8 * B_END);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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
    for (int s = 0; s < MF_SLOTS; ++s) { ptx::mbar_init(bar(B_RFULL + s), 1); ptx::mbar_init(bar(B_REMPTY + s), 1); }
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
  const int my_tiles = (int)blockIdx.x < fp.m_tiles ? (fp.m_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const uint32_t total_rounds = (uint32_t)my_tiles * ROUNDS;
  const bool tr = p.trace != nullptr && blockIdx.x == 0;   // GDRN_MLP_TRACE: cycle accounting of CTA 0
  long long tw[6] = {0, 0, 0, 0, 0, 0};
  const long long tr_start = tr ? clock64() : 0;
#define MF_TIMED_WAIT(slot_, b_, par_) do { const long long t0_ = tr ? clock64() : 0; ptx::mbar_wait((b_), (par_)); \
                                            if (tr) tw[slot_] += clock64() - t0_; } while (0)

  if (warp == 0 && lane == 0) {
    // ================= TMA producer: A(it), then slots in the MMA thread's consumption order =================
    int slot = 0;
    uint32_t sphase = 0;
    auto load_w1 = [&](uint32_t G) {       // W1 rows [r*128, +128), both k-chunks -> one slot
      const int r = (int)(G % ROUNDS);
      ptx::mbar_wait(bar(B_REMPTY + slot), sphase ^ 1);
      ptx::mbar_arrive_expect_tx(bar(B_RFULL + slot), MF_SLOT_BYTES);
      for (int kc = 0; kc < KC1; ++kc)
        ptx::tma_load_2d(ring_smem + slot * MF_SLOT_BYTES + kc * CHUNK_BYTES, &fp.tmap_w1, bar(B_RFULL + slot), kc * BLOCK_K, r * 128);
      if (++slot == MF_SLOTS) { slot = 0; sphase ^= 1; }
    };
    auto load_w2 = [&](uint32_t G) {       // W2 all C rows, hidden k-chunks 2r, 2r+1 -> one slot
      const int r = (int)(G % ROUNDS);
      ptx::mbar_wait(bar(B_REMPTY + slot), sphase ^ 1);
      ptx::mbar_arrive_expect_tx(bar(B_RFULL + slot), MF_SLOT_BYTES);
      for (int c = 0; c < 2; ++c)
        ptx::tma_load_2d(ring_smem + slot * MF_SLOT_BYTES + c * CHUNK_BYTES, &fp.tmap_w2, bar(B_RFULL + slot), (2 * r + c) * BLOCK_K, 0);
      if (++slot == MF_SLOTS) { slot = 0; sphase ^= 1; }
    };
    auto load_a = [&](uint32_t it) {
      const int tile = (int)blockIdx.x + (int)it * (int)gridDim.x;
      ptx::mbar_wait(bar(B_AEMPTY), (it & 1u) ^ 1u);
      ptx::mbar_arrive_expect_tx(bar(B_AFULL), A_BYTES);
      for (int kc = 0; kc < KC1; ++kc)
        ptx::tma_load_2d(a_smem + kc * A_STAGE_BYTES, &fp.tmap_a, bar(B_AFULL), kc * BLOCK_K, tile * BLOCK_M);
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
    constexpr uint32_t idesc = ptx::make_idesc_bf16(BLOCK_M, 128);
    int slot = 0;
    uint32_t sphase = 0;
    auto mma1 = [&](uint32_t G) {
      const uint32_t hb = G & 1u, n = G >> 1;
      const int r = (int)(G % ROUNDS);
      if (r == 0) MF_TIMED_WAIT(0, bar(B_AFULL), (G / ROUNDS) & 1u);
      MF_TIMED_WAIT(1, bar(B_HEMPTY + hb), (n & 1u) ^ 1u);
      MF_TIMED_WAIT(2, bar(B_RFULL + slot), sphase);
      ptx::tc_fence_after();
#pragma unroll
      for (int kc = 0; kc < KC1; ++kc) {
        const uint64_t adesc = ptx::make_sw128_kmajor_desc(a_smem + kc * A_STAGE_BYTES);
        const uint64_t bdesc = ptx::make_sw128_kmajor_desc(ring_smem + slot * MF_SLOT_BYTES + kc * CHUNK_BYTES);
#pragma unroll
        for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk)
          ptx::tc_mma_bf16(tmem_base + hb * 128, adesc + 2u * kk, bdesc + 2u * kk, idesc, (kc | kk) != 0 ? 1u : 0u);
      }
      ptx::tc_commit(bar(B_REMPTY + slot));
      if (++slot == MF_SLOTS) { slot = 0; sphase ^= 1; }
      if (r == ROUNDS - 1) ptx::tc_commit(bar(B_AEMPTY));   // A tile consumed
      ptx::tc_commit(bar(B_HFULL + hb));
    };
    auto mma2 = [&](uint32_t G) {
      const uint32_t hb = G & 1u, n = G >> 1;
      const int r = (int)(G % ROUNDS);
      const uint32_t it = G / ROUNDS, ob = it & 1u;
      MF_TIMED_WAIT(3, bar(B_APFULL + hb), n & 1u);
      if (r == 0) MF_TIMED_WAIT(4, bar(B_OEMPTY + ob), ((it >> 1) & 1u) ^ 1u);
      MF_TIMED_WAIT(2, bar(B_RFULL + slot), sphase);
      ptx::tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const uint64_t adesc = ptx::make_sw128_kmajor_desc(ap_smem + hb * AP_BYTES + c * A_STAGE_BYTES);
        const uint64_t bdesc = ptx::make_sw128_kmajor_desc(ring_smem + slot * MF_SLOT_BYTES + c * CHUNK_BYTES);
#pragma unroll
        for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk)
          ptx::tc_mma_bf16(tmem_base + O_COL + ob * 128, adesc + 2u * kk, bdesc + 2u * kk, idesc, (r | c | kk) != 0 ? 1u : 0u);
      }
      ptx::tc_commit(bar(B_REMPTY + slot));
      if (++slot == MF_SLOTS) { slot = 0; sphase ^= 1; }
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
    const int q = ew & 3, half = ew >> 2;      // TMEM lane quarter / 64-column half of the round = A' k-chunk
    const int row = q * 32 + lane;
    const int sw = row & 7;
    uint8_t* ap_gen = smem_gen + A_BYTES;
    uint32_t G = 0;
    // E2 (x += gamma * (O + b2)) of a tile is issued one tile LATE, as two 128-byte-column groups interleaved with the
    // E1 rounds of the next tile: by then O is long complete (no wait on the last MMA-2) and the TMA reduce-add of the
    // previous group has long finished reading the staging buffer (the L2 reduce path sustains only ~16 B/clk per SM).
    auto e2_group = [&](int jt, int grp) {
      if (jt < 0) return;
      const uint32_t job = (uint32_t)jt & 1u;
      const long long te1 = tr ? clock64() : 0;
      if (grp == 0) {
        ptx::mbar_wait(bar(B_OFULL + job), ((uint32_t)jt >> 1) & 1u);
        ptx::tc_fence_after();
      }
      const int jtile = (int)blockIdx.x + jt * (int)gridDim.x;
      epilogue_tile_tma<C, EPI_RESID, true, NUM_EPI_WARPS>(p, jtile, 0, tmem_base + O_COL + job * 128, ew, lane,
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
        MF_TIMED_WAIT(0, bar(B_HFULL + hb), n & 1u);
        MF_TIMED_WAIT(1, bar(B_APEMPTY + hb), (n & 1u) ^ 1u);   // MMA-2 that last read this A' buffer has retired
        ptx::tc_fence_after();
        const long long te0 = tr ? clock64() : 0;
        const uint32_t tmem_row = tmem_base + hb * 128 + ((uint32_t)(q * 32) << 16) + half * 64;
        uint4* dst = reinterpret_cast<uint4*>(ap_gen + hb * AP_BYTES + half * A_STAGE_BYTES + row * 128);
        const float* b1 = fp.b1 + r * 128 + half * 64;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float v[32];
          tmem_load_chunk<32>(tmem_row + c * 32, v);
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            const float4 ba = __ldg(reinterpret_cast<const float4*>(b1 + c * 32 + j));
            const float4 bb = __ldg(reinterpret_cast<const float4*>(b1 + c * 32 + j + 4));
            uint4 w;
            w.x = gelu_pack2_f16(v[j] + ba.x, v[j + 1] + ba.y); w.y = gelu_pack2_f16(v[j + 2] + ba.z, v[j + 3] + ba.w);
            w.z = gelu_pack2_f16(v[j + 4] + bb.x, v[j + 5] + bb.y); w.w = gelu_pack2_f16(v[j + 6] + bb.z, v[j + 7] + bb.w);
            dst[((c * 4) + (j >> 3)) ^ sw] = w;
          }
        }
        ptx::tc_fence_before();
        ptx::fence_proxy_async();   // A' was written through the generic proxy, tcgen05.mma reads it through the async proxy
        __syncwarp();
        if (lane == 0) { ptx::mbar_arrive(bar(B_HEMPTY + hb)); ptx::mbar_arrive(bar(B_APFULL + hb)); }
        if (tr) tw[2] += clock64() - te0;
        if (r == 0) e2_group(it - 1, 0);   // two rounds apart: each 32 KB batch of reduce-adds drains (~16 B/clk) before the next
        if (r == 2) e2_group(it - 1, 1);
      }
    }
    e2_group(my_tiles - 1, 0);
    e2_group(my_tiles - 1, 1);
    if (lane == 0) ptx::bulk_wait0();
    if (tr && ew == 0 && lane == 0) for (int i = 0; i < 5; ++i) p.trace[8 + i] = tw[i];
  }
#undef MF_TIMED_WAIT
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

template <int BLOCK_N, int EPI>
int launch_inst(const GemmPlan& plan, cudaStream_t stream) {
  using C = Cfg<BLOCK_N>;
  static bool configured = false;
  auto kfn = gemm_tc_kernel<BLOCK_N, EPI>;
  if (!configured) {
    GDRN_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    configured = true;
  }
  int total = plan.m_tiles * plan.n_tiles;
  if (total <= 0) return GDRN_OK;
  int grid = total < gdrn_num_sms() ? total : gdrn_num_sms();
  kfn<<<grid, NUM_THREADS, C::SMEM_BYTES, stream>>>(plan);
  GDRN_CHECK_CUDA(cudaGetLastError());
  gdrn_count_launch(1);
  return GDRN_OK;
}

}  // namespace

int gemm_tc_launch(const GemmPlan& plan_in, int block_n, cudaStream_t stream) {
  GemmPlan plan = plan_in;
  {  // experiment knob: GDRN_TILE_ORDER=0/1 forces the tile order of every multi-n-tile GEMM
    static int order = -2;
    if (order == -2) { const char* e = getenv("GDRN_TILE_ORDER"); order = e ? atoi(e) : -1; }
    if (order >= 0) plan.n_major = order;
  }
  GDRN_REQUIRE(plan.a_rank == 2 || plan.a_rank == 4 || plan.a_rank == 5, "gemm: bad a_rank");
  plan.use_tma_store = 0;
  plan.resid_reduce = 0;
  {
    static int tma_epi = -1;  // GDRN_TMA_STORE=0 falls back to the gather-store epilogue (A/B experiments)
    if (tma_epi < 0) { const char* e = getenv("GDRN_TMA_STORE"); tma_epi = e ? atoi(e) : 1; }
    const bool f32 = plan.epi == EPI_RESID || plan.epi == EPI_BIAS_LN || (plan.epi == EPI_STORE && plan.out_f32);
    if (tma_epi && plan.a_rank == 2 && !plan.split && block_n >= 128 && plan.N % 64 == 0 &&
        (plan.epi == EPI_STORE || plan.epi == EPI_GELU || plan.epi == EPI_RESID || plan.epi == EPI_BIAS_LN) &&
        ((uintptr_t)plan.out % 16 == 0) && ((plan.ldo * (f32 ? 4 : 2)) % 16 == 0) &&
        (plan.epi != EPI_RESID || ((uintptr_t)plan.resid % 16 == 0))) {
      const uint64_t dims[2] = {(uint64_t)plan.ldo, (uint64_t)plan.M};
      const uint64_t str[1] = {(uint64_t)plan.ldo * (f32 ? 4 : 2)};
      const uint32_t box[2] = {f32 ? 32u : 64u, 32u};
      int rc = make_tmap_store(&plan.tmap_out, plan.out, f32 ? 1 : 0, dims, str, box);
      if (rc != GDRN_OK) return rc;
      plan.use_tma_store = 1;
      static int red = -1;  // GDRN_RESID_REDUCE=0: read-modify-write in the SM instead of the L2 reduce-add
      if (red < 0) { const char* e = getenv("GDRN_RESID_REDUCE"); red = e ? atoi(e) : 1; }
      plan.resid_reduce = (red && plan.epi == EPI_RESID && plan.resid == plan.out) ? 1 : 0;
    }
  }
  GDRN_REQUIRE(plan.num_taps >= 1 && plan.num_taps <= GEMM_MAX_TAPS, "gemm: bad num_taps");
  {
    // CTA-pair path: rank-2, one tap, BLOCK_N = 256, TMA-store epilogue, enough tiles to fill the pairs
    static int pair_on = -1;  // GDRN_GEMM_PAIR=0 disables (A/B experiments)
    if (pair_on < 0) { const char* e = getenv("GDRN_GEMM_PAIR"); pair_on = e ? atoi(e) : 1; }
    // measured (profiles/): the pair kernel wins when the mainloop dominates (K >= 1024: fc2, stage-3 fc1) and loses
    // a few % on short-K tiles whose time is the GELU epilogue plus per-tile pair handshakes
    static int gelu16 = -1;   // GDRN_GELU16=0: GELU GEMMs keep the 8-warp epilogue (and the K threshold below)
    if (gelu16 < 0) { const char* e = getenv("GDRN_GELU16"); gelu16 = e ? atoi(e) : 1; }
    static int pair_min_k = -1;
    if (pair_min_k < 0) { const char* e = getenv("GDRN_GEMM_PAIR_MIN_KCHUNKS"); pair_min_k = e ? atoi(e) : 16; }
    if (pair_on && plan.use_tma_store && block_n == 256 && plan.num_taps == 1 && plan.taps[0].c0 == 0 &&
        plan.taps[0].d1 == 0 && plan.taps[0].b_off == 0 && plan.b_ptr != nullptr && plan.N % 256 == 0 &&
        (plan.epi == EPI_GELU || plan.epi == EPI_RESID || plan.epi == EPI_STORE) && plan.M >= 256 * 16 &&
        (plan.k_chunks >= pair_min_k || (plan.epi == EPI_GELU && gelu16))) {
      const uint64_t dims[2] = {(uint64_t)plan.b_ktot, (uint64_t)plan.b_rows};
      const uint64_t str[1] = {(uint64_t)plan.b_ktot * 2};
      const uint32_t box[2] = {64, 128};
      int rc = make_tmap_bf16(&plan.tmap_b, plan.b_ptr, 2, dims, str, box);
      if (rc != GDRN_OK) return rc;
      if (plan.epi == EPI_GELU) return gelu16 ? launch_pair<EPI_GELU, 16>(plan, stream) : launch_pair<EPI_GELU, 8>(plan, stream);
      if (plan.epi == EPI_RESID) return launch_pair<EPI_RESID, 8>(plan, stream);
      return launch_pair<EPI_STORE, 8>(plan, stream);
    }
  }
#define GDRN_GEMM_CASE(BN, E) \
  if (block_n == BN && plan.epi == E) return launch_inst<BN, E>(plan, stream);
  GDRN_GEMM_CASE(256, EPI_GELU)
  GDRN_GEMM_CASE(256, EPI_RESID)
  GDRN_GEMM_CASE(256, EPI_STORE)
  GDRN_GEMM_CASE(256, EPI_GNSTATS)
  GDRN_GEMM_CASE(128, EPI_GELU)
  GDRN_GEMM_CASE(128, EPI_RESID)
  GDRN_GEMM_CASE(128, EPI_STORE)
  GDRN_GEMM_CASE(128, EPI_GNSTATS)
  GDRN_GEMM_CASE(128, EPI_BIAS_LN)
  GDRN_GEMM_CASE(64, EPI_GELU)
  GDRN_GEMM_CASE(64, EPI_STORE)
  GDRN_GEMM_CASE(64, EPI_GNSTATS)
  GDRN_GEMM_CASE(16, EPI_STORE)
  GDRN_GEMM_CASE(80, EPI_OUTCONV)
#undef GDRN_GEMM_CASE
  gdrn_set_last_error(__FILE__, __LINE__, "gemm: unsupported (block_n, epilogue) combination");
  return GDRN_ERR_INVALID;
}

int mlp_fused_supported(int C, long long M) { return C == 128 && M % 128 == 0 && M >= 128 * 148; }

int mlp_fused_launch(const void* A, const void* W1, const float* b1, const void* W2, const float* b2, const float* gamma,
                     float* x, long long M, int C, cudaStream_t stream) {
  GDRN_REQUIRE(mlp_fused_supported(C, M), "mlp_fused: unsupported shape");
  MlpFusedPlan fp;
  memset(&fp, 0, sizeof(fp));
  {
    const uint64_t d[2] = {(uint64_t)C, (uint64_t)M}; const uint64_t st[1] = {(uint64_t)C * 2}; const uint32_t bx[2] = {64, 128};
    int rc = make_tmap_bf16(&fp.tmap_a, A, 2, d, st, bx);
    if (rc != GDRN_OK) return rc;
  }
  {
    const uint64_t d[2] = {(uint64_t)C, (uint64_t)4 * C}; const uint64_t st[1] = {(uint64_t)C * 2}; const uint32_t bx[2] = {64, 128};
    int rc = make_tmap_bf16(&fp.tmap_w1, W1, 2, d, st, bx);
    if (rc != GDRN_OK) return rc;
  }
  {
    const uint64_t d[2] = {(uint64_t)4 * C, (uint64_t)C}; const uint64_t st[1] = {(uint64_t)4 * C * 2}; const uint32_t bx[2] = {64, (uint32_t)C};
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
  constexpr int C_ = 128;
  constexpr int SMEM = (C_ / 64) * A_STAGE_BYTES + 2 * 2 * A_STAGE_BYTES + MF_SLOTS * MF_SLOT_BYTES + NUM_EPI_WARPS * EPI_STAGE_BYTES + 256 + 1024;
  auto kfn = mlp_fused_kernel<C_>;
  static bool configured = false;
  if (!configured) {
    GDRN_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    configured = true;
  }
  const int grid = fp.m_tiles < gdrn_num_sms() ? fp.m_tiles : gdrn_num_sms();
  static int trace_on = -1;
  if (trace_on < 0) trace_on = getenv("GDRN_MLP_TRACE") ? 1 : 0;
  static long long* d_trace = nullptr;
  if (trace_on) {
    if (!d_trace) GDRN_CHECK_CUDA(cudaMalloc(&d_trace, 16 * sizeof(long long)));
    GDRN_CHECK_CUDA(cudaMemsetAsync(d_trace, 0, 16 * sizeof(long long), stream));
    fp.g.trace = d_trace;
  }
  kfn<<<grid, NUM_THREADS, SMEM, stream>>>(fp);
  GDRN_CHECK_CUDA(cudaGetLastError());
  gdrn_count_launch(1);
  if (trace_on) {
    long long h[16];
    GDRN_CHECK_CUDA(cudaMemcpyAsync(h, d_trace, sizeof(h), cudaMemcpyDeviceToHost, stream));
    GDRN_CHECK_CUDA(cudaStreamSynchronize(stream));
    fprintf(stderr, "[mlp fused trace] cta0 cycles=%lld tiles=%lld | mma waits: A=%lld H-empty=%lld ring=%lld A'-full=%lld O-empty=%lld | "
                    "epi0: H-full wait=%lld A'-empty wait=%lld E1 busy=%lld O-full wait=%lld E2 busy=%lld\n",
            h[5], h[6], h[0], h[1], h[2], h[3], h[4], h[8], h[9], h[10], h[11], h[12]);
  }
  return GDRN_OK;
}

// ------------------------------------------------------------------------------------------------
// Tensor maps (driver entry point resolved at run time: the library does not link libcuda).
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  }
  return fn;
}

int make_tmap_f32_plain(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                        const uint64_t* strides_bytes, const uint32_t* box) {
  PFN_encodeTiled fn = get_encode_fn();
  GDRN_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[160];
    snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled(f32) failed: CUresult %d", (int)r);
    gdrn_set_last_error(__FILE__, __LINE__, msg);
    return GDRN_ERR_CUDA;
  }
  return GDRN_OK;
}

// rank-2 output map for the TMA-store epilogue: 128-byte rows, SWIZZLE_128B
int make_tmap_store(CUtensorMap* out, const void* base, int is_f32, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box) {
  PFN_encodeTiled fn = get_encode_fn();
  GDRN_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  cuuint64_t gdim[2] = {dims[0], dims[1]};
  cuuint64_t gstr[1] = {strides_bytes[0]};
  cuuint32_t bx[2] = {box[0], box[1]};
  cuuint32_t es[2] = {1, 1};
  CUresult r = fn(out, is_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                  const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[160];
    snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled(store) failed: CUresult %d (dims %llu,%llu)", (int)r,
             (unsigned long long)dims[0], (unsigned long long)dims[1]);
    gdrn_set_last_error(__FILE__, __LINE__, msg);
    return GDRN_ERR_CUDA;
  }
  return GDRN_OK;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box) {
  PFN_encodeTiled fn = get_encode_fn();
  GDRN_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[160];
    snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled failed: CUresult %d (rank %d dims %llu,%llu box %u,%u)", (int)r,
             rank, (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
    gdrn_set_last_error(__FILE__, __LINE__, msg);
    return GDRN_ERR_CUDA;
  }
  return GDRN_OK;
}
