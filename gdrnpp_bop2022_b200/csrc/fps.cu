// Farthest point sampling on the GPU, bit-exact with the reference CPU implementation
// (core/csrc/fps/src/farthest_point_sampling.cpp:40-160; built by fps/setup.py:5-7 with gcc -O2, i.e.
// NO fused multiply-add, float accumulation order (x*x + y*y) + z*z).
//
// One CTA per cloud.  The cloud (x,y,z SoA) and the running min-distance live in shared memory (up to
// 14 000 points) or, for larger clouds, min-distance in shared memory and the points streamed from
// L2 (up to 56 000 points); beyond that a cluster of 2 / 4 / 8 CTAs shares one cloud (fps_cluster_kernel, up to
// 448 000 points).  Every iteration: each thread relaxes its points against the last pick
// and keeps a local (value, index) arg-max; one shuffle tree + one __syncthreads per pick.
// Algorithmic traffic: 12*pn + 4*sn bytes per cloud; the op is latency-bound on sn block reductions.
#include <cooperative_groups.h>
#include <float.h>
#include <stdlib.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int FPS_THREADS = 1024;
constexpr int FPS_SMEM_ALL = 14000;  // points with xyz+d resident: 16 B each  (224 000 B + 1.3 KB static)
constexpr int FPS_SMEM_D = 56000;    // points with only d resident: 4 B each

struct Cand {
  float v;
  int i;
};

// reference find_max_dist_idx (:56-73): strict '>', scan order => lowest index among equal maxima,
// and (0, index 0) when nothing is > 0.
__device__ __forceinline__ Cand better(Cand a, Cand b) {
  if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
  return a;
}

__device__ __forceinline__ Cand warp_argmax(Cand c) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Cand t;
    t.v = __shfl_xor_sync(0xffffffffu, c.v, o);
    t.i = __shfl_xor_sync(0xffffffffu, c.i, o);
    c = better(c, t);
  }
  return c;
}

__device__ __forceinline__ float sqdist_nofma(float ax, float ay, float az, float bx, float by, float bz) {
  // Vec3::operator- then squared_norm(): x*x + y*y + z*z, each op rounded (no contraction)
  float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

template <bool RESIDENT>
__global__ void __launch_bounds__(FPS_THREADS, 1)
fps_kernel(const float* __restrict__ pts_all, int* __restrict__ idx_all, int pn, int sn,
           const int* __restrict__ start_idx) {
  extern __shared__ float sm[];
  const int cloud = blockIdx.x;
  const float* pts = pts_all + (size_t)cloud * pn * 3;
  int* out = idx_all + (size_t)cloud * sn;
  float* sd = sm;                          // [pn] running min distance (negative = already picked)
  float* sx = sm + pn;                     // RESIDENT: [pn] x, y, z
  float* sy = sx + pn;
  float* sz = sy + pn;
  __shared__ Cand red[2][32];
  __shared__ float redf[6][32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (RESIDENT) {
    for (int i = tid; i < pn; i += FPS_THREADS) {
      sx[i] = pts[3 * i];
      sy[i] = pts[3 * i + 1];
      sz[i] = pts[3 * i + 2];
    }
  }
  auto PX = [&](int i) { return RESIDENT ? sx[i] : __ldg(pts + 3 * i); };
  auto PY = [&](int i) { return RESIDENT ? sy[i] : __ldg(pts + 3 * i + 1); };
  auto PZ = [&](int i) { return RESIDENT ? sz[i] : __ldg(pts + 3 * i + 2); };
  __syncthreads();

  int cur;
  if (start_idx == nullptr) {
    // init_center (:118-160): bbox centre = (max+min)*(1/2), min_dist = |p - centre|^2
    float mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX}, mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
    for (int i = tid; i < pn; i += FPS_THREADS) {
      float x = PX(i), y = PY(i), z = PZ(i);
      mx[0] = fmaxf(mx[0], x); mx[1] = fmaxf(mx[1], y); mx[2] = fmaxf(mx[2], z);
      mn[0] = fminf(mn[0], x); mn[1] = fminf(mn[1], y); mn[2] = fminf(mn[2], z);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
        mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
      }
      if (lane == 0) { redf[k][warp] = mx[k]; redf[3 + k][warp] = mn[k]; }
    }
    __syncthreads();
    float c[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float a = redf[k][lane], b = redf[3 + k][lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a = fmaxf(a, __shfl_xor_sync(0xffffffffu, a, o));
        b = fminf(b, __shfl_xor_sync(0xffffffffu, b, o));
      }
      c[k] = __fmul_rn(__fadd_rn(a, b), 0.5f);  // Vec3 operator/(2.f) == *(1.f/2.f)
    }
    Cand best = {0.f, 0};
    for (int i = tid; i < pn; i += FPS_THREADS) {
      float d = sqdist_nofma(PX(i), PY(i), PZ(i), c[0], c[1], c[2]);
      d = fminf(d, FLT_MAX);
      sd[i] = d;
      best = better(best, Cand{d, i});
    }
    best = warp_argmax(best);
    if (lane == 0) red[1][warp] = best;
    __syncthreads();
    Cand t = warp_argmax(red[1][lane]);
    cur = t.i;
  } else {
    for (int i = tid; i < pn; i += FPS_THREADS) sd[i] = FLT_MAX;
    cur = start_idx[cloud];
    __syncthreads();
  }

  for (int s = 0; s < sn; ++s) {
    if (tid == 0) out[s] = cur;
    if (s == sn - 1) break;
    const float cx = PX(cur), cy = PY(cur), cz = PZ(cur);
    Cand best = {0.f, 0};
    for (int i = tid; i < pn; i += FPS_THREADS) {
      float d = sd[i];
      if (i == cur) { d = -1.f; sd[i] = d; }  // mask[cur] = true
      if (d >= 0.f) {
        float nd = sqdist_nofma(PX(i), PY(i), PZ(i), cx, cy, cz);
        if (nd < d) { d = nd; sd[i] = d; }
        best = better(best, Cand{d, i});
      }
    }
    best = warp_argmax(best);
    if (lane == 0) red[s & 1][warp] = best;
    __syncthreads();
    Cand t = warp_argmax(red[s & 1][lane]);
    cur = t.i;
  }
}

// Clouds beyond one CTA's shared memory (pn > 56 000; BOP meshes reach 10^5 vertices): a thread-block
// CLUSTER of CS CTAs per cloud.  CTA `rank` owns points [rank*chunk, (rank+1)*chunk) with its slice of the running
// min-distance in its own shared memory (points streamed from L2); per pick every CTA pushes its local arg-max
// candidate into all peers' shared memory (distributed shared memory stores), one cluster barrier, and every CTA
// reduces the CS candidates identically (strict '>' then lowest index: the reference's scan order).  Same arithmetic,
// same tie-breaking => the same indices as the single-CTA kernel and the reference.
template <int CS>
__global__ void __launch_bounds__(FPS_THREADS, 1)
fps_cluster_kernel(const float* __restrict__ pts_all, int* __restrict__ idx_all, int pn, int sn,
                   const int* __restrict__ start_idx) {
  extern __shared__ float sm[];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cloud = blockIdx.x / CS;
  const float* pts = pts_all + (size_t)cloud * pn * 3;
  int* out = idx_all + (size_t)cloud * sn;
  const int chunk = (pn + CS - 1) / CS;
  const int i0 = rank * chunk, i1 = min(pn, i0 + chunk);
  float* sd = sm;                         // [chunk] running min distance of this CTA's points
  __shared__ Cand red[32];
  __shared__ Cand xch[2][CS];             // candidates of every rank, double-buffered by pick parity
  __shared__ float redf[6][32];
  __shared__ float xbb[CS][6];            // bbox partials of every rank
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  auto PX = [&](int i) { return __ldg(pts + 3 * i); };
  auto PY = [&](int i) { return __ldg(pts + 3 * i + 1); };
  auto PZ = [&](int i) { return __ldg(pts + 3 * i + 2); };
  // block arg-max of `best` -> pushed to every rank's xch[buf][rank]; cluster barrier; identical final reduction
  auto cluster_argmax = [&](Cand best, int buf) {
    best = warp_argmax(best);
    if (lane == 0) red[warp] = best;
    __syncthreads();
    if (warp == 0) {
      Cand t = warp_argmax(red[lane]);
      if (lane < CS) *cluster.map_shared_rank(&xch[buf][rank], lane) = t;
    }
    cluster.sync();
    Cand t = xch[buf][0];
#pragma unroll
    for (int r = 1; r < CS; ++r) t = better(t, xch[buf][r]);
    return t.i;
  };

  int cur;
  if (start_idx == nullptr) {
    float mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX}, mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
    for (int i = i0 + tid; i < i1; i += FPS_THREADS) {
      float x = PX(i), y = PY(i), z = PZ(i);
      mx[0] = fmaxf(mx[0], x); mx[1] = fmaxf(mx[1], y); mx[2] = fmaxf(mx[2], z);
      mn[0] = fminf(mn[0], x); mn[1] = fminf(mn[1], y); mn[2] = fminf(mn[2], z);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
        mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
      }
      if (lane == 0) { redf[k][warp] = mx[k]; redf[3 + k][warp] = mn[k]; }
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        float a = redf[k][lane], b = redf[3 + k][lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          a = fmaxf(a, __shfl_xor_sync(0xffffffffu, a, o));
          b = fminf(b, __shfl_xor_sync(0xffffffffu, b, o));
        }
        if (lane < CS) { *cluster.map_shared_rank(&xbb[rank][k], lane) = a; *cluster.map_shared_rank(&xbb[rank][3 + k], lane) = b; }
      }
    }
    cluster.sync();
    float c[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float a = xbb[0][k], b = xbb[0][3 + k];
#pragma unroll
      for (int r = 1; r < CS; ++r) { a = fmaxf(a, xbb[r][k]); b = fminf(b, xbb[r][3 + k]); }   // max / min are order independent
      c[k] = __fmul_rn(__fadd_rn(a, b), 0.5f);
    }
    Cand best = {0.f, 0};
    for (int i = i0 + tid; i < i1; i += FPS_THREADS) {
      float d = sqdist_nofma(PX(i), PY(i), PZ(i), c[0], c[1], c[2]);
      d = fminf(d, FLT_MAX);
      sd[i - i0] = d;
      best = better(best, Cand{d, i});
    }
    cur = cluster_argmax(best, 1);
  } else {
    for (int i = i0 + tid; i < i1; i += FPS_THREADS) sd[i - i0] = FLT_MAX;
    cur = start_idx[cloud];
    __syncthreads();
  }

  for (int s = 0; s < sn; ++s) {
    if (tid == 0 && rank == 0) out[s] = cur;
    if (s == sn - 1) break;
    const float cx = PX(cur), cy = PY(cur), cz = PZ(cur);
    Cand best = {0.f, 0};
    for (int i = i0 + tid; i < i1; i += FPS_THREADS) {
      float d = sd[i - i0];
      if (i == cur) { d = -1.f; sd[i - i0] = d; }
      if (d >= 0.f) {
        float nd = sqdist_nofma(PX(i), PY(i), PZ(i), cx, cy, cz);
        if (nd < d) { d = nd; sd[i - i0] = d; }
        best = better(best, Cand{d, i});
      }
    }
    cur = cluster_argmax(best, s & 1);
  }
  cluster.sync();   // no CTA exits while a peer may still write its shared memory
}

template <int CS>
int fps_launch_cluster(const float* pts, int* idxs, int pn, int sn, int batch, const int* start_idx, cudaStream_t st) {
  auto kfn = fps_cluster_kernel<CS>;
  const size_t smem = (size_t)((pn + CS - 1) / CS) * 4;
  GDRN_OPT_IN_SMEM(kfn, 224000);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(batch * CS);
  cfg.blockDim = dim3(FPS_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  GDRN_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kfn, pts, idxs, pn, sn, start_idx));
  gdrn_count_launch(1);
  return GDRN_OK;
}

int fps_launch(const float* pts, int* idxs, int pn, int sn, int batch, const int* start_idx, cudaStream_t st) {
  GDRN_REQUIRE(pn > 0 && sn > 0 && batch > 0, "fps: pn, sn, batch must be positive");
  GDRN_REQUIRE(pn <= 8 * FPS_SMEM_D, "fps: more than 448000 points per cloud are not supported (8-CTA cluster x 56000)");
  if (pn > FPS_SMEM_D) {   // cluster of 2 / 4 / 8 CTAs per cloud
    if (pn <= 2 * FPS_SMEM_D) return fps_launch_cluster<2>(pts, idxs, pn, sn, batch, start_idx, st);
    if (pn <= 4 * FPS_SMEM_D) return fps_launch_cluster<4>(pts, idxs, pn, sn, batch, start_idx, st);
    return fps_launch_cluster<8>(pts, idxs, pn, sn, batch, start_idx, st);
  }
  {
    // Few large clouds (a mesh being sampled, not a batch of ROI clouds): one CTA per cloud leaves most SMs idle and, beyond
    // 14 000 points, re-reads the coordinates from L2 every step.  Spread each cloud over a cluster when the SMs are there:
    // the per-step scan shrinks by the cluster size and the slices fit shared memory again.  Same arg-max order, same bits.
    static int min_pn = -1;   // GDRN_FPS_CLUSTER_MIN_PN (0 = never)
    if (min_pn < 0) { const char* e = getenv("GDRN_FPS_CLUSTER_MIN_PN"); min_pn = e ? atoi(e) : 16384; }
    if (min_pn > 0 && pn >= min_pn) {
      const int room = gdrn_num_sms() / batch;
      if (room >= 8) return fps_launch_cluster<8>(pts, idxs, pn, sn, batch, start_idx, st);
      if (room >= 4) return fps_launch_cluster<4>(pts, idxs, pn, sn, batch, start_idx, st);
      if (room >= 2) return fps_launch_cluster<2>(pts, idxs, pn, sn, batch, start_idx, st);
    }
  }
  GDRN_OPT_IN_SMEM(fps_kernel<true>, 224000);
  GDRN_OPT_IN_SMEM(fps_kernel<false>, 224000);
  if (pn <= FPS_SMEM_ALL) {
    fps_kernel<true><<<batch, FPS_THREADS, (size_t)pn * 16, st>>>(pts, idxs, pn, sn, start_idx);
  } else {
    fps_kernel<false><<<batch, FPS_THREADS, (size_t)pn * 4, st>>>(pts, idxs, pn, sn, start_idx);
  }
  GDRN_CHECK_CUDA(cudaGetLastError());
  gdrn_count_launch(1);
  return GDRN_OK;
}

unsigned g_fps_seed = 0;
bool g_fps_seeded = false;

void fps_host(float* pts, int* idxs, int pn, int sn, bool init_center) {
  // Host-pointer entry with the reference's cffi signature (void return, no error channel):
  // failures are reported on stderr and leave idxs zero-filled.
  for (int i = 0; i < sn; ++i) idxs[i] = 0;
  if (pn <= 0 || sn <= 0) return;
  float* d_pts = nullptr;
  int* d_idx = nullptr;
  int* d_start = nullptr;
  cudaError_t e = cudaMalloc(&d_pts, (size_t)pn * 12);
  if (e == cudaSuccess) e = cudaMalloc(&d_idx, (size_t)sn * 4 + 4);
  if (e == cudaSuccess) e = cudaMemcpy(d_pts, pts, (size_t)pn * 12, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && !init_center) {
    // reference: srand(time(0)); cur_idx = rand() % pn (:93-94).  We keep rand()%pn but let the caller seed.
    if (g_fps_seeded) { srand(g_fps_seed); g_fps_seeded = false; }
    int start = rand() % pn;
    d_start = d_idx + sn;
    e = cudaMemcpy(d_start, &start, 4, cudaMemcpyHostToDevice);
  }
  int rc = GDRN_OK;
  if (e == cudaSuccess) rc = fps_launch(d_pts, d_idx, pn, sn, 1, d_start, 0);
  if (e == cudaSuccess && rc == GDRN_OK) e = cudaMemcpy(idxs, d_idx, (size_t)sn * 4, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess || rc != GDRN_OK)
    fprintf(stderr, "gdrn farthest_point_sampling failed: %s\n", e != cudaSuccess ? cudaGetErrorString(e) : gdrn_last_error());
  cudaFree(d_pts);
  cudaFree(d_idx);
}

}  // namespace

extern "C" int gdrn_fps_cuda(const float* pts, int* idxs, int pn, int sn, int batch, const int* start_idx,
                             void* stream) {
  return fps_launch(pts, idxs, pn, sn, batch, start_idx, (cudaStream_t)stream);
}
extern "C" void gdrn_fps_set_seed(unsigned seed) {
  g_fps_seed = seed;
  g_fps_seeded = true;
}
extern "C" void farthest_point_sampling(float* pts, int* idxs, int pn, int sn) { fps_host(pts, idxs, pn, sn, false); }
extern "C" void farthest_point_sampling_init_center(float* pts, int* idxs, int pn, int sn) {
  fps_host(pts, idxs, pn, sn, true);
}
