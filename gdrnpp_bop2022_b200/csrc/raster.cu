// CUDA depth rasteriser (no EGL / GL interop) + the fast depth-refinement step.
//
// Replaces, for depth output, lib/render_vispy/renderer.py:126-182,363-407,461-476 (the renderer
// fast depth refine calls) and lib/egl_renderer/egl_renderer_v3.py:755-783,838-1228 (pc_cam[...,2]).
// GL conventions reproduced: camera-space (X,Y,Z) in OpenCV axes projects to
// u = fx*X/Z + skew*Y/Z + cx, v = fy*Y/Z + cy; pixel (row r, col c) is sampled at (c+0.5, r+0.5);
// coverage = sample inside the projected triangle (either winding, no culling: renderer.py:102);
// the z-buffer value is affine in window space, i.e. 1/Z is interpolated with screen-space
// barycentrics (== the perspective-correct interpolation of camera-space Z in the EGL shader);
// nearest fragment wins; background = 0.  Triangles with a vertex closer than znear are dropped
// (GL would clip them; objects in front of the camera never hit this).
//
// Algorithmic HBM traffic per render: 12*V + 12*F bytes of mesh + 8*H*W z-buffer scratch + 4*H*W out.
// One thread per (render, triangle); 64-bit atomicMin on (z_bits << 32 | triangle) resolves visibility.
#include "common.cuh"

namespace {

__global__ void rast_clear_kernel(unsigned long long* zbuf, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    zbuf[i] = 0xFFFFFFFFFFFFFFFFull;
}

__global__ void __launch_bounds__(256)
rast_tri_kernel(const float* __restrict__ verts, const int* __restrict__ faces, int V, int F,
                const float* __restrict__ poses, const float* __restrict__ Ks, int H, int W, float znear, float zfar,
                unsigned long long* __restrict__ zbuf) {
  const int r = blockIdx.y;
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  const float* P = poses + r * 12;
  const float* K = Ks + r * 9;
  float u[3], v[3], iz[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int vi = faces[f * 3 + k];
    const float x = verts[vi * 3], y = verts[vi * 3 + 1], z = verts[vi * 3 + 2];
    const float X = P[0] * x + P[1] * y + P[2] * z + P[3];
    const float Y = P[4] * x + P[5] * y + P[6] * z + P[7];
    const float Z = P[8] * x + P[9] * y + P[10] * z + P[11];
    if (Z < znear) return;
    iz[k] = 1.0f / Z;
    u[k] = (K[0] * X + K[1] * Y) * iz[k] + K[2];
    v[k] = (K[4] * Y) * iz[k] + K[5];
  }
  const float area = (u[1] - u[0]) * (v[2] - v[0]) - (u[2] - u[0]) * (v[1] - v[0]);
  if (area == 0.f) return;
  const float inv_area = 1.0f / area;
  int c0 = max(0, (int)floorf(fminf(fminf(u[0], u[1]), u[2]) - 0.5f));
  int c1 = min(W - 1, (int)ceilf(fmaxf(fmaxf(u[0], u[1]), u[2]) - 0.5f));
  int r0 = max(0, (int)floorf(fminf(fminf(v[0], v[1]), v[2]) - 0.5f));
  int r1 = min(H - 1, (int)ceilf(fmaxf(fmaxf(v[0], v[1]), v[2]) - 0.5f));
  unsigned long long* zb = zbuf + (long long)r * H * W;
  for (int py = r0; py <= r1; ++py) {
    const float sy = (float)py + 0.5f;
    for (int px = c0; px <= c1; ++px) {
      const float sx = (float)px + 0.5f;
      // barycentrics (signed areas / area)
      const float w0 = ((u[1] - sx) * (v[2] - sy) - (u[2] - sx) * (v[1] - sy)) * inv_area;
      const float w1 = ((u[2] - sx) * (v[0] - sy) - (u[0] - sx) * (v[2] - sy)) * inv_area;
      const float w2 = 1.0f - w0 - w1;
      if (w0 < 0.f || w1 < 0.f || w2 < 0.f) continue;
      const float izp = w0 * iz[0] + w1 * iz[1] + w2 * iz[2];
      const float zp = 1.0f / izp;
      if (!(zp >= znear) || zp > zfar) continue;
      const unsigned long long key = ((unsigned long long)__float_as_uint(zp) << 32) | (unsigned)f;
      atomicMin(zb + (long long)py * W + px, key);
    }
  }
}

// ---- mesh registry: meshes uploaded once (rast_upload_mesh), renders pick one per ROI by id --------------------------
struct MeshEntry {
  const float* verts;
  const int* faces;
  int V, F;
};
constexpr int RAST_MAX_MESHES = 1024;

// one thread per (ROI, triangle of that ROI's mesh); grid.x covers the largest registered mesh
__global__ void __launch_bounds__(256)
rast_tri_multi_kernel(const MeshEntry* __restrict__ table, const int* __restrict__ mesh_ids, int n_meshes,
                      const float* __restrict__ poses, const float* __restrict__ Ks, int H, int W, float znear, float zfar,
                      unsigned long long* __restrict__ zbuf) {
  const int r = blockIdx.y;
  int mid = mesh_ids[r];
  if (mid < 0 || mid >= n_meshes) return;   // unknown id: the ROI renders as background
  const MeshEntry me = table[mid];
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= me.F) return;
  const float* P = poses + r * 12;
  const float* K = Ks + r * 9;
  float u[3], v[3], iz[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int vi = me.faces[f * 3 + k];
    const float x = me.verts[vi * 3], y = me.verts[vi * 3 + 1], z = me.verts[vi * 3 + 2];
    const float X = P[0] * x + P[1] * y + P[2] * z + P[3];
    const float Y = P[4] * x + P[5] * y + P[6] * z + P[7];
    const float Z = P[8] * x + P[9] * y + P[10] * z + P[11];
    if (Z < znear) return;
    iz[k] = 1.0f / Z;
    u[k] = (K[0] * X + K[1] * Y) * iz[k] + K[2];
    v[k] = (K[4] * Y) * iz[k] + K[5];
  }
  const float area = (u[1] - u[0]) * (v[2] - v[0]) - (u[2] - u[0]) * (v[1] - v[0]);
  if (area == 0.f) return;
  const float inv_area = 1.0f / area;
  int c0 = max(0, (int)floorf(fminf(fminf(u[0], u[1]), u[2]) - 0.5f));
  int c1 = min(W - 1, (int)ceilf(fmaxf(fmaxf(u[0], u[1]), u[2]) - 0.5f));
  int r0 = max(0, (int)floorf(fminf(fminf(v[0], v[1]), v[2]) - 0.5f));
  int r1 = min(H - 1, (int)ceilf(fmaxf(fmaxf(v[0], v[1]), v[2]) - 0.5f));
  unsigned long long* zb = zbuf + (long long)r * H * W;
  for (int py = r0; py <= r1; ++py) {
    const float sy = (float)py + 0.5f;
    for (int px = c0; px <= c1; ++px) {
      const float sx = (float)px + 0.5f;
      const float w0 = ((u[1] - sx) * (v[2] - sy) - (u[2] - sx) * (v[1] - sy)) * inv_area;
      const float w1 = ((u[2] - sx) * (v[0] - sy) - (u[0] - sx) * (v[2] - sy)) * inv_area;
      const float w2 = 1.0f - w0 - w1;
      if (w0 < 0.f || w1 < 0.f || w2 < 0.f) continue;
      const float izp = w0 * iz[0] + w1 * iz[1] + w2 * iz[2];
      const float zp = 1.0f / izp;
      if (!(zp >= znear) || zp > zfar) continue;
      const unsigned long long key = ((unsigned long long)__float_as_uint(zp) << 32) | (unsigned)f;
      atomicMin(zb + (long long)py * W + px, key);
    }
  }
}

__global__ void rast_resolve_kernel(const unsigned long long* __restrict__ zbuf, const float* __restrict__ Ks, int n,
                                    int H, int W, float znear, float zfar, int quantize_bits,
                                    float* __restrict__ depth, float* __restrict__ xyz_cam) {
  const long long total = (long long)n * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const unsigned long long key = zbuf[i];
    float z = 0.f;
    if (key != 0xFFFFFFFFFFFFFFFFull) {
      z = __uint_as_float((unsigned)(key >> 32));
      if (quantize_bits > 0) {
        // fixed-point window depth d = (1/z - 1/n) / (1/f - 1/n), read back as float32 and decoded like
        // renderer.py:176-182: dep = mult / (d + addi)
        const double dn = (double)znear, df = (double)zfar;
        double d = (1.0 / (double)z - 1.0 / dn) / (1.0 / df - 1.0 / dn);
        const double q = (double)((1ull << quantize_bits) - 1);
        d = floor(d * q + 0.5) / q;
        const float dfl = (float)d;
        const float mult = (float)((dn * df) / (dn - df)), addi = (float)(df / (dn - df));
        z = (dfl == 1.0f) ? 0.f : mult / (dfl + addi);
      }
    }
    depth[i] = z;
    if (xyz_cam) {
      const int px = (int)(i % W), py = (int)((i / W) % H);
      const int r = (int)(i / ((long long)H * W));
      const float* K = Ks + r * 9;
      float X = 0.f, Y = 0.f;
      if (z > 0.f) {
        Y = ((float)py + 0.5f - K[5]) / K[4] * z;
        X = ((float)px + 0.5f - K[2] - K[1] * Y / z) / K[0] * z;
      }
      xyz_cam[i * 3] = X; xyz_cam[i * 3 + 1] = Y; xyz_cam[i * 3 + 2] = z;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Fast depth refine, one CTA per ROI (gdrn_evaluator.py:539-561):
//   w = ||xyz|| * mask * (ren>0) * (sensor>0);  w /= sum(w);  sel = w > thresh*max(w)
//   dz = median(sensor[sel] - ren[sel]);  (y,x) = sum(w * (row, col));  ray = K^-1 (x,y,1) / z;  t += ray * dz
constexpr int DR_THREADS = 1024;

__global__ void __launch_bounds__(DR_THREADS)
depth_refine_kernel(const float* __restrict__ xyz, const float* __restrict__ mask, const float* __restrict__ sensor,
                    const float* __restrict__ ren, const float* __restrict__ Kc, float* __restrict__ trans, int hw,
                    float thresh, int mask_mode) {
  extern __shared__ float sm[];  // [npix] weights, then [npow2] diffs for the sort
  const int roi = blockIdx.x;
  const int npix = hw * hw;
  float* sw = sm;
  __shared__ double red_d[32];
  __shared__ float red_f[32];
  __shared__ int s_count;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* X = xyz + (long long)roi * 3 * npix;
  // get_out_mask (engine/engine_utils.py:313-333) folded in: mask_mode 0 = already normalised, 1 = raw L1 output ->
  // per-ROI min-max normalisation (m - min) / (max - min), 2 = raw logits -> sigmoid (BCE / RW_BCE / dice heads)
  float m_lo = 0.f, m_scale = 1.f;
  if (mask_mode == 1) {
    float mn = INFINITY, mx = -INFINITY;
    for (int i = tid; i < npix; i += DR_THREADS) {
      const float m = mask[(long long)roi * npix + i];
      mn = fminf(mn, m); mx = fmaxf(mx, m);
    }
    for (int o = 16; o > 0; o >>= 1) {
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    __shared__ float red_mm[2][32];
    if (lane == 0) { red_mm[0][warp] = mn; red_mm[1][warp] = mx; }
    __syncthreads();
    mn = INFINITY; mx = -INFINITY;
    for (int k = 0; k < DR_THREADS / 32; ++k) { mn = fminf(mn, red_mm[0][k]); mx = fmaxf(mx, red_mm[1][k]); }
    m_lo = mn; m_scale = mx - mn;    // applied as a DIVISION below, like torch's (mask - min) / (max - min)
  }
  // pass 1: weights, sum (double), max
  double lsum = 0.0;
  float lmax = 0.f;
  for (int i = tid; i < npix; i += DR_THREADS) {
    const float a = X[i], b = X[npix + i], c = X[2 * npix + i];
    float mv = mask[(long long)roi * npix + i];
    if (mask_mode == 1) mv = (mv - m_lo) / m_scale;
    else if (mask_mode == 2) mv = 1.0f / (1.0f + expf(-mv));
    float wv = sqrtf(a * a + b * b + c * c) * mv;
    const float rd = ren[(long long)roi * npix + i], sd = sensor[(long long)roi * npix + i];
    if (!(rd > 0.f) || !(sd > 0.f)) wv = 0.f;
    sw[i] = wv;
    lsum += (double)wv;
    lmax = fmaxf(lmax, wv);
  }
  for (int o = 16; o > 0; o >>= 1) {
    lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
    lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
  }
  if (lane == 0) { red_d[warp] = lsum; red_f[warp] = lmax; }
  if (tid == 0) s_count = 0;
  __syncthreads();
  double tsum = 0.0;
  float tmax = 0.f;
  for (int k = 0; k < DR_THREADS / 32; ++k) { tsum += red_d[k]; tmax = fmaxf(tmax, red_f[k]); }
  if (tsum == 0.0) return;  // reference: `continue` (pose unchanged)
  const float inv = (float)(1.0 / tsum);
  const float cut = (tmax * inv) * thresh;
  // pass 2: weighted centroid (double) and selection of depth differences
  int npow2 = 1;
  while (npow2 < npix) npow2 <<= 1;
  float* sdiff = sm + npix;
  for (int i = tid; i < npow2; i += DR_THREADS) sdiff[i] = INFINITY;
  __syncthreads();
  double ly = 0.0, lx = 0.0;
  for (int i = tid; i < npix; i += DR_THREADS) {
    const float wn = sw[i] * inv;
    ly += (double)wn * (double)(i / hw);
    lx += (double)wn * (double)(i % hw);
    if (wn > cut) {
      const int slot = atomicAdd(&s_count, 1);
      sdiff[slot] = sensor[(long long)roi * npix + i] - ren[(long long)roi * npix + i];
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    ly += __shfl_xor_sync(0xffffffffu, ly, o);
    lx += __shfl_xor_sync(0xffffffffu, lx, o);
  }
  __syncthreads();  // red_d reads above are complete
  __shared__ double red_y[32], red_x[32];
  if (lane == 0) { red_y[warp] = ly; red_x[warp] = lx; }
  __syncthreads();
  // bitonic sort of sdiff[0..npow2) ascending (INF padding sorts last)
  for (int k = 2; k <= npow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < npow2; i += DR_THREADS) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const float a = sdiff[i], b = sdiff[ixj];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { sdiff[i] = b; sdiff[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
  if (tid == 0) {
    const int cnt = s_count;
    if (cnt > 0) {
      // np.median: mean of the two middle elements for an even count
      const float med = (cnt & 1) ? sdiff[cnt >> 1] : 0.5f * (sdiff[(cnt >> 1) - 1] + sdiff[cnt >> 1]);
      double cy = 0.0, cx = 0.0;
      for (int k = 0; k < DR_THREADS / 32; ++k) { cy += red_y[k]; cx += red_x[k]; }
      // ray = inv(K) @ (x, y, 1), normalised by z (general 3x3 inverse in double)
      const float* K = Kc + roi * 9;
      const double a = K[0], b = K[1], c = K[2], d = K[3], e = K[4], f = K[5], g = K[6], h = K[7], i9 = K[8];
      const double det = a * (e * i9 - f * h) - b * (d * i9 - f * g) + c * (d * h - e * g);
      const double i00 = (e * i9 - f * h) / det, i01 = (c * h - b * i9) / det, i02 = (b * f - c * e) / det;
      const double i10 = (f * g - d * i9) / det, i11 = (a * i9 - c * g) / det, i12 = (c * d - a * f) / det;
      const double i20 = (d * h - e * g) / det, i21 = (b * g - a * h) / det, i22 = (a * e - b * d) / det;
      const double rx = i00 * cx + i01 * cy + i02, ry = i10 * cx + i11 * cy + i12, rz = i20 * cx + i21 * cy + i22;
      float* t = trans + roi * 3;
      t[0] = (float)((double)t[0] + rx / rz * (double)med);
      t[1] = (float)((double)t[1] + ry / rz * (double)med);
      t[2] = (float)((double)t[2] + 1.0 * (double)med);
    }
  }
}

}  // namespace

extern "C" size_t rast_scratch_bytes(int n, int H, int W) { return (size_t)n * H * W * 8; }

extern "C" int rast_render_depth(const float* verts, const int* faces, int V, int F, const float* poses,
                                 const float* Ks, int n, int H, int W, float znear, float zfar, int quantize_bits,
                                 float* depth, float* xyz_cam, unsigned long long* zbuf_scratch, void* stream) {
  GDRN_REQUIRE(verts && faces && poses && Ks && depth && zbuf_scratch, "rast: null argument");
  GDRN_REQUIRE(V > 0 && F > 0 && n > 0 && H > 0 && W > 0, "rast: empty input");
  GDRN_REQUIRE(quantize_bits == 0 || quantize_bits == 16 || quantize_bits == 24 || quantize_bits == 32,
               "rast: quantize_bits must be 0, 16, 24 or 32");
  GDRN_REQUIRE(znear > 0.f && zfar > znear, "rast: need 0 < znear < zfar");
  cudaStream_t st = (cudaStream_t)stream;
  const long long npx = (long long)n * H * W;
  int blocks = (int)((npx + 255) / 256 > 4096 ? 4096 : (npx + 255) / 256);
  rast_clear_kernel<<<blocks, 256, 0, st>>>(zbuf_scratch, npx);
  rast_tri_kernel<<<dim3((F + 255) / 256, n), 256, 0, st>>>(verts, faces, V, F, poses, Ks, H, W, znear, zfar,
                                                            zbuf_scratch);
  rast_resolve_kernel<<<blocks, 256, 0, st>>>(zbuf_scratch, Ks, n, H, W, znear, zfar, quantize_bits, depth, xyz_cam);
  GDRN_CHECK_CUDA(cudaGetLastError());
  gdrn_count_launch(3);
  return GDRN_OK;
}

// ---- mesh registry (process-wide, on the device that is current at the first upload) ------------------------------
namespace {
struct MeshRegistry {
  int device = -1;
  int count = 0;
  int max_faces = 0;
  MeshEntry host[RAST_MAX_MESHES];
  MeshEntry* dev_table = nullptr;
};
MeshRegistry g_meshes;
}  // namespace

extern "C" int rast_upload_mesh(const float* verts, int V, const int* faces, int F) {
  // verts / faces: host OR device pointers (cudaMemcpyDefault); the library keeps its own device copy.
  // Returns the mesh id (>= 0) or a negative GDRN_ERR_* code.
  if (!(verts && faces && V > 0 && F > 0)) { gdrn_set_last_error(__FILE__, __LINE__, "rast_upload_mesh: empty mesh"); return GDRN_ERR_INVALID; }
  int dev = 0;
  cudaGetDevice(&dev);
  if (g_meshes.device < 0) g_meshes.device = dev;
  if (g_meshes.device != dev) { gdrn_set_last_error(__FILE__, __LINE__, "rast_upload_mesh: the mesh registry lives on another device"); return GDRN_ERR_STATE; }
  if (g_meshes.count >= RAST_MAX_MESHES) { gdrn_set_last_error(__FILE__, __LINE__, "rast_upload_mesh: registry full (1024 meshes)"); return GDRN_ERR_STATE; }
  float* dv = nullptr;
  int* df = nullptr;
  cudaError_t e = cudaMalloc(&dv, (size_t)V * 12);
  if (e == cudaSuccess) e = cudaMalloc(&df, (size_t)F * 12);
  if (e == cudaSuccess) e = cudaMemcpy(dv, verts, (size_t)V * 12, cudaMemcpyDefault);
  if (e == cudaSuccess) e = cudaMemcpy(df, faces, (size_t)F * 12, cudaMemcpyDefault);
  if (e == cudaSuccess && !g_meshes.dev_table) e = cudaMalloc(&g_meshes.dev_table, sizeof(MeshEntry) * RAST_MAX_MESHES);
  if (e == cudaSuccess) {
    const int id = g_meshes.count;
    g_meshes.host[id] = MeshEntry{dv, df, V, F};
    e = cudaMemcpy(g_meshes.dev_table + id, &g_meshes.host[id], sizeof(MeshEntry), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
      g_meshes.count = id + 1;
      if (F > g_meshes.max_faces) g_meshes.max_faces = F;
      return id;
    }
  }
  (void)cudaGetLastError();
  cudaFree(dv); cudaFree(df);
  gdrn_set_last_error(__FILE__, __LINE__, cudaGetErrorString(e));
  return GDRN_ERR_CUDA;
}

extern "C" int rast_mesh_count(void) { return g_meshes.count; }

extern "C" void rast_free_meshes(void) {
  for (int i = 0; i < g_meshes.count; ++i) { cudaFree(const_cast<float*>(g_meshes.host[i].verts)); cudaFree(const_cast<int*>(g_meshes.host[i].faces)); }
  if (g_meshes.dev_table) cudaFree(g_meshes.dev_table);
  g_meshes = MeshRegistry();
}

extern "C" int rast_render_meshes(const int* mesh_ids, const float* poses, const float* Ks, int n, int H, int W,
                                  float znear, float zfar, int quantize_bits, float* depth, float* xyz_cam,
                                  unsigned long long* zbuf_scratch, void* stream) {
  GDRN_REQUIRE(mesh_ids && poses && Ks && depth && zbuf_scratch, "rast_render_meshes: null argument");
  GDRN_REQUIRE(n > 0 && H > 0 && W > 0, "rast_render_meshes: empty input");
  GDRN_REQUIRE(g_meshes.count > 0, "rast_render_meshes: no mesh uploaded (rast_upload_mesh)");
  GDRN_REQUIRE(quantize_bits == 0 || quantize_bits == 16 || quantize_bits == 24 || quantize_bits == 32,
               "rast: quantize_bits must be 0, 16, 24 or 32");
  GDRN_REQUIRE(znear > 0.f && zfar > znear, "rast: need 0 < znear < zfar");
  cudaStream_t st = (cudaStream_t)stream;
  const long long npx = (long long)n * H * W;
  int blocks = (int)((npx + 255) / 256 > 4096 ? 4096 : (npx + 255) / 256);
  rast_clear_kernel<<<blocks, 256, 0, st>>>(zbuf_scratch, npx);
  rast_tri_multi_kernel<<<dim3((g_meshes.max_faces + 255) / 256, n), 256, 0, st>>>(g_meshes.dev_table, mesh_ids, g_meshes.count, poses, Ks,
                                                                                 H, W, znear, zfar, zbuf_scratch);
  rast_resolve_kernel<<<blocks, 256, 0, st>>>(zbuf_scratch, Ks, n, H, W, znear, zfar, quantize_bits, depth, xyz_cam);
  GDRN_CHECK_CUDA(cudaGetLastError());
  gdrn_count_launch(3);
  return GDRN_OK;
}

extern "C" int gdrn_depth_refine_step(const float* xyz, const float* mask, const float* depth_sensor,
                                      const float* ren_depth, const float* K_crop, float* trans, int n, int hw,
                                      float thresh, void* stream) {
  return gdrn_depth_refine_step_ex(xyz, mask, 0, depth_sensor, ren_depth, K_crop, trans, n, hw, thresh, stream);
}

extern "C" int gdrn_depth_refine_step_ex(const float* xyz, const float* mask, int mask_mode, const float* depth_sensor,
                                         const float* ren_depth, const float* K_crop, float* trans, int n, int hw,
                                         float thresh, void* stream) {
  GDRN_REQUIRE(xyz && mask && depth_sensor && ren_depth && K_crop && trans, "depth_refine: null argument");
  GDRN_REQUIRE(mask_mode >= 0 && mask_mode <= 2, "depth_refine: mask_mode must be 0 (normalised), 1 (raw L1) or 2 (logits)");
  GDRN_REQUIRE(n > 0 && hw > 0 && hw <= 128, "depth_refine: need 0 < hw <= 128");
  int npix = hw * hw, npow2 = 1;
  while (npow2 < npix) npow2 <<= 1;
  size_t smem = (size_t)(npix + npow2) * 4;
  GDRN_OPT_IN_SMEM(depth_refine_kernel, 160 * 1024);
  depth_refine_kernel<<<n, DR_THREADS, smem, (cudaStream_t)stream>>>(xyz, mask, depth_sensor, ren_depth, K_crop, trans,
                                                                    hw, thresh, mask_mode);
  GDRN_CHECK_CUDA(cudaGetLastError());
  gdrn_count_launch(1);
  return GDRN_OK;
}
