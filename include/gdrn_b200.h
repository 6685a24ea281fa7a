/* libgdrn_b200.so -- C ABI of the B200-native GDRNPP per-ROI pose-inference hot path.
 *
 * Plain pointers and sizes only (no torch types).  Unless stated otherwise every pointer is a DEVICE
 * pointer, every call is asynchronous on the given CUDA stream (passed as void* = cudaStream_t), never
 * allocates, never synchronises, and returns 0 (GDRN_OK) or a negative error code; the message of the
 * last error is available from gdrn_last_error().  One process per GPU; the caller owns all buffers.
 *
 * Each entry point names the reference interface it replaces (paths relative to the reference tree
 * shanice-l/gdrnpp_bop2022 @ b80383bd).  The reference-side bindings are shown in INTEGRATION.md.
 */
#ifndef GDRN_B200_H_
#define GDRN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GDRN_OK 0
#define GDRN_ERR_INVALID (-1) /* bad argument / unsupported configuration */
#define GDRN_ERR_CUDA (-2)    /* CUDA runtime or driver error */
#define GDRN_ERR_STATE (-3)   /* model not fully loaded, workspace too small, ... */

const char* gdrn_last_error(void);
int gdrn_version(void);
/* number of CUDA kernels this library has launched in the calling process (monotonic) */
long long gdrn_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * Dense path: GDRN_DoubleMask.forward (eval, do_loss=False)
 *   replaces core/gdrn_modeling/models/GDRN_double_mask.py:66-214 and everything it calls:
 *   timm convnext features (core/utils/timm_utils.py:9-35), TopDownDoubleMaskXyzRegionHead.forward
 *   (models/heads/top_down_doublemask_xyz_region_head.py:177-211), ConvPnPNet.forward
 *   (models/heads/conv_pnp_net.py:120-183), rot6d_to_mat_batch (core/utils/rot_reps.py:34-55),
 *   pose_from_predictions_test (models/pose_from_pred_centroid_z.py:56-154) and
 *   allocentric_to_egocentric (core/utils/utils.py:31-88).
 * ------------------------------------------------------------------------------------------- */
typedef struct GdrnModel GdrnModel;

/* arch: "convnext_base" | "convnext_small" | "convnext_tiny"; num_classes: class-aware head size (21). */
int gdrn_model_create(GdrnModel** out, const char* arch, int num_classes, int max_batch);
/* precision: 0 = bf16 tensor-core operands, fp32 accumulate / residual stream (the throughput mode; what the
 *               reference's AMP test path does, engine_utils.py autocast);
 *            1 = split-bf16 ("bf16x3"): every GEMM operand is the pair (hi, lo) of bf16 values and every GEMM
 *               accumulates A_lo*W_hi + A_hi*W_lo + A_hi*W_hi in fp32; exact-erf GELU; fp32 Patch-PnP FC stack.
 *               Reproduces the reference's fp32 forward to ~1e-5 relative (the BASELINE.json parity bar).
 * gdrn_model_create() uses precision 0 unless the environment variable GDRN_PRECISION says otherwise. */
int gdrn_model_create_ex(GdrnModel** out, const char* arch, int num_classes, int max_batch, int precision);
int gdrn_model_precision(const GdrnModel* m);
void gdrn_model_destroy(GdrnModel* m);

/* Feed one fp32 tensor of a reference checkpoint (state_dict key names, e.g.
 * "backbone.stages_2.blocks.5.mlp.fc1.weight", "geo_head_net.features.0.weight", "pnp_net.fc_r.bias";
 * loader: core/utils/my_checkpoint.py:35-83).  The tensor is re-packed into kernel-native layouts on
 * `stream`; the source buffer may be freed after the stream reaches this point. Unknown keys return
 * GDRN_ERR_INVALID. */
int gdrn_model_load_tensor(GdrnModel* m, const char* key, const float* data, int64_t numel, void* stream);
/* number of tensors still missing (0 = ready) */
int gdrn_model_missing(const GdrnModel* m);

size_t gdrn_model_workspace_bytes(const GdrnModel* m, int batch);

typedef struct GdrnMaps { /* optional class-gathered maps (GDRN_double_mask.py:203-214), fp32 NCHW, or all NULL */
  float* mask;      /* [B,1,64,64] */
  float* full_mask; /* [B,1,64,64] */
  float* coor_x;    /* [B,1,64,64] */
  float* coor_y;    /* [B,1,64,64] */
  float* coor_z;    /* [B,1,64,64] */
  float* region;    /* [B,65,64,64] */
} GdrnMaps;

/* roi_img [B,3,256,256] fp32 NCHW in [0,1]; roi_classes [B] int64; roi_coord_2d [B,2,64,64]; roi_cams [B,3,3];
 * roi_centers [B,2]; roi_whs [B,2]; resize_ratios [B]; roi_extents [B,3]  ->  out_rot [B,3,3] (egocentric,
 * row-major), out_trans [B,3].  out_raw (optional, [B,9]) receives the Patch-PnP head output (rot6d, t_). */
int gdrn_model_forward(GdrnModel* m, const float* roi_img, const int64_t* roi_classes, const float* roi_coord_2d,
                       const float* roi_cams, const float* roi_centers, const float* roi_whs,
                       const float* resize_ratios, const float* roi_extents, int batch, float* out_rot,
                       float* out_trans, float* out_raw, const GdrnMaps* maps, void* workspace,
                       size_t workspace_bytes, void* stream);

/* Per-category device timing of the LAST forward (CUDA events on the launch stream): categories
 * 0 = tcgen05 GEMM kernel, 1 = depthwise-7x7+LayerNorm kernel, 2 = other CUDA-core kernels.
 * get_profile blocks until that forward has finished. */
int gdrn_model_set_profiling(GdrnModel* m, int enable);
int gdrn_model_get_profile(GdrnModel* m, float* ms_out, int* launches_out);

/* debugging / parity hooks: copy an internal activation (after the last forward on the same workspace)
 * as fp32 into dst.  name: "conv_feat" ([B,8,8,C3] NHWC) | "stage0".."stage3" | "head16"|"head32"|"head64"
 * ([B,h,w,256] NHWC, post GN+GELU) | "pnp_in" ([B,64,64,128]) . Returns element count or <0. */
int64_t gdrn_model_debug_read(GdrnModel* m, const char* name, int batch, float* dst, void* workspace, void* stream);

/* Plain GEMM through the same tcgen05 kernel, exported for tests and roofline measurement:
 * out[M,N] = epi(A[M,K] @ W[N,K]^T + bias); A, W bf16 row-major (K contiguous, lda/ldw = K).
 * epi: 0 = store (out bf16 or fp32 by out_f32), 1 = GELU -> bf16, 2 = resid + gamma*(.) -> fp32. */
int gdrn_gemm_bf16(const void* A, const void* W, const float* bias, const float* gamma, const float* resid,
                   void* out, int M, int N, int K, int epi, int out_f32, int block_n, void* stream);

/* The same for the split-bf16 ("bf16x3") precision mode, exported for tests and roofline measurement:
 * A [M, 2K] and W [N, 2K] hold [hi K | lo K] rows (hi = bf16(v), lo = bf16(v - hi)); the kernel accumulates
 * A_lo*W_hi + A_hi*W_lo + A_hi*W_hi in fp32.  epi: 0 = store fp32 [M,N]; 1 = exact-erf GELU -> split bf16 [M, 2N];
 * 2 = resid + gamma*(.) -> fp32 [M,N] (in place allowed).  block_n in {64, 128, 256}. */
int gdrn_gemm_x3(const void* A, const void* W, const float* bias, const float* gamma, const float* resid, void* out,
                 int M, int N, int K, int epi, int block_n, void* stream);

/* In-place residual form (epi 2) of gdrn_gemm_x3 with the balanced k-split schedule of the CTA-pair kernel allowed:
 * x [M,N] fp32 += gamma * (A @ W^T + bias).  `flags` = flag_words ZEROED device words (>= 16 per 256 x block_n tile); the
 * kernel leaves them zeroed.  Tiles whose K range is split between two CTA pairs are reduce-added in a fixed order, so the
 * result is run-to-run deterministic.  The schedule is used when GDRN_X3_KSPLIT=1 (default 0 = whole tiles: measured no
 * faster, the GEMMs are clock-limited by the power cap).  Exported for tests and measurement (the model forward
 * passes the same flag words for its fc2 GEMMs: timm ConvNeXtBlock.mlp.fc2 + layer scale + shortcut). */
int gdrn_gemm_x3_ksplit(const void* A, const void* W, const float* bias, const float* gamma, float* x, int M, int N, int K,
                        int block_n, unsigned* flags, int flag_words, void* stream);

/* Fused ConvNeXt MLP half-block of the split-bf16 mode (stage 0, C = 128), exported for tests and measurement:
 * x [M,C] fp32 += gamma * (W2 . gelu(W1 . a + b1) + b2) with a [M,2C], W1 [4C,2C], W2 [C,8C] in [hi | lo] bf16 rows; the
 * 4C-wide hidden activation stays on chip.  Replaces two gdrn_gemm_x3 calls (epi 1 then epi 2); same arithmetic.
 * Reference op: timm ConvNeXtBlock.mlp + layer scale + shortcut (models/GDRN_double_mask.py:102 via the backbone). */
int gdrn_mlp_fused_x3(const void* A, const void* W1, const float* b1, const void* W2, const float* b2, const float* gamma,
                      float* x, long long M, int C, void* stream);

/* ConvNeXt block front half exported for tests and roofline measurement: depthwise 7x7 (pad 3) + bias + LayerNorm(C)
 * on an NHWC fp32 tensor x [B,H,W,C] with tap-major weights w49c [49][C] -> bf16 [B*H*W, C] (split = 1: [hi C | lo C]).
 * variant: -1 default, 0 = one-tile-per-CTA cluster kernel, 1 = persistent two-warpgroup ping-pong kernel. */
int gdrn_dwconv_ln(const float* x, const float* w49c, const float* bias, const float* ln_w, const float* ln_b, void* out,
                   int B, int H, int W, int C, float eps, int split, int variant, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Farthest point sampling -- replaces core/csrc/fps/src/farthest_point_sampling.cpp:166-204
 * (cffi surface core/csrc/fps/src/ext.h:1-14, Python wrapper core/csrc/fps/fps_utils.py:6-21).
 * The two host-pointer symbols keep the reference's exact signature (blocking; they stage through the
 * GPU).  `farthest_point_sampling` (random start in the reference: srand(time(0)), :93-94) starts at
 * index `gdrn_fps_set_seed`-derived rand()%pn here; `_init_center` is deterministic and bit-exact.
 * ------------------------------------------------------------------------------------------- */
void farthest_point_sampling(float* pts, int* idxs, int pn, int sn);
void farthest_point_sampling_init_center(float* pts, int* idxs, int pn, int sn);
void gdrn_fps_set_seed(unsigned seed);
/* batched device version: pts [batch,pn,3] f32, idxs [batch,sn] i32.  start_idx: NULL -> bbox-centre
 * initialisation (init_center), else [batch] i32 explicit first index. */
int gdrn_fps_cuda(const float* pts, int* idxs, int pn, int sn, int batch, const int* start_idx, void* stream);

/* ---------------------------------------------------------------------------------------------
 * PVNet RANSAC voting -- replaces core/csrc/ransac_voting/src/ransac_voting.cpp:30-109 (pybind module
 * `ransac_voting`: generate_hypothesis, voting_for_hypothesis, *_vanishing_point) and the kernels
 * src/ransac_voting_kernel.cu:11-49,88-126,170-229,268-310.
 *   direct [tn,vn,2] f32, coords [tn,2] f32, idxs [hn,vn,2] i32, hypo [hn,vn,2|3] f32 (pre-zeroed by the
 *   caller like at::zeros), inliers [hn,vn,tn] u8 (in/out, pre-zeroed by the caller).
 * ------------------------------------------------------------------------------------------- */
int rv_generate_hypothesis(const float* direct, const float* coords, const int* idxs, float* hypo, int tn, int vn,
                           int hn, void* stream);
int rv_voting_for_hypothesis(const float* direct, const float* coords, const float* hypo, unsigned char* inliers,
                             int tn, int vn, int hn, float inlier_thresh, void* stream);
int rv_generate_hypothesis_vanishing_point(const float* direct, const float* coords, const int* idxs, float* hypo,
                                           int tn, int vn, int hn, void* stream);
int rv_voting_for_hypothesis_vanishing_point(const float* direct, const float* coords, const float* hypo,
                                             unsigned char* inliers, int tn, int vn, int hn, float inlier_thresh,
                                             void* stream);
/* fused round: votes without materialising the [hn,vn,tn] mask; counts [hn,vn] i32 (overwritten). */
int rv_vote_count(const float* direct, const float* coords, const float* hypo, int* counts, int tn, int vn, int hn,
                  float inlier_thresh, int vanishing_point, void* stream);

/* The whole RANSAC voting layer for a batch of images on the device (SURVEY.md 8b "rv_ransac_round"): replaces the
 * Python driver ransac_voting_layer / ransac_voting_layer_v3 (core/csrc/ransac_voting/ransac_voting_gpu.py:7-104,
 * 123-218).  mask [b,h,w] f32 (non-zero = foreground), vertex [b,h,w,vn,2] f32 -> win_pts [b,vn,2]: ordered
 * compaction of the foreground (tn stays on the device; images with fewer than min_num pixels give zeros, more than
 * max_num are randomly subsampled), hn hypotheses per keypoint from pixel pairs idxs [b,hn,vn,2] i32 (optional;
 * taken modulo tn; NULL = counter-based RNG(seed)), fused vote + count, per-keypoint winner, inlier set of the winner
 * and least-squares refit of its lines.  No host synchronisation.  Optional outputs: hypo [b,hn,vn,2], counts
 * [b,hn,vn] i32, tn [b] i32, inliers [b,vn,h*w] u8 (compacted pixel order, first tn[b] entries valid). */
size_t rv_layer_workspace_bytes(int b, int h, int w, int vn, int hn);
int rv_ransac_voting_layer(const float* mask, const float* vertex, int b, int h, int w, int vn, int hn,
                           float inlier_thresh, int min_num, int max_num, unsigned seed, const int* idxs, float* win_pts,
                           float* hypo, int* counts, int* tn, unsigned char* inliers, void* workspace,
                           size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Chamfer / NN distance -- replaces core/csrc/torch_nndistance/src/nnd_cuda.cpp:37-84
 * (module torch_nndistance_aten: nnd_forward_cuda / nnd_backward_cuda; kernels nnd_cuda_kernel.cu:8-183).
 *   xyz1 [b,n,3], xyz2 [b,m,3] f32 -> dist1 [b,n], dist2 [b,m] f32, idx1 [b,n], idx2 [b,m] i32.
 *   backward: grad buffers are ACCUMULATED into (caller zeroes them, as torch_nndistance.py:52-53 does).
 * Returns 1 on success / 0 on failure like the reference launchers.
 * ------------------------------------------------------------------------------------------- */
int nnd_forward_cuda(const float* xyz1, const float* xyz2, float* dist1, float* dist2, int* idx1, int* idx2, int b,
                     int n, int m, void* stream);
int nnd_backward_cuda(const float* xyz1, const float* xyz2, float* gradxyz1, float* gradxyz2, const float* graddist1,
                      const float* graddist2, const int* idx1, const int* idx2, int b, int n, int m, void* stream);

/* ---------------------------------------------------------------------------------------------
 * DeepIM depth-reprojection flow -- replaces core/csrc/flow/src/flow_cuda.cpp:30-42 (module flow_cuda,
 * kernel flow_cuda_kernel.cu:26-65). depth_src/tgt [B,1,H,W], KT [B,3,4], Kinv [B,3,3] ->
 * flow [B,2,H,W] (ch0 = dh, ch1 = dw), valid [B,1,H,W]; fp32.
 * ------------------------------------------------------------------------------------------- */
int flow_forward_cuda(const float* depth_src, const float* depth_tgt, const float* KT, const float* Kinv, float* flow,
                      float* valid, int batch, int height, int width, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Uncertainty PnP -- replaces core/csrc/uncertainty_pnp/src/uncertainty_pnp.cpp:61-92 (cffi surface
 * src/ext.h; Ceres AutoDiff + DENSE_SCHUR LM).  Host f64 pointers, blocking, same signature.
 * upnp_batched: device f64, `n_problems` independent problems of `pn` points each.
 * ------------------------------------------------------------------------------------------- */
void uncertainty_pnp(double* pts2d, double* pts3d, double* wgt2d, double* K, double* init_rt, double* result_rt,
                     int pn);
int upnp_batched(const double* pts2d, const double* pts3d, const double* wgt2d, const double* K,
                 const double* init_rt, double* result_rt, int pn, int n_problems, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Depth rasteriser -- replaces the GL renderers used for depth: lib/render_vispy/renderer.py:126-182,
 * 363-407 (Renderer.set_cam/draw_model/finish -> depth) and lib/egl_renderer/egl_renderer_v3.py:838-1228
 * (EGLRenderer.render(..., pc_cam_tensor=) -> pc_cam[...,2]); C++ side lib/egl_renderer/cpp/
 * egl_renderer.cpp:262-310.
 *   verts [V,3] f32 (model space, metres), faces [F,3] i32; poses [n,3,4] f32 (R|t, OpenCV camera),
 *   Ks [n,3,3] f32 -> depth [n,H,W] f32 (0 = background), optional xyz_cam [n,H,W,3].
 *   quantize_bits: 0 = float depth of the interpolated camera-space z (EGL path), 24/16 = emulate the
 *   fixed-point z-buffer decode of the vispy path (renderer.py:176-182) with znear/zfar.
 * ------------------------------------------------------------------------------------------- */
int rast_render_depth(const float* verts, const int* faces, int V, int F, const float* poses, const float* Ks,
                      int n, int H, int W, float znear, float zfar, int quantize_bits, float* depth,
                      float* xyz_cam, unsigned long long* zbuf_scratch, void* stream);
size_t rast_scratch_bytes(int n, int H, int W);
/* Mesh registry (rast_upload_mesh + rast_render_depth(mesh_ids, ...)): meshes are uploaded once
 * (host or device pointers; the library keeps a device copy) and a batch of ROIs renders one mesh EACH, selected by
 * mesh_ids [n] i32 (device) -- the multi-object form of the reference's per-object draw_model loop
 * (engine/gdrn_evaluator.py:520-526) without a host loop or a sync.  rast_upload_mesh returns the id (>= 0) or a
 * negative error code. */
int rast_upload_mesh(const float* verts, int V, const int* faces, int F);
int rast_mesh_count(void);
void rast_free_meshes(void);
int rast_render_meshes(const int* mesh_ids, const float* poses, const float* Ks, int n, int H, int W, float znear,
                       float zfar, int quantize_bits, float* depth, float* xyz_cam, unsigned long long* zbuf_scratch,
                       void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fast depth refinement -- replaces GDRN_Evaluator.process_depth_refine
 * (core/gdrn_modeling/engine/gdrn_evaluator.py:521-561) == GdrnPredictor.process_depth_refine
 * (core/gdrn_modeling/demo/predictor_gdrn.py:239-286), one iteration after a render:
 *   xyz [n,3,64,64], mask [n,64,64], depth_sensor [n,64,64], ren_depth [n,64,64], K_crop [n,3,3],
 *   trans [n,3] in/out; thresh = DEPTH_REFINE_THRESHOLD (0.8).
 * ------------------------------------------------------------------------------------------- */
int gdrn_depth_refine_step(const float* xyz, const float* mask, const float* depth_sensor, const float* ren_depth,
                           const float* K_crop, float* trans, int n, int hw, float thresh, void* stream);
/* same with get_out_mask (engine/engine_utils.py:313-333) folded into the kernel: mask_mode 0 = `mask` is already
 * normalised, 1 = raw L1 mask output -> per-ROI min-max normalisation, 2 = raw logits -> sigmoid. */
int gdrn_depth_refine_step_ex(const float* xyz, const float* mask, int mask_mode, const float* depth_sensor,
                              const float* ren_depth, const float* K_crop, float* trans, int n, int hw, float thresh,
                              void* stream);

/* ---------------------------------------------------------------------------------------------
 * ROI crop + resize -- replaces crop_resize_by_warp_affine = cv2.warpAffine
 * (core/utils/data_utils.py:115-133) as called once per ROI by GdrnPredictor.preprocessing
 * (core/gdrn_modeling/demo/predictor_gdrn.py:417-438) and the test data loader
 * (core/gdrn_modeling/datasets/data_loader.py:758-797).  OpenCV's warpAffine arithmetic is reproduced exactly
 * (bit-exact for uint8 and nearest, bit-exact with the scalar float path for float bilinear; zero border).
 *   M [n,6] f64 (device): forward (source -> crop) 2x3 matrices, row-major, as get_affine_transform returns them.
 *   gdrn_crop_resize_u8 : image [H,W,C] u8 (BGR as cv2.imread, device), INTER_LINEAR, followed by
 *                         normalize_image ((v - pixel_mean[c]) / pixel_std[c] in f64 -> f32; host arrays [C])
 *                         -> out [n,C,out_h,out_w] f32 (the roi_img the model consumes).
 *   gdrn_crop_resize_f32: src [H,W,C] f32 (device), INTER_LINEAR (nearest = 0; roi_coord_2d) or INTER_NEAREST
 *                         (nearest = 1; roi_depth) -> out [n,C,out_h,out_w] f32.
 * ------------------------------------------------------------------------------------------- */
int gdrn_crop_resize_u8(const uint8_t* image, int H, int W, int C, const double* M, int n, int out_h, int out_w,
                        const double* pixel_mean, const double* pixel_std, float* out, void* stream);
int gdrn_crop_resize_f32(const float* src, int H, int W, int C, const double* M, int n, int out_h, int out_w,
                         int nearest, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Batched RANSAC-PnP -- replaces the per-ROI host loop get_pnp_ransac_pose
 * (core/gdrn_modeling/engine/gdrn_evaluator.py:1122-1221) + misc.pnp_v2 (lib/pysixd/misc.py:153-208) =
 * cv2.solvePnPRansac(EPNP, reprojectionError 3, 100 iterations), the reference predictor's default TEST.USE_PNP path.
 * One CTA per ROI: correspondence selection from the maps, `iters` P3P hypotheses, inlier counts, LM refit on the
 * inliers (csrc/pnp_ransac.cu).  poses [n,3,4] (R|t, -100 everywhere when fewer than 4 correspondences / no model),
 * n_inliers [n] (optional), inlier_mask [n, hw*hw | npts] u8 over the INPUT pixels / points (optional).
 * idxs (optional) [n, iters, 4] i32: the sampled correspondences of every hypothesis (taken modulo the number of
 * selected correspondences) instead of the built-in counter-based RNG(seed).
 *   _maps  : coor_x/y/z, mask [n, hw, hw] raw network outputs, roi_coord_2d [n,2,hw,hw], im_hw [n,2] = (im_H, im_W),
 *            extents [n,3], Ks [n,3,3]; mask_thr = GEO_HEAD.MASK_THR_TEST (0.5).
 *   _points: explicit correspondences pts3d [n,npts,3], pts2d [n,npts,2] (all used), npts <= 4096.
 * ------------------------------------------------------------------------------------------- */
int gdrn_pnp_ransac_maps(const float* coor_x, const float* coor_y, const float* coor_z, const float* mask,
                         const float* roi_coord_2d, const float* im_hw, const float* extents, const float* Ks,
                         const int* idxs, int n, int hw, int iters, float mask_thr, float reproj_thr, unsigned seed,
                         float* poses, int* n_inliers, unsigned char* inlier_mask, void* stream);
int gdrn_pnp_ransac_points(const float* pts3d, const float* pts2d, const float* Ks, const int* idxs, int n, int npts,
                           int iters, float reproj_thr, unsigned seed, float* poses, int* n_inliers,
                           unsigned char* inlier_mask, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Online training targets -- after rast_render_meshes, replaces the torch ops of
 * core/gdrn_modeling/engine/engine_utils.py:131-187: misc.calc_xyz_bp_batch (lib/pysixd/misc.py:412-457),
 * roi_mask_obj, xyz_to_region_batch (core/utils/data_utils.py:283-301) and the xyz normalisation, in one kernel.
 *   depth [n,H,W], R [n,3,3] (ego rotation), T [n,3], K [n,3,3] (zoomed intrinsics), fps_points [n,F,3], extents [n,3]
 *   -> roi_xyz [n,3,H,W] (xyz / extent + 0.5), xyz_raw [n,H,W,3] (object-space xyz), mask_obj [n,H,W] f32,
 *      region [n,H,W] i64 (1..F, 0 = background).  Any output may be NULL.
 * ------------------------------------------------------------------------------------------- */
int gdrn_xyz_region_targets(const float* depth, const float* R, const float* T, const float* K, const float* fps_points,
                            const float* extents, int n, int H, int W, int F, float* roi_xyz, float* xyz_raw,
                            float* mask_obj, long long* region, void* stream);

/* ---------------------------------------------------------------------------------------------
 * YOLOX head post-processing -- replaces YOLOXHead.decode_outputs
 * (det/yolox/models/yolo_head.py:239-255) and postprocess (det/yolox/utils/boxes.py:34-80: cxcywh -> xyxy, class
 * max, obj * cls_conf >= conf_thre, torchvision batched_nms / nms) for a batch of images, no host synchronisation.
 *   preds [B, A, 5 + num_classes] f32 (obj / class scores already sigmoid-ed like the head emits them);
 *   n_levels > 0: raw head outputs, decoded here with hw [n_levels,2] i32 (rows, cols per level) and strides [n_levels]
 *   i32 (both device arrays); n_levels = 0: preds are already decoded (cx, cy, w, h).
 *   dets [B, max_out, 7] = (x1, y1, x2, y2, obj_conf, class_conf, class) in descending score order, n_det [B] i32.
 * ------------------------------------------------------------------------------------------- */
size_t yolox_postprocess_workspace_bytes(int B, int A);
int yolox_postprocess(const float* preds, int B, int A, int num_classes, const int* hw, const int* strides, int n_levels,
                      float conf_thre, float nms_thre, int class_agnostic, int max_out, float* dets, int* n_det,
                      void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GDRN_B200_H_ */
