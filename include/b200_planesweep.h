/* C ABI of libb200planesweep.so -- the sm_100a plane-sweep hot path of implicit-depth.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to float32 data unless stated otherwise; buffers are
 *     dense row-major in the shape given; nothing is allocated or freed by the library;
 *   - `stream` is a cudaStream_t passed as void*; calls only enqueue work on it (no host
 *     synchronisation), so they can sit between the CUDA events of test_bd.py:196-212;
 *   - return value 0 = ok; negative = error (-1 bad argument, -2 CUDA error), message via
 *     b200_last_error().  Nothing aborts.
 *   - "pixel-major" features come in two gather layouts: layout 0 = texel records [n_images, h*w, 16] (one 64-byte
 *     record per texel; consumed by b200_cv_dot), layout 1 = quarter-planar [n_images, 4, h*w, 4] (channels
 *     4q..4q+3 of texel t at ((img*4 + q)*h*w + t)*4; consumed by b200_fv_mlp_*).
 *
 * Citations are file:line in the reference repository (nianticlabs/implicit-depth).
 */
#ifndef B200_PLANESWEEP_H
#define B200_PLANESWEEP_H

#ifdef __cplusplus
extern "C" {
#endif

int b200_abi_version(void);           /* 2: CTA caps are per-call arguments (`max_ctas`), no process-global state */
/* A dedicated non-blocking CUDA stream for the host-side mirror (Plan scheduler pools, encoder / copy streams): streams
 * that must be distinct within one CUDA-graph capture must not come from PyTorch's 32-entry stream pool, which aliases
 * once a process has created more.  No reference counterpart (the reference runs on one stream). */
int b200_stream_create(int priority, void** stream_out);
int b200_stream_destroy(void* stream);
const char* b200_last_error(void);    /* message of the last failing call on this thread */
const char* b200_source_digest(void); /* sha256 of the sources the library was built from (checked by the loader) */
/* `max_ctas` (b200_conv_desc, b200_fv_mlp_tc, b200_stem_conv7_tc): upper bound on the CTAs of that persistent launch,
 * 0 = all SMs -- keeps SMs free for work the caller runs concurrently on another stream. */

/* Per-batch set-up, replaces the tensor algebra at the top of every manager's
 * build_cost_volume: P = (K_src @ T_src<-cur)[:3] (utils/geometry_utils.py:82-84) folded with
 * invK_cur (:60), the source-camera centres (modules/cost_volume.py:1088), pose_distance
 * (utils/geometry_utils.py:183-195), the log-spaced depth planes
 * (modules/cost_volume.py:98-132) and, when W1/b1 are given, the part of the first MLP layer
 * that is constant over (pixel, plane): mask (==1) and pose channels (:684, :690-692).
 *   src_Ks, src_extrinsics, src_poses [B,K,4,4]; cur_invK [B,4,4];
 *   min_depth, max_depth: device scalars, or [B] values when range_per_frame != 0 (the reference broadcasts its
 *   [1,1,1,1] or [B,1,1,1] tensors, modules/cost_volume.py:117-126); used when planes_in == NULL; planes_in [B,D] or NULL;
 *   W1 [128, 26K+20] and b1 [128] in the reference's channel order, or NULL;
 *   out: cams [B,K,32], planes [B,D], bias_eff [B,128] (only with W1). */
int b200_volume_prepare(const float* src_Ks, const float* src_extrinsics, const float* src_poses,
                        const float* cur_invK, const float* min_depth, const float* max_depth,
                        const float* planes_in, const float* W1, const float* b1, float* cams, float* planes,
                        float* bias_eff, int B, int K, int D, int C, int range_per_frame, void* stream);

/* [n_img, C=16, HW] planes (image stride / channel stride in elements) -> pixel-major, layout 0 | 1 (see top). */
int b200_feats_to_pixel_major(const float* in, float* out, int n_img, int C, int HW, long long img_stride,
                              long long ch_stride, int layout, void* stream);

/* argmax over planes (first maximum, modules/cost_volume.py:352-356) + gather of the plane depth.
 *   vol [B,D,N]; planes [B,D]; out: lowest [B,N] float, best_idx [B,N] int32 or NULL. */
int b200_volume_argmax(const float* vol, const float* planes, float* lowest, int* best_idx, int B, int D, int N,
                       void* stream);

/* Fused dot-product plane sweep = CostVolumeManager.build_cost_volume + forward
 * (modules/cost_volume.py:221-358) / EfficientCostVolumeManager (:1245-1304):
 * warp (grid_sample bilinear/zeros/align_corners=False, :192-198), dot with the current
 * features, sum over views (:302-311), argmax + depth gather (:352-356).
 *   cur [B,N,16], src [B,K,N,16] pixel-major layout 0; cams/planes from b200_volume_prepare;
 *   out: cost [B,D,h,w]; lowest [B,h,w] or NULL; best_idx [B,h,w] int32 or NULL. */
int b200_cv_dot(const float* cur, const float* src, const float* cams, const float* planes, float* cost,
                float* lowest, int* best_idx, int B, int K, int C, int h, int w, int D, void* stream);
/* Same contract, same results bit for bit, shared-memory-band version (csrc/cv_dot_band.cu): per (8x8 pixel tile, 4
 * planes, view) ONE TMA box of 20x20 texel records is staged in shared memory (zero fill = grid_sample's zeros padding)
 * and the four bilinear taps are LDS.128; iterations whose bounding box does not fit gather from global memory. */
int b200_cv_dot_band(const float* cur, const float* src, const float* cams, const float* planes, float* cost,
                     float* lowest, int* best_idx, int B, int K, int C, int h, int w, int D, void* stream);

/* Fused metadata-MLP plane sweep = FeatureVolumeManager.build_cost_volume
 * (modules/cost_volume.py:437-706) / FastFeatureVolumeManager (:938-1146) with the MLP of
 * modules/networks.py:218-233, strict-fp32 CUDA-core version.
 *   W1p [KP,128]: first layer, transposed, columns permuted to the kernel's view-major channel
 *   order (implicit_depth_b200/cost_volume.py: channel_permutation), zero rows up to KP;
 *   W2t [128,128] = W2^T; b2 [128]; w3 [128]; b3 [1]; bias_eff [B,128] from b200_volume_prepare;
 *   out: vol [B,D,h,w]; mask_out [B,h,w] uint8 (overall_mask_bhw, last plane only, :603-615) or NULL. */
int b200_fv_mlp_simt(const float* cur, const float* src, const float* cams, const float* cur_invK,
                     const float* planes, const float* bias_eff, const float* W1p, const float* W2t,
                     const float* b2, const float* w3, const float* b3, float* vol, unsigned char* mask_out, int B,
                     int K, int C, int h, int w, int D, int KP, void* stream);

/* Same contract as b200_fv_mlp_simt on the tcgen05 tensor cores: the input rows are written by their
 * owning threads straight into tensor memory as split-bf16 A operands, both weight matrices stay in
 * shared memory, three bf16 MMAs (hi*hi + hi*lo + lo*hi) per k-step give fp32-grade results.
 *   wimage: b200_fv_tc_wimage_bytes(K) bytes = pre-swizzled split-bf16 image of the permuted, zero-padded
 *   W1 [128, 64*ceil((22K+20)/64)] (hi tiles, lo tiles) followed by W2 [128,128] (hi tiles, lo tiles);
 *   16 KB tiles of [128 x 64] bf16 in the 128-byte-swizzled K-major UMMA layout, 16-byte aligned. */
int b200_fv_mlp_tc(const float* cur, const float* src, const float* cams, const float* cur_invK,
                   const float* planes, const float* bias_eff, const void* wimage, const float* b2,
                   const float* w3, const float* b3, float* vol, unsigned char* mask_out, int B, int K, int C,
                   int h, int w, int D, int max_ctas, void* stream);
int b200_fv_tc_wimage_bytes(int K);
/* K-dimension layout of that image: out[0] = views built by role A (the rest by role B), out[1] / out[2] =
 * 32-channel halves of role A / B, out[3] = 64-channel chunks; returns the image size in bytes
 * (implicit_depth_b200/cost_volume.py: tc_channel_layout turns it into the column permutation of W1). */
int b200_fv_tc_layout(int K, int* out);

/* ---- tensor-core convolution (csrc/conv_tc.cu) ------------------------------------------------
 * Implicit-GEMM conv over NHWC split-bf16 activations (two bf16 planes hi/lo with x = hi + lo).
 * Replaces nn.Conv2d + bias + activation + residual + torch.cat of modules/layers.py:8-95 (BasicBlock),
 * modules/networks.py:186-215 (CVEncoder), :20-84 (BDDecoderPP), modules/networks_fast.py:10-99.
 * A conv is described once (b200_conv_create encodes the TMA tensor maps) and then launched any number
 * of times (b200_conv_run only enqueues a kernel, so it is CUDA-graph capturable). */
typedef struct {
  const void* in_hi; /* NHWC bf16 [B,H,W,C], 16-byte aligned, C % 8 == 0 */
  const void* in_lo;
  int H, W, C, ksize /*1|3*/, stride /*1|2*/, pad /* top / left */;
  int pad_hi;        /* bottom / right padding: == pad for torch-style symmetric padding; (pad, pad_hi) = (0, 1) is the
                        TF "SAME" padding of a stride-2 3x3 conv on an even-sized map (timm tf_efficientnetv2_s,
                        reference call site experiment_modules/bd_model.py:46-51) */
} b200_conv_seg;
typedef struct {
  b200_conv_seg seg[4]; /* K-segments: concatenated inputs and/or the shortcut conv of a BasicBlock */
  int nseg;
  const void* wimage;   /* packed weights, b200_conv_wimage_bytes() bytes: [n-tile][chunk][hi NT rows | lo NT rows],
                           chunks ordered (segment, tap row-major, channel block), zero-padded channel tails;
                           b200_conv_uses_halo(desc) == 1: 32-channel blocks, 64-byte-swizzled K-major tiles,
                           == 0: 64-channel blocks, 128-byte swizzle */
  const float* bias;    /* [Cout] or NULL */
  const void* res_hi;   /* optional identity residual NHWC [B,OH,OW,Cout] (layers.py:92) */
  const void* res_lo;
  void* out_hi;         /* NHWC [B,OH,OW,Cout] bf16 planes (both or neither) */
  void* out_lo;
  float* out_f32;       /* optional NHWC fp32 copy of the output */
  int B, OH, OW, Cout /* %16==0; >128 => %128==0 */, act /*0 none,1 leaky(slope),2 ELU,3 ReLU*/;
  float slope;
  int max_ctas;         /* CTA cap of this persistent launch, 0 = all SMs */
  int tile_hint;        /* 0 = automatic; 64-channel halo convs: 1 = M=128 items, two CTAs per SM; 2 = M=256 items */
} b200_conv_desc;
int b200_conv_ntile(int Cout);
/* N tile of this particular conv (64 instead of 128 for ring-kernel convs with few M tiles); the weight image must be
 * packed for it. */
int b200_conv_ntile_for(const b200_conv_desc* d);
/* which kernel (and weight-image layout) a conv gets: 1 = halo-patch kernel (all segments stride 1, no fp32 copy,
 * Cout % 64 == 0 -- or Cout == 16 with 3x3 segments and no residual: the replicate-pad head of the matching encoder,
 * modules/networks.py:279-282), 0 = per-tap kernel.  Depends only on the geometry fields of the descriptor. */
int b200_conv_uses_halo(const b200_conv_desc* desc);
long long b200_conv_wimage_bytes(const int* seg_C, const int* seg_ksize, int nseg, int Cout, int halo);
int b200_conv_create(const b200_conv_desc* desc, void** plan_out);
int b200_conv_run(void* plan, void* stream);
int b200_conv_destroy(void* plan);

/* Stem of the image-prior encoder: Conv2d(3, Cout, 3, stride 2) + folded BatchNorm + SiLU straight from the fp32 image
 * (replaces timm `conv_stem` + `bn1` + SiLU of `tf_efficientnetv2_s`, reference call site
 * experiment_modules/bd_model.py:46-51; torchvision `features[0]` alike).  img [B,3,H,W] fp32 with element strides
 * sB,sC,sH,sW; w [27][Cp] (k = (c*3+dy)*3+dx, output channels zero-padded to Cp % 8 == 0), bias [Cp]; pad_lo = 0 (TF
 * "SAME" at stride 2 on even sizes: padding (0,1)) or 1 (symmetric); out: split-bf16 NHWC [B,OH,OW,Cp]. */
int b200_stem3x3_s2_silu(const float* img, const float* w, const float* bias, void* out_hi, void* out_lo, int B, int H,
                         int W, int Cp, int pad_lo, long long sB, long long sC, long long sH, long long sW,
                         void* stream);

/* ---- layout / elementwise kernels of the conv path (csrc/elementwise.cu) -------------------- */
/* fp32 [B,C,H,W] with element strides (NCHW or channels_last) -> NHWC split-bf16 planes. */
int b200_f32_to_split(const float* in, void* hi, void* lo, int B, int C, int H, int W, long long sB, long long sC,
                      long long sH, long long sW, void* stream);
/* NHWC split-bf16 -> dense fp32 NCHW (the layout the reference's modules return). */
int b200_split_to_nchw(const void* hi, const void* lo, float* out, int B, int C, int H, int W, void* stream);
/* x2 upsampling, mode 0 bilinear align_corners=False (utils/generic_utils.py:94-103), 1 nearest
 * (modules/networks_fast.py:42); out is [B,2H,2W,C]. */
int b200_upsample2x(const void* in_hi, const void* in_lo, void* out_hi, void* out_lo, int B, int H, int W, int C,
                    int mode, void* stream);
/* nn.InstanceNorm2d (affine=False, biased variance; modules/networks.py:277,283) + optional LeakyReLU;
 * out = split NHWC with a replicated border of `pad` pixels, and/or fp32 pixel-major in layout `f32_layout` (pad 0).
 * Workspaces sized by b200_instance_norm_ws_bytes. Deterministic, batch-invariant reduction. */
int b200_instance_norm(const void* in_hi, const void* in_lo, double* partial_ws, float* stats_ws, void* out_hi,
                       void* out_lo, float* out_f32, int B, int H, int W, int C, int pad, int act, float slope,
                       float eps, int f32_layout, void* stream);
int b200_instance_norm_ws_bytes(int B, int C, long long* partial_bytes, long long* stats_bytes);
/* ---- EfficientNetV2 image-prior encoder (reference call site bd_model.py:46-51: timm tf_efficientnetv2_s,
 * features_only; torchvision efficientnet_v2_s layout).  Its dense convolutions run on b200_conv_* (act 4 = SiLU);
 * these are the memory-bound parts of the MBConv blocks, NHWC split-bf16 in and out. ---- */
/* depthwise 3x3 (stride 1|2) + bias (BatchNorm folded) + SiLU; wt [9][C] tap-major fp32.  Padding (pad_lo, 1):
 * pad_lo = 1 is torch's symmetric pad 1, pad_lo = 0 the TF "SAME" padding of a stride-2 conv on an even-sized map. */
int b200_dwconv3x3_silu(const void* in_hi, const void* in_lo, const float* wt, const float* bias, void* out_hi,
                        void* out_lo, int B, int H, int W, int C, int stride, int pad_lo, void* stream);
/* SqueezeExcitation: mean over H*W -> fc1 w1 [S][C] + SiLU -> fc2 (transposed: w2t [S][C]) + sigmoid -> x * scale
 * (in place when out == in).  mean_ws / scale_ws: [B,C] fp32.  Deterministic reduction. */
int b200_squeeze_excite(const void* in_hi, const void* in_lo, const float* w1, const float* b1, const float* w2t,
                        const float* b2, float* mean_ws, float* scale_ws, void* out_hi, void* out_lo, int B, int HW,
                        int C, int S, void* stream);
/* The two calls above as one chain with the squeeze fused into the depthwise conv and the excitation spread over
 * many blocks (what the encoder plan uses): out = dw(x) * sigmoid(fc2(silu(fc1(mean_hw(dw(x)))))), torchvision
 * MBConv.block[1:3].  Workspaces (fp32): partial_ws [B][ceil(OH*OW / b200_mbconv_pool_block())][C], s1_ws [B][S],
 * scale_ws [B][C].  Fixed-order reductions: deterministic and batch-invariant. */
int b200_mbconv_dw_se(const void* in_hi, const void* in_lo, const float* wt, const float* bias, const float* w1,
                      const float* b1, const float* w2t, const float* b2, float* partial_ws, float* s1_ws,
                      float* scale_ws, void* out_hi, void* out_lo, int B, int H, int W, int C, int stride, int S,
                      int pad_lo /* like b200_dwconv3x3_silu */, void* stream);
int b200_mbconv_pool_block(void);
/* out = a + b on split activations of n elements (n % 8 == 0). */
int b200_split_add(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo, void* out_hi, void* out_lo,
                   long long n, void* stream);

/* 1x1 convolution to one channel (+ exp): the last layer of the regression heads of DepthDecoderPP
 * (modules/networks.py:160-163) / SkipDecoderRegression (networks_fast.py:106-112) and the exp() of
 * DepthModel.forward (depth_model.py:426-435).  in: split NHWC [n_pix, C]; out_log, out_exp (or NULL): [n_pix]. */
int b200_channel_dot_exp(const void* in_hi, const void* in_lo, const float* w, const float* bias, float* out_log,
                         float* out_exp, long long n_pix, int C, void* stream);

/* Output side of the evaluation scripts: sigmoid_custom (modules/layers.py:138-139) + F.interpolate to the
 * ground-truth size, bilinear (align_corners=False) or nearest (test_bd.py:225-243, inference/inference.py:159-162).
 * in [N,h,w] -> out [N,H,W] fp32; apply_sigmoid = 0 resizes only (rendered_depth / search_depths, test_bd.py:245-271). */
int b200_sigmoid_resize(const float* in, float* out, int N, int h, int w, int H, int W, float multiplier, int nearest,
                        int apply_sigmoid, void* stream);

/* Matching-encoder stem conv 7x7/2 (3->64, BatchNorm folded) + ReLU (modules/networks.py:264-266):
 * img fp32 NCHW [n,3,H,W]; wt [147,64] tap-major (c,dy,dx); out NHWC split [n,H/2,W/2,64]. */
int b200_stem_conv7(const float* img, const float* wt, const float* bias, void* out_hi, void* out_lo, int n_img,
                    int H, int W, void* stream);
/* Same layer on the tensor cores (csrc/stem_tc.cu): implicit GEMM with K = (c*7 + dy)*8 + dx built row by row into
 * tensor memory.  wimage: 48 KB = 3 K-chunks x [W_hi 64x64 | W_lo 64x64] bf16 in the 128-byte-swizzled K-major
 * layout (BatchNorm folded, k >= 168 and dx == 7 zero). */
int b200_stem_conv7_tc(const float* img, const void* wimage, const float* bias, void* out_hi, void* out_lo,
                       int n_img, int H, int W, int max_ctas, void* stream);
/* MaxPool2d(2, stride 1) + BlurPool(4x4 binomial, stride 2, reflect pad) of antialiased-cnns 0.3
 * (call site modules/networks.py:267); [B,H,W,C] -> [B,H/2,W/2,C]. */
int b200_maxblurpool(const void* in_hi, const void* in_lo, void* out_hi, void* out_lo, int B, int H, int W, int C,
                     void* stream);
/* ---- binary-occupancy MLP (modules/networks.py:87-115 BinaryMLPNetwork scale 0; BDModel.run_mlp_val
 * experiment_modules/bd_model.py:412-442 looped over rendered planes :293-304; infer_depth bisection :273-292) ----
 * One fused tcgen05 kernel: a 128-pixel tile of the 64-channel decoder feature (NHWC split-bf16) is TMA-loaded
 * once and every plane / bisection step is evaluated from it.  The plan owns the TMA descriptors. */
typedef struct {
  const void* feat_hi;  /* NHWC bf16 [npix, 64] (feature_s0, hi and lo halves) */
  const void* feat_lo;
  long long npix;       /* B * H * W */
  int HW;               /* H * W */
  const void* wimage;   /* 96 KB: W1[:,1:65] hi|lo, W2[:, :64] hi|lo, W2[:, 64:] hi|lo as 128x64 bf16 SW128 tiles */
  const float* vecs;    /* [6][128] fp32: b1, W1[:,0] (depth column), W1[:,65] (prior column) or 0, b2, w3, {b3} */
  int use_prior;
} b200_binary_mlp_desc;
int b200_binary_mlp_create(const b200_binary_mlp_desc* desc, void** plan_out);
/* depth [B,P,H,W] fp32; prior [B,1,H,W] fp32 or NULL (-1 everywhere if the model uses a prior); pred [B,P,H,W]. */
int b200_binary_mlp_planes(void* plan, const float* depth, int P, const float* prior, float* pred, void* stream);
/* per-pixel bisection: `iters` evaluations starting at first_depth inside [min_bound, max_bound]; a pixel is "visible"
 * at query depth z when sigmoid(logit) < threshold(z): 0.5, or -- with the evaluation's depth-dependent Thresholder
 * (utils/binary_metrics_utils.py:42-52, set by test_bd.py:91-102) -- thr_vals[bucketize(z, thr_bins)] (device arrays
 * of n_thr floats, or NULL); search_out [B,1,H,W] = final query depth, pred_out [B,1,H,W] = logit of the last
 * evaluation. */
int b200_binary_mlp_search(void* plan, const float* prior, int iters, float min_bound, float max_bound,
                           float first_depth, const float* thr_bins, const float* thr_vals, int n_thr,
                           float* search_out, float* pred_out, void* stream);
int b200_binary_mlp_destroy(void* plan);
/* BDModel.sample_prior (experiment_modules/bd_model.py:395-410): warp the previous frame's prediction into the
 * current frame through the rendered depth, nearest sampling, -1 where the rendered depth is not positive.
 *   P [B,12] = (K_s0 @ prior_cam_T_world @ world_T_cam)[:3,:4]; invK [B,16] = invK_s0; all maps [B,1,H,W]. */
int b200_sample_prior(const float* rendered_depth, const float* prior_prediction, const float* P, const float* invK,
                      float* out, int B, int H, int W, void* stream);

/* Input side of the step.
 * Relative poses of BDModel.forward (experiment_modules/bd_model.py:196-204), one launch for both products:
 *   src_cam_T_cur_cam[b,k] = src_cam_T_world[b,k] @ cur_world_T_cam[b]   (cur -> src, the managers' src_extrinsics)
 *   cur_cam_T_src_cam[b,k] = cur_cam_T_world[b] @ src_world_T_cam[b,k]   (src -> cur, the managers' src_poses)
 *   src_* [B,K,4,4]; cur_* [B,4,4]; outputs [B,K,4,4]. */
int b200_relative_poses(const float* src_cam_T_world, const float* src_world_T_cam, const float* cur_cam_T_world,
                        const float* cur_world_T_cam, float* src_cam_T_cur_cam, float* cur_cam_T_src_cam, int B,
                        int K, void* stream);
/* Intrinsics pyramid of the datasets (datasets/scannet_dataset.py:479-484): K_s[i] = K_s0 with rows 0,1 divided by
 * 2^i, invK_s[i] = inverse(K_s[i]) (fp64 adjugate rounded to fp32; the reference inverts with fp32 LAPACK).
 *   K_s0 [n,4,4]; out: K_s, invK_s [levels,n,4,4]. */
int b200_intrinsics_pyramid(const float* K_s0, float* K_s, float* invK_s, int n, int levels, void* stream);

#ifdef __cplusplus
}
#endif
#endif
