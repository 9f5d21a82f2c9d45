/*
 * libhandnet_b200 -- C ABI of the B200 (sm_100a) kernels behind the HandNet detect -> crop -> pose path.
 *
 * The reference (IRVLUTD/handnet-pipeline) is pure Python/PyTorch and has no FFI of its own
 * (SURVEY.md section 8b); each entry point below names the reference code it replaces.  The Python
 * modules in handnet-pipeline_b200/{handnet_pipeline,fcos_utils,a2j}/ bind these symbols with ctypes
 * (hn_b200/_lib.py) and keep the reference's class / function surface.
 *
 * Conventions
 *   - plain C types only; every pointer is a DEVICE pointer unless its name ends in _host;
 *   - every function returns 0 on success or a negative hn_status; hn_last_error() gives the text;
 *   - work is enqueued on the cudaStream_t passed as `stream` (void* here); nothing synchronises the
 *     device and nothing allocates device memory: scratch space is supplied by the caller;
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails with HN_ERR_CUDA.
 *
 * Activation layout ("haloed NHWC"): bf16 [N][H + 2*halo][W + 2*halo][C], halo rows/columns are zero and
 * are never written by any kernel, so a KxK convolution is a GEMM whose A operand for tap (r,s) is the
 * same matrix shifted by (r*Wp + s) rows.  Stride-2 convolutions read a "phase-split" copy
 * [4][N][ceil(H/2) + 2*halo][ceil(W/2) + 2*halo][C] (phase = (h&1)*2 + (w&1)) that the producing layer
 * writes next to its normal output.
 */
#ifndef HANDNET_B200_H
#define HANDNET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  HN_OK = 0,
  HN_ERR_ARG = -1,         /* bad argument / unsupported shape */
  HN_ERR_CUDA = -2,        /* CUDA runtime or driver error (incl. no device) */
  HN_ERR_UNSUPPORTED = -3  /* device is not sm_100 */
} hn_status;

/* ---- library ------------------------------------------------------------------------------------------ */
const char* hn_last_error(void);
int hn_version(void);
/* sm_count / cc_major / cc_minor of the current device. */
int hn_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* Number of kernels this library has launched since load (bench.py's gpu_launches). */
int64_t hn_launch_count(void);

/* ---- T1: GeneralizedRCNNTransform (fcos_utils/fcos.py:505,709; torchvision transform.py:119-255) ---------
 * (x - mean) / std, bilinear resize (align_corners=False, scale = in/out), zero-padded canvas.
 * images_host[i] points at DEVICE memory holding image i as fp32 [3][in_h[i]][in_w[i]] in 0..1; the arrays
 * themselves (images_host, in_h_host, ...) live on the host.  Output: bf16 [batch][canvas_h][canvas_w][4]
 * (channel 3 is zero); pixels outside (out_h[i], out_w[i]) are zero. */
int hn_preprocess_resize_pad(const float* const* images_host, const int* in_h_host, const int* in_w_host,
                             const int* out_h_host, const int* out_w_host, int batch, const float* mean3_host,
                             const float* std3_host, void* canvas_bf16, int canvas_h, int canvas_w, void* stream);

/* Same, but the canvas is the sub-rectangle at (pad_top, pad_left) of a larger frame of pitch_h x pitch_w pixels whose
 * remaining pixels are not touched (the caller keeps them zero): the zero-padded input of hn_conv2d_bf16's direct 7x7
 * stem (stem_pitch_*).  The frame stores its rows in PAIRS, bf16 [batch][pitch_h / 2][pitch_w][2][4]: frame pixel
 * (fy, fx) lives at [fy / 2][fx][fy & 1] (pitch_h even). */
int hn_preprocess_resize_pad_framed(const float* const* images_host, const int* in_h_host, const int* in_w_host,
                                    const int* out_h_host, const int* out_w_host, int batch, const float* mean3_host,
                                    const float* std3_host, void* canvas_bf16, int canvas_h, int canvas_w, int pad_top,
                                    int pad_left, int pitch_h, int pitch_w, void* stream);

/* ---- stem: 7x7 stride-2 pad-3 patches as GEMM rows (fcos backbone.body.conv1; a2j/resnet.py:105) ----------
 * K is laid out as 8 kernel rows x 8 pixels x C: k = (r*8 + px)*C + ch with input pixel (2*oy - 3 + r,
 * 2*ox - 4 + px); r = 7 and px = 0 are padding that meets zero weights (see pack_stem_weight in hn_b200/ops.py).
 *   in_is_f32 == 0: in = bf16 [n][h][w][4] (the T1 canvas), c = 4, k_pad = 256
 *   in_is_f32 == 1: in = fp32 [n][h][w]    (depth crops),   c = 1, k_pad = 64
 * out: bf16 [n * ceil(h/2) * ceil(w/2)][k_pad]. */
int hn_im2col_7x7s2(const void* in, int in_is_f32, int n, int h, int w, int c, void* out_bf16, int k_pad,
                    void* stream);

/* ---- convolution as shifted GEMM on tcgen05 (all 3x3 / 1x1 convs of B1, B2, H1, H2, J1, J2) ---------------- */
typedef struct hn_conv_desc {
  /* input, haloed NHWC bf16.  For stride 2, `in` is the phase-split copy and (h, w) are the per-phase
   * (= output) sizes.  For the stem GEMM use kh = kw = 1, halo_in = 0, cin = k_pad. */
  const void* in;
  int n, h, w, cin, halo_in, in_phases; /* in_phases: 1, or 4 for a phase-split input */
  /* weights: bf16, K = kh*kw*cin in tap-major order (k = (r*kw + s)*cin + c), stored k-block major:
   * [K/64][cout_pad][64], so that the tile (64 k, block_n output channels) the kernel streams is one contiguous
   * run of block_n*128 bytes.  cout_pad a multiple of block_n (rows >= cout are zero). */
  const void* weight;
  int cout, cout_pad, kh, kw, stride, dilation;
  /* epilogue: y = acc * scale[c] + shift[c] (+ residual) ; relu on channels [relu_lo, relu_hi) ; store */
  const float* scale; /* may be NULL (== 1) */
  const float* shift; /* may be NULL (== 0) */
  int relu_lo, relu_hi;
  const void* res; /* bf16 haloed NHWC with `cout` channels, or NULL */
  int res_mode;    /* 1: same pixel (res_h,res_w == output size); 2: nearest 2x upsample of a coarser map */
  int res_h, res_w, res_halo;
  /* output */
  void* out;
  int out_kind; /* 0: bf16 haloed NHWC [n][h+2*out_halo][w+2*out_halo][cout];
                   1: fp32 rows: out[(img*out_rows_per_image + out_row_offset + pix)*out_ld + c],
                      pix = h*W + w, or w*H + h when out_transpose_hw;
                   2: fp32 channel planes: out[(img*out_ld + c)*out_rows_per_image + out_row_offset + pix], out_ld = planes
                      per image (>= cout): a warp's stores of one channel are one contiguous run and a consumer that reads a
                      few channels of every location (hn_fcos_decode_select) moves no padding bytes */
  int out_halo;
  int out_rows_per_image, out_row_offset, out_ld, out_transpose_hw;
  void* out_phase; /* optional second bf16 output, phase-split with halo out_phase_halo; NULL if unused */
  int out_phase_halo;
  int64_t* gn_stats; /* optional [n][gn_groups][2] (sum, sum of squares) over the bf16-rounded output as 40.24 FIXED-POINT
                        integers (value * 2^24), accumulated with integer atomics -- caller zeroes it.  Integer addition is
                        associative: the statistics, and the GroupNorm output, are bit-identical from run to run. */
  int gn_groups;
  int block_n; /* 0 = choose automatically among 16/32/64/128/256 */
  int cluster; /* 0 or 1 (reserved: CTA pairs that multicast the weight tile measured +-1 % in round 1 and were removed) */
  int debug;   /* 0 in production; timing experiments of -DHN_CONV_DEBUG builds (bits 4, 5, 15, 16 select code paths in every build) */
  /* optional split-K for short, deep layers that cannot fill the GPU with tiles: an fp32 scratch holding one slice of
   * ceil(rows/128)*128*cout_pad*4 bytes per K split (contents irrelevant on entry) and one uint32 counter per output
   * tile, ZERO on entry and left zero on exit; both exclusive to this convolution while it runs.  Every split stores
   * its partial tile into its own slice and the last one to arrive adds the slices in split order, so the result does
   * not depend on the order in which CTAs finish.  splits = 0 lets the library choose, limited to the slices that fit
   * the scratch (1 = off; an explicit value must fit). */
  void* splitk_ws;
  int64_t splitk_ws_bytes;
  void* splitk_counters;
  int splitk_counters_len;
  int splits;
  /* bring-up only (-DHN_CONV_DEBUG builds; ignored otherwise): CTA 0 logs (clock64, tag) pairs of its producer, MMA and
   * first epilogue warp into trace[3][2048][2] (int64). */
  void* trace;
  /* Direct 7x7 stride-2 pad-3 stem over a 4-channel canvas (no im2col buffer).  0 = ordinary convolution.  Otherwise
   * `in` is the row-pair frame bf16 [n][stem_pitch_h / 2][stem_pitch_w][2][4] (frame pixel (fy, fx) at [fy / 2][fx][fy & 1]),
   * all zero except the canvas at row 3, column 4 (stem_pitch_h >= 2*h + 6 and even, stem_pitch_w >= 2*w + 8); (h, w)
   * are the OUTPUT sizes, kh = kw = 1, cin = 256, halo_in = 0, and the weights are pack_stem_weight's K = 4 kernel-row
   * pairs x 8 pixels x 2 rows x 4 channels (k-block major like every weight): a k-block of an output pixel is one
   * contiguous 128-byte run of the frame.  TMA gathers the runs straight from the frame with an overlapping-stride
   * tensor map. */
  int stem_pitch_h, stem_pitch_w;
  /* 1: WINDOW stem (the default of the detector): `in` is the plain row-major canvas bf16 [n][stem_pitch_h][stem_pitch_w][4]
   * without any frame (stem_pitch_w even, (stem_pitch_h + 1) / 2 == h, (stem_pitch_w + 1) / 2 == w, cout = cout_pad = 64);
   * borders are zero-filled by TMA.  Weights: pack_stem_weight(..., order="window"), k = ky * 32 + px * 4 + ch.  One TMA box
   * of 8 canvas rows x 256 pixels serves a tile of 125 output pixels: the im2col matrix of a kernel row is the canvas row
   * itself, read through an un-swizzled UMMA descriptor whose rows start 16 bytes apart (overlapping core matrices). */
  int stem_window;
} hn_conv_desc;
int hn_conv2d_bf16(const hn_conv_desc* desc, void* stream);

/* The same convolution over several pyramid levels in ONE launch (fcos_utils/fcos.py:278-289, 378-380: the FCOS towers and
 * output convolutions apply the same modules to P3, P4 and P5 in a Python loop).  descs[0..n_levels) (n_levels <= 3)
 * describe the levels: they share weight, scale, shift, channel counts, kernel shape, ReLU range, block_n and output kind and
 * differ in geometry (n, h, w), in, out (bf16 outputs; fp32-row outputs share one buffer and differ in out_row_offset) and
 * gn_stats.  Only plain 3x3 stride-1 convolutions (no residual, phase copy, split-K or stem).  The tiles of all levels
 * form one list that is dealt over the persistent CTAs, so the small levels no longer leave most of the 148 SMs idle.
 * Fused tile configurations: 256-wide bf16 layers (cout a multiple of 256) and 16-wide fp32-row outputs with enough tiles
 * for resident weights; other shapes run level by level inside the call (same results). */
int hn_conv2d_bf16_levels(const hn_conv_desc* descs, int n_levels, void* stream);

/* Upper bound on the CTAs (SMs) the following hn_conv2d_bf16* launches of this process use (0 = all SMs, the default).
 * The runtime caps the latency-bound pose-net launches so that they fit next to the detector kernels of the next step. */
int hn_conv_set_cta_cap(int max_ctas);
/* Programmatic dependent launch for the following hn_conv2d_bf16* launches (default 1: a kernel's CTAs may become resident
 * and wait inside the kernel while its predecessor in the stream is still running; 0: plain stream order). */
int hn_conv_set_pdl(int enabled);

/* GroupNorm + ReLU (hn_groupnorm_relu below) over several pyramid levels in one launch: x[i] has n[i] x h[i] x w[i] pixels,
 * stats[i] is that level's [n][groups][2] array; c, halo, groups, gamma, beta and eps are shared.  n_levels <= 3; the
 * arrays live on the host. */
int hn_groupnorm_relu_levels(void* const* x_host, const int* n_host, const int* h_host, const int* w_host, int n_levels, int c,
                             int halo, const int64_t* const* stats_host, int groups, const float* gamma, const float* beta,
                             float eps, void* stream);

/* ---- frame ingest (ros_demo.py:227-238, 266-267): what the camera delivers -> what HandNet.forward takes ---------
 * bgr_u8: DEVICE uint8 [n][h][w][3] (cv2 BGR) -> rgb_out fp32 [n][3][h][w] = float32(byte) / 255 in R,G,B order;
 * depth_u16: DEVICE uint16 [n][h][w] millimetres -> depth_out fp32 [n][1][h][w] = float32(mm) / 1000.  Either pair may
 * be NULL.  Bit-exact with numpy's astype(float32) / 255.0 and / 1000.0. */
int hn_ingest_frames(const void* bgr_u8, const void* depth_u16, int n, int h, int w, float* rgb_out, float* depth_out,
                     void* stream);

/* fp32 NCHW images -> the zero-framed 4-channel bf16 canvas of the direct stem (hn_conv_desc.stem_pitch_*; row-pair
 * layout, see hn_preprocess_resize_pad_framed): frame pixel (pad_top + y, pad_left + x), channel j = src channel
 * chan_map4_host[j] (-1 = zero).  The RGBD A2J variant feeds its
 * 4-channel crops this way with the reference's [2,1,0,3] reorder (handnet_pipeline.py:102) as the channel map. */
int hn_pack_nhwc4_frame(const float* src, int n, int c, int h, int w, const int* chan_map4_host, void* frame_bf16,
                        int pad_top, int pad_left, int pitch_h, int pitch_w, void* stream);

/* a2j.convert_joints + uvd2xyz, batched on the device (a2j/a2j.py:17-43, datasets3d/a2jdataset.py:31-38).
 * uvd [n][joints][3] crop-space joints, crops int64 [n][4] = (x_min, y_min, x_max, y_max), paras4_dev = (fx, fy, cx, cy)
 * in DEVICE memory as 4 doubles (paras_is_f64 = 1) or 4 floats (0), or NULL (pixels only).  out [n][joints][3]:
 * u*(x_max-x_min)/crop_w + x_min etc. evaluated in float64 and stored as float32, as numpy >= 2 evaluates a float32
 * array times an int64 scalar; with paras, (uv - c) * z / f in the intrinsics' precision -> float32, then * 1000 in
 * float32.  has_hand (int32 [n]) may be NULL; frames with has_hand == 0 produce zeros. */
int hn_convert_joints(const float* uvd, const int64_t* crops, const int* has_hand, const void* paras4_dev,
                      int paras_is_f64, int n, int joints, int crop_w, int crop_h, float* out, void* stream);

/* ---- 3x3 stride-2 pad-1 max pool (torchvision resnet maxpool; a2j/resnet.py:108) ----------------------------
 * in: bf16 [n][h][w][c] (no halo) -> out: bf16 haloed NHWC [n][oh+2*halo][ow+2*halo][c], oh = (h+1)/2. */
int hn_maxpool3x3s2(const void* in, int n, int h, int w, int c, void* out, int out_halo, void* stream);

/* ---- GroupNorm + ReLU of the FCOS towers (fcos_utils/fcos.py:232-240, 352-360) ------------------------------
 * x: bf16 haloed NHWC, normalised in place: relu((x - mean) * rstd * gamma + beta) with mean/rstd from
 * stats[n][groups][2] (fixed-point sums as produced by hn_conv2d_bf16) over h*w*(c/groups) elements.  Halo stays zero. */
int hn_groupnorm_relu(void* x, int n, int h, int w, int c, int halo, const int64_t* stats, int groups,
                      const float* gamma, const float* beta, float eps, void* stream);

/* ---- P1..P4: decode + score + threshold (fcos_utils/fcos.py:591-632; det_utils.py:266-294;
 *      anchor_utils.py:56-132) -------------------------------------------------------------------------------
 * Head tensors are fp32 [batch][locs][channels] in ANY layout: element (b, loc, c) of a tensor lives at
 * ptr[b * img_stride + loc * loc_stride + c * chan_stride] (the detector's fused output convolutions write channel planes:
 * loc_stride 1, chan_stride locs; row layouts [batch][locs][ld] are loc_stride ld, chan_stride 1).
 * Levels are described by level_h/w/stride (host arrays).  For every
 * location: box = anchor-centre -/+ reg * anchor-size (anchors are generated on the fly), score = max over classes
 * of sqrt(sigmoid(cls_c) * sigmoid(ctr)), label = first arg-max class (class 0 included), candidate when
 * score > score_thresh (a double, compared the way `scores > 0.7` is in the reference).  The sigmoids / square roots are
 * evaluated only where the outcome is open (a location whose centre-ness or largest class logit is below
 * logit(score_thresh^2) - 0.01 cannot pass).  One kernel: every block of 1024 locations scores, counts, looks back over
 * the counts of the blocks before it in the same image (one state word per block in `workspace`, cleared by a memset node
 * in front of the kernel) and writes its survivors in place.  Candidates are written per image in ascending location order:
 *   cand_count[batch], cand_loc[batch][locs] (int32), cand_score[batch][locs] (fp32),
 *   cand_label[batch][locs] (int32), cand_box[batch][locs][4] (fp32, canvas pixels, NOT clipped). */
int hn_fcos_decode_select(const float* cls_logits, int64_t cls_img_stride, int cls_loc_stride, int cls_chan_stride,
                          const float* bbox_ctrness, int64_t ctr_img_stride, int ctr_loc_stride,
                          const float* bbox_regression, int64_t reg_img_stride, int reg_loc_stride, int reg_chan_stride,
                          int batch, int locs, int num_classes, int num_levels, const int* level_h_host,
                          const int* level_w_host, const int* level_stride_h_host, const int* level_stride_w_host,
                          const int* level_anchor_host, double score_thresh, int* cand_count, int* cand_loc,
                          float* cand_score, int* cand_label, float* cand_box, void* workspace,
                          int64_t workspace_bytes, void* stream);
int64_t hn_fcos_select_workspace_bytes(int batch, int locs);

/* ---- P5: batched NMS (fcos_utils/fcos.py:635 -> torchvision.ops.boxes.batched_nms, CPU semantics) ----------
 * Per image: stable descending score order; class-aware greedy suppression with IoU evaluated in fp32 and
 * compared as torchvision's CPU kernel does ((double)iou > (double)iou_thresh).  When 4*count <=
 * coord_trick_max_numel the "coordinate trick" of torchvision (boxes + label*(max+1)) is reproduced, above it
 * suppression is restricted to equal labels.  keep[batch][cap] receives indices INTO THE CANDIDATE LIST in
 * score-descending order (ties: ascending candidate index), keep_count[batch] their number.
 * workspace: hn_nms_workspace_bytes(batch, cap) bytes. */
int64_t hn_nms_workspace_bytes(int batch, int cap);
int hn_nms_batched(const float* cand_box, const float* cand_score, const int* cand_label, const int* cand_count,
                   int batch, int cap, double iou_thresh, int coord_trick_max_numel, int* keep, int* keep_count,
                   void* workspace, int64_t workspace_bytes, void* stream);

/* ---- P6: gather kept detections + resize_boxes (fcos_utils/fcos.py:648-669, 770-783) -----------------------
 * ratio_h[i] = orig_h / resized_h, ratio_w likewise, as fp32 (host arrays).  Outputs are dense per image with
 * capacity cap: boxes[batch][cap][4], scores[batch][cap], labels[batch][cap] (int64), sides[batch][cap]
 * (int64, argmax of hand_lr), level[batch][cap] (fp32 pyramid level = the reference's feature_idx).
 * hand_lr is fp32 rows [batch][locs][lr_ld] (2 used).  Optional ext heads (pass NULL to skip):
 * contact_logits rows [..][contact_ld] (5 used) -> contacts[batch][cap] (int64, argmax of sigmoid);
 * dxdy rows [..][dxdy_ld] (3 used, already ReLU'd) -> dxdymags[batch][cap][3] = (d0, 0.1 * normalize(d1, d2))
 * (fcos_utils/fcos.py:299-303). */
int hn_fcos_gather(const int* keep, const int* keep_count, const int* cand_loc, const float* cand_score,
                   const int* cand_label, const float* cand_box, const float* hand_lr, int64_t lr_img_stride,
                   int lr_loc_stride, int lr_chan_stride, const float* contact_logits, int64_t contact_img_stride,
                   int contact_loc_stride, int contact_chan_stride, const float* dxdy, int64_t dxdy_img_stride,
                   int dxdy_loc_stride, int dxdy_chan_stride, int batch, int cap, int locs, int num_levels,
                   const int* level_start_host,
                   const float* ratio_h_host, const float* ratio_w_host, float* boxes, float* scores, int64_t* labels,
                   int64_t* sides, float* level, int64_t* contacts, float* dxdymags, void* stream);

/* ---- S1 + S2: hand select, box pad, depth crop, nearest resize (handnet_pipeline.py:74-102) ------------------
 * For image i: first kept detection with label == hand_label; truncate to integers; pad by 0.4*w / 0.4*h in
 * fp32 and clamp to the image; crop depth[i][:, y1:y2+1, x1:x2+1] and resize to out_size x out_size with the
 * legacy nearest rule.  depth: fp32 [batch][depth_c][img_h][img_w].  Outputs: crops[batch][4] (int64),
 * has_hand[batch] (int32), depth_batch fp32 [batch][depth_c][out][out] (zeros where no hand). */
int hn_select_crop_resize(const float* boxes, const int64_t* labels, const int* keep_count, int batch, int cap,
                          int hand_label, const float* depth, int depth_c, int img_h, int img_w, int out_size,
                          int64_t* crops, int* has_hand, float* depth_batch, void* stream);
/* The same per-box path for the first `max_hands` hand detections of every image (an extension: the reference keeps
 * boxes[:1], handnet_pipeline.py:84-85; BASELINE.json config 5 names "up to 4 hands/frame").  Output slot i*max_hands + h
 * holds the h-th kept detection of image i with label == hand_label: crops[batch*max_hands][4], has_hand[batch*max_hands],
 * depth_batch [batch*max_hands][depth_c][out][out]; slots without a detection are zero.  max_hands == 1 is
 * hn_select_crop_resize. */
int hn_select_crop_resize_multi(const float* boxes, const int64_t* labels, const int* keep_count, int batch, int cap,
                                int hand_label, int max_hands, const float* depth, int depth_c, int img_h, int img_w,
                                int out_size, int64_t* crops, int* has_hand, float* depth_batch, void* stream);

/* ---- J4: anchor aggregation (a2j/anchor.py:57-82) ------------------------------------------------------------
 * cls [n][anchors][joints], reg [n][anchors][joints][2], depth [n][anchors][joints] fp32; anchor_xy
 * [anchors][2].  out [n][joints][3] = (sum w*(a0+r0), sum w*(a1+r1), sum w*d), w = softmax over anchors.
 * workspace: hn_a2j_workspace_bytes(n, joints). */
int64_t hn_a2j_workspace_bytes(int n, int joints);
int hn_a2j_aggregate(const float* cls, const float* reg, const float* depth, const float* anchor_xy, int n,
                     int anchors, int joints, float* out, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- pose2mesh lifting (last row): models.pose2mesh_net.FlatPose2Mesh ---------------------------
 * fp32 SIMT kernels; one hand is 21 joints -> ~1000 vertices x <= 256 features (weight streaming and 5-nonzero rows).
 *
 * Chebyshev term of a graph convolution (pose2mesh/lib/models/backbones/cheby_graph_conv.py:26-31): with the rescaled graph
 * Laplacian L [verts x verts] in CSR form and x, z, out [batch][verts][feats]:  out = alpha * L x + beta * z  (z may be NULL).
 * T1 = L T0 is (1, 0); T2 = 2 L T1 - T0 is (2, -1) with z = T0. */
int hn_cheby_spmm(const int* row_ptr, const int* col, const float* val, const float* x, const float* z, float alpha, float beta,
                  int batch, int verts, int feats, float* out, void* stream);
/* y[m][n] = post(sum_kk pre(A[m][kk]) * weight[n][kk] + bias[n]) (+ res[m][n]).  A is `planes` (1..3) matrices a0, a1, a2 of
 * [m][fin] and kk = f * planes + p reads plane p, feature f: the order in which graph_conv_cheby flattens its Chebyshev terms
 * (cheby_graph_conv.py:33-35); planes = 1 is an ordinary [m][fin] matrix.  pre(a) = relu(a * in_scale[kk] + in_shift[kk]) when
 * in_scale != NULL (eval-mode BatchNorm1d + ReLU in front of the layer, posenet.py:25-28); post(t) = t * out_scale[n] +
 * out_shift[n] when out_scale != NULL (eval-mode BatchNorm1d behind it, cheby_graph_conv.py:40-41), then ReLU when relu_out.
 * weight [n][fin * planes], bias / res may be NULL. */
int hn_linear_f32(const float* a0, const float* a1, const float* a2, int planes, int m, int fin, const float* weight,
                  const float* bias, int n, const float* in_scale, const float* in_shift, const float* out_scale,
                  const float* out_shift, int relu_out, const float* res, float* y, void* stream);
/* Residual of a mesh block (meshnet.py:107-114, 69-76): out[r * up + j][f] = x[r][f] + interp(skip[r][0..fskip), f) for
 * j < up, where interp is F.interpolate(mode='linear', align_corners=False) to fout samples along the FEATURE axis and up = 2
 * is the nearest x2 vertex upsample.  x [rows][fout], skip [rows][fskip], out [rows * up][fout]. */
int hn_mesh_residual_upsample(const float* x, const float* skip, int rows, int fout, int fskip, int up, float* out,
                              void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HANDNET_B200_H */
