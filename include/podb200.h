/*
 * podb200.h -- C ABI of the B200-native probabilistic-inference path.
 *
 * Drop-in boundary for the hot path of asharakeh/pod_compare (reference paths relative to
 * /root/reference/src).  The reference has no native layer: this path is pure Python over
 * torch / torchvision / detectron2, so each entry point below names the reference *Python*
 * code whose arithmetic it replaces; INTEGRATION.md shows the ctypes binding a maintainer
 * adds in probabilistic_inference.py to call them.
 *
 * Conventions
 *   - every function returns 0 on success, <0 for a bad argument, >0 for a CUDA error code;
 *     pod_last_error() returns the thread-local message.  No exceptions cross the ABI.
 *   - all pointers are DEVICE pointers unless the name ends in _host; the caller owns every
 *     buffer (torch-allocated); the library allocates nothing that outlives a call.
 *   - `stream` is a cudaStream_t passed as void*; no call synchronises the device.
 *   - activations are channels-last: map n, pixel (y*W+x), channel c  ->  ((n*H+y)*W+x)*C+c.
 *   - "split" tensors are the fp16 pair (hi, lo) with  x*scale ~= hi + lo  (hi = rn_fp16(x*scale),
 *     lo = rn_fp16(x*scale - hi)); three fp16 tensor-core products hi*hi' + hi*lo' + lo*hi'
 *     accumulated in fp32 reproduce the fp32 convolution to ~2^-22 relative.
 */
#ifndef PODB200_H_
#define PODB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define POD_ABI_VERSION 2

const char* pod_last_error(void);
int pod_version(void);
/* 1 if the current device can run the sm_100a kernels (compute capability 10.x). */
int pod_device_ok(void);
/* Device-side error word of the library, read and cleared (this call synchronises the device).  0 = ok;
 * 1..99 = a bounded mbarrier wait of the tcgen05 kernels expired (the code names the wait); 100 = a hidden tower
 * activation left the fp16 split range (|x * out_scale| > 65504: the reference computes in fp32 and has no such
 * limit, so the condition is reported instead of propagating inf / NaN); 101 = the same in pod_mask_expand_split;
 * 102 = non-finite input feature (pod_absmax_accumulate).  The product path (predictor.infer_from_features) calls
 * it once per inference call and raises PodError on a non-zero word. */
int pod_status(int* status_host);

/* ---- counter-based random streams (parity contract with oracle/philox.py) -------------------
 * Replace torch's generator at the three places the reference draws randomness:
 * nn.Dropout (probabilistic_modeling/probabilistic_retinanet.py:422-424), Normal.rsample
 * (probabilistic_inference/probabilistic_inference.py:291-294) and MultivariateNormal.rsample
 * (:351-356).  These three fill functions expose the streams for tests. */
int pod_philox_dropout_mask(uint8_t* keep_hwc, int H, int W, int C, uint64_t seed, int image, int sample,
                            int pass, int tower, int layer, int level, double p, void* stream);
int pod_philox_logit_normals(float* out_draws_n_k, int draws, int n_anchor, int K, uint64_t seed, int image,
                             int level, void* stream);
int pod_philox_box_normals(float* out_draws_m_4, const int64_t* anchor_ids, int M, int draws, uint64_t seed,
                           int image, void* stream);

/* ---- operand preparation ---------------------------------------------------------------------
 * FPN feature maps (NCHW fp32, as the backbone returns them at probabilistic_retinanet.py:99-100)
 * -> channels-last fp16 split pair. */
int pod_nchw_to_nhwc_split(const float* src_nchw, int NB, int C, int H, int W, float scale,
                           void* dst_hi, void* dst_lo, void* stream);
/* fp16 split scale of the input maps without a host round trip: pod_absmax_accumulate folds max|x| of a tensor into
 * *amax_bits (device word, zero it first; bit pattern of a non-negative float), pod_pow2_scale_from_absmax writes
 * min(max_scale, largest power of two s with amax * s <= target) to *scale_dev, and pod_nchw_to_nhwc_split_dev /
 * pod_conv_args.in_scale_dev read it from device memory. */
int pod_absmax_accumulate(const float* x, int64_t n, uint32_t* amax_bits, void* stream);
int pod_pow2_scale_from_absmax(const uint32_t* amax_bits, float max_scale, float target, float* scale_dev, void* stream);
int pod_nchw_to_nhwc_split_dev(const float* src_nchw, int NB, int C, int H, int W, const float* scale_dev,
                               void* dst_hi, void* dst_lo, void* stream);
/* Same layout change, fp32 out (input of the SIMT cross-check convolution). */
int pod_nchw_to_nhwc_f32(const float* src_nchw, int NB, int C, int H, int W, float* dst, void* stream);
/* nn.Conv2d weight (Cout, Cin, 3, 3) fp32 -> K-major GEMM operand rows [Cout_pad][9*Cin] with
 * k = (ky*3+kx)*Cin + ci, as fp16 split pair scaled by `scale` (rows >= Cout are zero).
 * Replaces nothing arithmetic: it is the weight-pack step of the checkpoint loader
 * (probabilistic_inference.py:72-84). */
int pod_pack_conv_weight(const float* w_oihw, int Cout, int Cin, int Cout_pad, float scale,
                         void* dst_hi, void* dst_lo, void* stream);
/* Same K-major order, fp32, transposed to [9*Cin][Cout_pad] for the SIMT convolution. */
int pod_pack_conv_weight_f32(const float* w_oihw, int Cout, int Cin, int Cout_pad, float* dst, void* stream);

/* Dropout description for one convolution launch: map index n of the OUTPUT decodes as
 * image = image0 + n / (samples*passes), sample = (n / passes) % samples, pass = pass0 + n % passes.
 * p == 0 disables dropout. */
typedef struct pod_dropout {
  double p;
  uint64_t seed;
  int image0;
  int samples;
  int passes;
  int pass0;
  int tower;
  int layer;
  int level;
} pod_dropout;

/* MC-dropout replication of the (mask-independent) first tower layer, SURVEY Q2:
 * x (NB_in, H*W, C) fp32 post-ReLU  ->  NB_in*samples*passes masked, rescaled, split copies.
 * Replaces the first nn.Dropout of every tower evaluation (probabilistic_retinanet.py:422-424,518-523).
 * live_reps (0 = all): only the first live_reps of the samples*passes copies of every image are written
 * (see pod_conv_args.map_live). */
int pod_mask_expand_split(const float* x, int NB_in, int HW, int C, const pod_dropout* d, float scale,
                          void* dst_hi, void* dst_lo, int live_reps, const float* scale_dev /* nullable: replaces scale */,
                          void* stream);

/* ---- the head convolutions (probabilistic_retinanet.py:401-441,458-484,517-523) -------------
 * 3x3 / stride 1 / pad 1 convolution over NB channels-last maps as an implicit GEMM on the
 * tcgen05 tensor cores (TMA-fed, fp32 accumulation in TMEM, fp16x3 split operands).
 *   in_hi/in_lo : (NB, H, W, Cin) fp16 split pair scaled by in_scale; consecutive maps are
 *                 in_map_stride ELEMENTS apart (lets a launch read every 2nd map).
 *   w_hi/w_lo   : [Cout_pad][9*Cin] from pod_pack_conv_weight, scaled by w_scale
 *   bias        : Cout_pad fp32
 * mode POD_OUT_HIDDEN : y = dropout(relu(conv + bias)) -> split pair (NB,H,W,Cout_pad) scaled by out_scale
 * mode POD_OUT_RAW    : y = conv + bias (relu if `relu`) -> fp32, element (n, pixel, c<Cout) at
 *                       out_f32[n*out_map_stride + pixel*out_pixel_stride + c]; with Cout = A*K this
 *                       IS permute_to_N_HWA_K (probabilistic_retinanet.py:343-349). */
enum { POD_OUT_HIDDEN = 0, POD_OUT_RAW = 1 };
typedef struct pod_conv_args {
  const void* in_hi;
  const void* in_lo;
  int64_t in_map_stride;
  float in_scale;
  int NB, H, W, Cin;
  const void* w_hi;
  const void* w_lo;
  float w_scale;
  const float* bias;
  int Cout, Cout_pad;
  int mode;
  int relu;
  void* out_hi;
  void* out_lo;
  float out_scale;
  float* out_f32;
  int64_t out_map_stride;
  int64_t out_pixel_stride;
  pod_dropout drop;
  /* POD_OUT_RAW only, optional second destination: output channels c >= split_col go to
   * out2_f32[n*out2_map_stride + pixel*out2_pixel_stride + (c - split_col)]; out2_f32 == NULL means a single
   * destination.  split_col may be <= 0 (every channel of this launch belongs to the second destination).
   * Lets one launch evaluate two heads that read the same tower output (eval mode: cls_score|cls_var and
   * bbox_pred|bbox_cov, probabilistic_retinanet.py:518-523 with identical tower passes). */
  float* out2_f32;
  int split_col;
  int64_t out2_map_stride;
  int64_t out2_pixel_stride;
  /* Optional live-map selection (0 = every map): the NB maps come in groups of map_group (one image's samples x
   * passes) and only the first map_live maps of each group are evaluated; the others' outputs are left
   * untouched.  Used to leave out the tower passes of the last MC sample / ensemble member, whose class logits
   * and variance outputs the reference computes but never reads (probabilistic_inference.py:216-267: the sample
   * "mean" runs over range(len-1); only box_delta of the last sample is used, :326-331). */
  int map_group;
  int map_live;
  /* Optional device-resident split scales (powers of two written by pod_pow2_scale_from_absmax); when non-NULL they
   * replace in_scale / out_scale.  The engine derives ONE scale per call from max|feature| and uses it for every
   * activation of the towers, so that inputs of any magnitude stay inside the fp16 split range without a host
   * round trip (values are computed in fp32 registers in true units; only the stored pair is scaled). */
  const float* in_scale_dev;
  const float* out_scale_dev;
  /* Optional Q1 sample accumulation (POD_OUT_HIDDEN, 256 channels, CTA-pair kernel): the NB maps are
   * images x q1_samples x q1_passes (sample-major, pass-minor).  For every pass p whose bit is set in q1_acc_mask the
   * layer's output is NOT written per sample; instead the reference's sample sum  2*y_0 + y_1 + ... + y_{L-1}
   * (probabilistic_inference.py:214-270, quirk Q1; L = q1_live[p] samples are evaluated, the others are neither
   * computed nor read) is accumulated in fp32 into
   *     q1_acc[((image * n_acc + a) * G + g) * H*W*Cout + pixel*Cout + c],   a = index of p among the accumulated passes,
   * as G = ceil(q1_samples / q1_group) partial sums over consecutive groups of q1_group samples (a fixed group size
   * keeps the summation order independent of the batch).  pod_q1_finish adds the partial sums, divides by q1_samples
   * and writes the split pair that the (linear) output convolutions cls_score / cls_var / bbox_cov then read ONCE per
   * image:  mean_s conv(x_s) == conv(mean_s x_s).  Passes not in the mask are written per sample as usual
   * (bbox_pred: every sample's deltas enter the epistemic covariance, :326-331).  q1_acc == NULL: off. */
  float* q1_acc;
  int q1_samples, q1_passes;
  int q1_live[2];
  int q1_acc_mask;
  int q1_group;
  /* Optional in-kernel dropout of the INPUT (first masked tower layer; POD_OUT_HIDDEN, 256 -> 256, CTA-pair kernel):
   * with mask_in = 1 the input holds ONE map per image -- the mask-independent first tower layer, already multiplied by
   * 1/(1-p) (produced by a launch with drop_scale_only = 1) -- and output map n = (image, sample, pass) reads it with
   * the keep mask of stream layer mask_in_layer applied to the staged tile in shared memory.  Replaces
   * pod_mask_expand_split + a plain launch bit for bit; the samples x passes masked copies never exist in HBM
   * (probabilistic_retinanet.py:422-424 applied 16N times per level in the reference). */
  int mask_in;
  int mask_in_layer;
  /* POD_OUT_HIDDEN with drop.p > 0: multiply by 1/(1-p) and keep every element (no mask). */
  int drop_scale_only;
} pod_conv_args;
int pod_conv3x3_tc(const pod_conv_args* a, void* stream);
/* Channels per pipeline stage of the tcgen05 kernel: 64 (SWIZZLE_128B operand tiles, default) or
 * 32 (SWIZZLE_64B, deeper pipeline).  Process-wide tuning knob; results are identical. */
int pod_conv3x3_tc_set_kblock(int bk);
/* Accumulation chunk = how much of the K loop is summed inside the tensor core (fp32 TMEM, truncating adds) before
 * the epilogue warps add the partial sum in fp32 round-to-nearest.  Two knobs, ONE wins: chunk_kblocks (default 12
 * K-blocks of 64 channels = 3 taps) is used whenever it is > 0 and divides the K-block count of the convolution;
 * only otherwise (or after pod_conv3x3_tc_set_chunk_kblocks(0)) the taps setting applies (default 1 tap per chunk;
 * 3 or 9 = coarser; 9 = a single 2304-long chain, which drifts ~7e-6 relative without the compensation below). */
int pod_conv3x3_tc_set_chunk_taps(int taps);
int pod_conv3x3_tc_set_chunk_kblocks(int kb);
/* Compensation of the tensor core's accumulate-with-truncation: expected loss per MMA accumulation in fp32 ulps
 * of the running sum (default 0.27, measured; 0 = off).  The epilogue scale is multiplied by 1 + ulps * (MMAs per TMEM chain) * 2^-24. */
int pod_conv3x3_tc_set_trunc_comp(float ulps_per_mma);
/* Weights-as-A kernel for RAW convolutions of <= 64 output channels (default on): the weight tile with its hi and
 * lo halves stacked along M is the A operand and 256 pixels are the N operand, two full-size MMAs per K-step.
 * Taken only when w_lo == w_hi + 64 rows (pod_pack_conv_weight into one buffer) and Cout_pad == 64. */
int pod_conv3x3_tc_set_wt(int on);
/* Row-halo activation staging (K-block 64 only): one TMA box two rows taller than the pixel tile per column shift
 * serves the three row-shifted taps through descriptor offsets, cutting activation bytes into the SM ~2.5x.
 * mode bit 0: pixels-as-M kernels, bit 1: weights-as-A kernel (default 3 = both).  Process-wide tuning
 * knob; the K order of the accumulation differs (results agree to fp32 round-off). */
int pod_conv3x3_tc_set_halo(int mode);
/* 256-output-channel convolutions on CTA pairs (tcgen05 cta_group::2, default on) or on single CTAs. */
int pod_conv3x3_tc_set_pair(int on);
/* Pixel-tile width of the CTA-pair row-halo kernel: 16 (16 x 8 tiles), 32 (32 x 4 tiles) or 0 = chosen per map shape
 * (default: 32 x 4 where it covers the map with >= 2 % fewer tiles, e.g. 160 x 92).  Every output pixel sees the same
 * MMAs in the same order either way: results are bit-identical.  Process-wide tuning / test knob. */
int pod_conv3x3_tc_set_tile_width(int tw);
/* Device-side error word of the tcgen05 kernels only (see pod_status).  Host pointer out. */
int pod_conv3x3_tc_status(int* status_host);
/* Test hooks: the cycle budget of every bounded mbarrier wait (default ~2 s; <= 0 restores it) and a fault
 * injection that makes the TMA producer of the first CTA / CTA pair skip its loads, so that the bounded waits of that
 * CTA expire and the error path (pod_status != 0 -> PodError in the predictor) can be exercised. */
int pod_conv3x3_tc_set_wait_limit(long long cycles);
int pod_conv3x3_tc_debug_fault(int on);
/* Measurement hook: enable (out == NULL), then after CTA-pair launches read {SM cycles, nanoseconds} of the last one
 * (out = 2 words, synchronises): cycles / ns is the clock the tensor kernel really ran at under the power cap. */
int pod_conv3x3_tc_debug_clock(int enable, long long* out_cycles_ns_host);

/* ---- general convolution for the ResNet-50-FPN backbone (detectron2 build_retinanet_resnet_fpn_backbone,
 *      call sites probabilistic_retinanet.py:96-101) -------------------------------------------------------------------
 * 1x1 or 3x3 (padding ksize/2), stride 1 or 2, channels-last fp16 split pair in, on the same tcgen05 implicit-GEMM
 * kernel as the head (fp16x3 split operands, fp32 TMEM accumulation, chunked K loop).  FrozenBatchNorm is folded into
 * the packed weights (scale) and `bias`.  One launch evaluates ONE column block of block_cols (64 / 128 / 256) output
 * channels starting at col0 of a weight matrix packed by pod_pack_conv_weight_k with Cout_rows rows.
 *   out_hi/out_lo : (NB, Hout, Wout, out_ch_stride) split pair scaled by out_scale; this block writes channels
 *                   [col0, col0 + block_cols);  y = relu?(conv + bias (+ residual))
 *   res_hi/res_lo : optional residual with the layout of the output (bottleneck shortcut), scaled by res_scale
 *   out_f32       : alternative fp32 channels-last output (NB, Hout, Wout, out_ch_stride) -- the FPN maps handed to the head */
typedef struct pod_convg_args {
  const void* in_hi;
  const void* in_lo;
  float in_scale;
  int NB, Hin, Win, Cin;
  int ksize, stride;
  int Hout, Wout;
  const void* w_hi;
  const void* w_lo;
  float w_scale;
  int Cout_rows;          /* rows of the packed weight matrix (multiple of block_cols) */
  int Cout;               /* real output channels */
  int col0, block_cols;
  const float* bias;      /* Cout_rows fp32 */
  int relu;
  void* out_hi;
  void* out_lo;
  float out_scale;
  int out_ch_stride;
  const void* res_hi;
  const void* res_lo;
  float res_scale;
  float* out_f32;
} pod_convg_args;
int pod_conv_tc_general(const pod_convg_args* a, void* stream);
/* nn.Conv2d weight (Cout, Cin, k, k) -> K-major rows [Cout_pad][k*k*Cin], column (ky*k+kx)*Cin + ci, fp16 split pair. */
int pod_pack_conv_weight_k(const float* w_oihw, int Cout, int Cin, int ksize, int Cout_pad, float scale,
                           void* dst_hi, void* dst_lo, void* stream);
/* Stem of the ResNet: (image - mean) / std, zero-padded from (Himg, Wimg) to the size-divisible (H, W) (detectron2
 * preprocess_image: ImageList.from_tensors pads the NORMALISED image with zeros) -> 7x7 / stride 2 / pad 3 convolution
 * 3 -> 64 (+ folded FrozenBN, ReLU) -> 3x3 / stride 2 / pad 1 max-pool, SIMT fp32.  images (NB, 3, Himg, Wimg) uint8
 * (is_u8) or fp32; w [(ky*7+kx)*3+c][64] with the BN scale folded in, bias 64.  out: (NB, Hp, Wp, 64) split pair,
 * Hp = ceil(ceil(H/2)/2). */
int pod_stem_conv7_pool(const void* images_nchw, int is_u8, int NB, int Himg, int Wimg, int H, int W,
                        const float* mean3_host, const float* std3_host, const float* w147x64, const float* bias,
                        float* scratch /* NB*Hc*Wc*64 fp32, Hc = ceil(H/2) */, void* out_hi, void* out_lo, float out_scale,
                        void* stream);
/* FPN top-down step on fp32 channels-last maps: dst (NB,H,W,C) += nearest-upsample-by-2 of src (NB,ceil(H/2),ceil(W/2),C). */
int pod_upsample2_add(float* dst, const float* src, int NB, int H, int W, int C, void* stream);
/* fp32 channels-last -> fp16 split pair with a device-resident scale (FPN maps -> head operand), optional ReLU first. */
int pod_split_f32(const float* src, int64_t n, float scale, const float* scale_dev, int relu, void* dst_hi, void* dst_lo,
                  void* stream);

/* Plain fp32 SIMT convolution with the same semantics (cross-check of the tensor-core kernel;
 * not on the product path). in (NB,H,W,Cin) fp32, w from pod_pack_conv_weight_f32. */
int pod_conv3x3_simt(const float* in, int NB, int H, int W, int Cin, const float* w_kc, const float* bias,
                     int Cout, int Cout_pad, int relu, const pod_dropout* drop, float* out,
                     int64_t out_map_stride, int64_t out_pixel_stride, void* stream);

/* Second half of the Q1 sample accumulation: acc (n_maps, groups, n) fp32 partial sums -> (n_maps, n) split pair of
 * ((g_0 + g_1) + ...) / samples, scaled by *scale_dev (or `scale` when scale_dev is NULL). */
int pod_q1_finish(const float* acc, int n_maps, int groups, int64_t n, int samples, float scale, const float* scale_dev,
                  void* dst_hi, void* dst_lo, void* stream);

/* Streaming form of the same mean (the default, measured faster than the epilogue accumulation: DESIGN.md 3.7): the last
 * tower layer writes its per-sample split pairs as usual (images x samples x passes maps of n elements) and this kernel
 * reads them once: for every pass in `mask`, ((x_0 + x_0) + x_1 + ... + x_{live[p]-1}) / samples -> split pair map
 * image * n_acc + a of dst.  HBM-bound. */
int pod_q1_mean_act(const void* in_hi, const void* in_lo, int images, int samples, int passes, int mask,
                    const int* live_host, int64_t n, float scale, const float* scale_dev, void* dst_hi, void* dst_lo,
                    void* stream);

/* ---- per-anchor sample statistics (probabilistic_inference.py:214-270, quirk Q1) -------------
 * x (B, S, n) fp32 -> out (B, n):  ((x0 + x0) + x1 + ... + x_{S-2}) / S  in that fp32 order
 * (S == 1: copy). */
int pod_sample_mean_q1(const float* x, int B, int S, int64_t n, float* out, void* stream);

/* ---- scores and per-level top-k (probabilistic_inference.py:283-308) ------------------------
 * For image b, level l (anchors [off_l, off_l+n_l) of R): class probabilities
 *   with logvar : mean_j sigmoid(mu + eps_j * sqrt(exp(logvar))), j < draws (STREAM_LOGIT)
 *   without     : sigmoid(mu)
 * then max/argmax over K.  probs (B,R,K), score (B,R), cls (B,R) int32. */
int pod_scores(const float* logits, const float* logvar /*nullable*/, int B, int R, int K, int n_levels,
               const int* level_off /*host, n_levels+1*/, int draws, uint64_t seed, int image0, int runs,
               float* probs, float* score, int* cls, void* stream);
/* `runs` (>= 1): row b of the batch is run (b % runs) of image image0 + b / runs -- the post-NMS merge modes
 * evaluate every MC sample / ensemble member as its own inference with its own noise draws
 * (probabilistic_inference.py:444-461,506-511); the run index enters the Philox counter. */
/* Per (image, level): the min(topk, n_l) highest scores, descending, ties -> lower anchor index,
 * then those > thresh.  cand_idx (B, cap) holds GLOBAL anchor ids, level l's segment starts at
 * seg_off[l] (host array, n_levels+1, seg_off[l+1]-seg_off[l] = min(topk, n_l)); cand_cnt (B, n_levels). */
int pod_topk_levels(const float* score, int B, int R, int n_levels, const int* level_off_host,
                    const int* seg_off_host, int topk, float thresh, int* cand_idx, int* cand_cnt, void* stream);

/* ---- decode + covariance of the candidates (probabilistic_inference.py:310-388,
 *      inference_utils.py:337-371,510-547, modeling_utils.py:4-22) ----------------------------
 * One warp per candidate.  Outputs are written at the COMPACTED position
 * m = sum_{l'<l} cnt[l'] + rank (the order of the reference's torch.cat over levels).
 *   mean_delta (B,R,4); mean_regvar (B,R,cd) or NULL; sample_delta (B,S,R,4) or NULL (S>1: epistemic)
 *   anchors (R,4); probs/score/cls from pod_scores
 * out: boxes (B,cap,4) cov (B,cap,16) scores (B,cap) classes (B,cap) i32 probs (B,cap,K) count (B)
 *      cand_anchor (B,cap) global anchor id per compacted candidate; has_cov = cov is meaningful. */
typedef struct pod_decode_args {
  const float* mean_delta;
  const float* mean_regvar;
  int cov_dims;
  const float* sample_delta;
  int S;
  const float* anchors;
  const float* probs;
  const float* score;
  const int* cls;
  const int* cand_idx;
  const int* cand_cnt;
  int B, R, K, n_levels, cap;
  const int* seg_off_host;
  int box_draws;
  uint64_t seed;
  int image0;
  int runs;   /* see pod_scores */
  float wx, wy, ww, wh;
  float* out_boxes;
  float* out_cov;
  float* out_scores;
  int* out_classes;
  float* out_probs;
  int* out_count;
  int* out_anchor;
  /* regression weights of the SAMPLED decode (inference_utils.py:510-547 SampleBox2BoxTransform, built from
   * cfg.MODEL.RPN.BBOX_REG_WEIGHTS at probabilistic_inference.py:175-176) -- the deterministic / per-sample decodes
   * above use the RetinaNet weights (wx..wh).  All four 0 = same as wx..wh. */
  float swx, swy, sww, swh;
} pod_decode_args;
int pod_decode_cov(const pod_decode_args* a, void* stream);

/* ---- NMS / BayesOD fusion / rescale (inference_utils.py:12-54,292-334,374-425;
 *      probabilistic_inference.py:536-636; torchvision ops/boxes.py:51-120 + cpu/nms_kernel.cpp) ---
 * One CTA per image.  mode: 0 standard NMS, 1 BayesOD, 2 anchor statistics (inference_utils.py:57-162).
 * nms_variant: 0 per-class ("vanilla"),
 * 1 coordinate-offset trick, 2 auto = torchvision's CPU rule (4*M > 4000 -> vanilla).
 * box_merge: 0 bayesian_inference, 1 covariance_intersection; cls_merge: 0 max_score,
 * 1 bayesian_inference (mean of member vectors).
 * in : candidates from pod_decode_cov (B,cap,...) + count (B); has_cov: 0 -> zero covariances
 * out: det_* (B,max_dets,...) + det_count (B) after scale/clip/nonempty; keep (B,max_dets) = NMS
 *      survivor indices into the candidate list BEFORE the nonempty filter, keep_count (B);
 *      det_src (B,max_dets) = the candidate index of every final row. */
typedef struct pod_nms_args {
  const float* boxes;
  const float* cov;
  const float* scores;
  const int* classes;
  const float* probs;
  const int* count;
  int B, cap, K;
  int has_cov;
  int mode, nms_variant, box_merge, cls_merge;
  double nms_thresh;
  double affinity;
  int max_dets;
  int in_h, in_w, out_h, out_w;
  float* det_boxes;
  float* det_cov;
  float* det_scores;
  int* det_classes;
  float* det_probs;
  int* det_count;
  int* keep;
  int* keep_count;
  int* det_src;  /* (B,max_dets): candidate index each FINAL detection row came from */
  int skip_post; /* 1: stop after NMS/fusion (no rescale, clip, nonempty filter, covariance conditioning) */
} pod_nms_args;
int pod_nms_fuse(const pod_nms_args* a, void* stream);

/* ---- post-NMS merging of per-run detections (inference_utils.py:165-289,
 *      general_black_box_ensembles_post_processing; reached from probabilistic_inference.py:444-481,506-534) ----
 * in : per-run detections from pod_nms_fuse(skip_post=1), rows b*runs + r: det_* (B*runs, max_dets, ...), det_count.
 * The runs of an image are concatenated run-major; boxes are clustered sequentially (seed i unless already a
 * member of an earlier cluster; members = IoU >= affinity and same class); every cluster yields the mean box,
 * sample covariance (+ mean member covariance), mean probability vector and its max/argmax.
 * out: cluster_* (B, runs*max_dets, ...) + cluster_count (B): a candidate set for pod_nms_fuse (final NMS + rescale). */
typedef struct pod_merge_args {
  const float* det_boxes;
  const float* det_cov;
  const float* det_probs;
  const int* det_classes;
  const int* det_count;
  int B, runs, max_dets, K;
  double affinity;
  float* out_boxes;
  float* out_cov;
  float* out_scores;
  int* out_classes;
  float* out_probs;
  int* out_count;
  int* seed_scratch;   /* (B, runs*max_dets) ints: seed list handed from the clustering kernel to the statistics kernel */
} pod_merge_args;
int pod_cluster_merge(const pod_merge_args* a, void* stream);

/* ---- wire format (inference_utils.py:428-502 covar_xyxy_to_xywh / instances_to_json; src/apply_net.py:53-102) ----
 * Detections of a whole batch (pod_nms_fuse outputs) -> fixed-size fp32 records, one row of
 * 1 + max_dets * (4 + 1 + 1 + K + 16) floats per image: count, then per detection box(4) score class probs(K)
 * cov(16); rows past `count` are zero.
 *   xywh == 0: boxes xyxy and covariance as they are -- the record of the multi-GPU all-gather;
 *   xywh == 1: the layout of coco_instances_results.json -- boxes XYWH_ABS, covariance T Sigma T^T
 *              (T = [[1,0,0,0],[0,1,0,0],[-1,0,1,0],[0,-1,0,1]]), bit-identical to the reference's fp32 arithmetic.
 * cat_map (device, K ints, nullable): contiguous class id -> dataset category id, -1 = class has no id in the test
 * dataset (the host writer drops those rows, inference_utils.py:489). */
typedef struct pod_wire_args {
  const float* det_boxes;
  const float* det_cov;
  const float* det_scores;
  const int* det_classes;
  const float* det_probs;
  const int* det_count;
  int B, max_dets, K;
  int xywh;
  const int* cat_map;
  float* records;
} pod_wire_args;
int pod_wire_records(const pod_wire_args* a, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PODB200_H_ */
