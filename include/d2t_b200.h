/*
 * d2t_b200.h -- C ABI of libd2t_b200.so: the Detect-to-Track per-frame-pair hot path as
 * hand-written sm_100a CUDA kernels.  Plain pointers and sizes only; no torch types.
 *
 * Part 1 re-exports, name for name and argument for argument, the `extern "C"` launchers
 * the reference's own C glue links against; each prototype cites
 * the reference declaration it replaces (paths relative to /root/reference/lib/model/).
 * Part 2 is the stream-ordered / batched / workspace-explicit surface the Python host
 * layer actually calls.
 *
 * Conventions for every entry point:
 *   - all tensor pointers are DEVICE pointers to contiguous fp32 / int32 NCHW data;
 *   - return 1 = success, 0 = failure (never exit(), never printf); after a 0 the message
 *     is available from d2t_last_error();
 *   - kernels are launched on the stream passed in; nothing synchronises the device except
 *     nms_cuda_compute(), whose reference contract is synchronous;
 *   - the caller owns every buffer.  Part 1 symbols that need scratch use a per-device
 *     cache inside the library (grown with cudaMalloc on first use, guarded by a mutex:
 *     one in-flight call per device per op); Part 2 symbols take the workspace explicitly.
 */
#ifndef D2T_B200_H
#define D2T_B200_H

#include <stddef.h>
#include <stdint.h>

#ifndef __DRIVER_TYPES_H__
typedef struct CUstream_st* cudaStream_t;
#endif

#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif
#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ library ---------- */
const char* d2t_version(void);
const char* d2t_last_error(void);          /* thread-local, "" when no error */
int d2t_device_sm_count(void);             /* SMs of the current device (148 on B200) */

/* ====================================================================================
 * Part 1 -- reference launcher symbols
 * ==================================================================================== */

/* correlation/src/correlation_cuda_kernel.h:5-39.  rInput1/rInput2 (the reference's padded
 * NHWC scratch) are accepted and ignored.  Strides are accepted; tensors must be contiguous
 * NCHW (the reference kernels assume it too).  `output` need not be pre-zeroed. */
int Correlation_forward_cuda_kernel(
    float* output, int ob, int oc, int oh, int ow, int osb, int osc, int osh, int osw,
    float* input1, int ic, int ih, int iw, int isb, int isc, int ish, int isw,
    float* input2, int gc, int gsb, int gsc, int gsh, int gsw,
    float* rInput1, float* rInput2,
    int pad_size, int kernel_size, int max_displacement, int stride1, int stride2,
    int corr_type_multiply, cudaStream_t stream);

/* correlation/src/correlation_cuda_kernel.h:41-88.  Writes the exact adjoint of the forward
 * into gradInput1/gradInput2 (every element is written; pre-zeroing is not required).  See
 * DESIGN.md "Correlation backward" for where the reference kernels deviate from it. */
int Correlation_backward_cuda_kernel(
    float* gradOutput, int gob, int goc, int goh, int gow, int gosb, int gosc, int gosh, int gosw,
    float* input1, int ic, int ih, int iw, int isb, int isc, int ish, int isw,
    float* input2, int gsb, int gsc, int gsh, int gsw,
    float* gradInput1, int gisb, int gisc, int gish, int gisw,
    float* gradInput2, int ggc, int ggsb, int ggsc, int ggsh, int ggsw,
    float* rInput1, float* rInput2,
    int pad_size, int kernel_size, int max_displacement, int stride1, int stride2,
    int corr_type_multiply, cudaStream_t stream);

/* psroi_pooling/src/psroi_pooling_kernel.h:8-11.  mapping_channel may be NULL (skipped). */
int PSROIPoolForwardLauncher(
    const float* bottom_data, const float spatial_scale, const int num_rois, const int height,
    const int width, const int channels, const int pooled_height, const int pooled_width,
    const float* bottom_rois, const int group_size, const int output_dim, float* top_data,
    int* mapping_channel, cudaStream_t stream);

/* psroi_pooling/src/psroi_pooling_kernel.h:14 (pooled_width BEFORE pooled_height, as in the
 * reference).  Adds into bottom_diff, which the caller zero-fills (functions/psroi_pool.py:40).
 * mapping_channel may be NULL: the channel is a pure function of the output index
 * (group_size is then taken to be pooled_width). */
int PSROIPoolBackwardLauncher(
    const float* top_diff, const int* mapping_channel, const int batch_size, const int num_rois,
    const float spatial_scale, const int channels, const int height, const int width,
    const int pooled_width, const int pooled_height, const int output_dim, float* bottom_diff,
    const float* bottom_rois, cudaStream_t stream);

/* roi_align/src/roi_align_kernel.h:18-22 and :29-33 (sic: "Laucher"). */
int ROIAlignForwardLaucher(
    const float* bottom_data, const float spatial_scale, const int num_rois, const int height,
    const int width, const int channels, const int aligned_height, const int aligned_width,
    const float* bottom_rois, float* top_data, cudaStream_t stream);
int ROIAlignBackwardLaucher(
    const float* top_diff, const float spatial_scale, const int batch_size, const int num_rois,
    const int height, const int width, const int channels, const int aligned_height,
    const int aligned_width, const float* bottom_rois, float* bottom_diff, cudaStream_t stream);

/* roi_pooling/src/roi_pooling_kernel.h:8-18 (sic).  argmax_data may be NULL in forward. */
int ROIPoolForwardLaucher(
    const float* bottom_data, const float spatial_scale, const int num_rois, const int height,
    const int width, const int channels, const int pooled_height, const int pooled_width,
    const float* bottom_rois, float* top_data, int* argmax_data, cudaStream_t stream);
int ROIPoolBackwardLaucher(
    const float* top_diff, const float spatial_scale, const int batch_size, const int num_rois,
    const int height, const int width, const int channels, const int pooled_height,
    const int pooled_width, const float* bottom_rois, float* bottom_diff, const int* argmax_data,
    cudaStream_t stream);

/* roi_crop/src/roi_crop_cuda_kernel.h:6-17 and :19-32.  inputImages is NCHW despite the
 * "BHWD" name; grids is [ob, oh, ow, 2] in (y, x) order.  Only contiguous tensors. */
int BilinearSamplerBHWD_updateOutput_cuda_kernel(
    int oc, int ow, int oh, int ob, int ic, int ih, int iw, int ib,
    float* inputImages, int isb, int isc, int ish, int isw,
    float* grids, int gsb, int gsc, int gsh, int gsw,
    float* output, int osb, int osc, int osh, int osw, cudaStream_t stream);
int BilinearSamplerBHWD_updateGradInput_cuda_kernel(
    int goc, int gow, int goh, int gob, int ic, int ih, int iw, int ib,
    float* inputImages, int isb, int isc, int ish, int isw,
    float* grids, int gsb, int gsc, int gsh, int gsw,
    float* gradInputImages, int gisb, int gisc, int gish, int gisw,
    float* gradGrids, int ggsb, int ggsc, int ggsh, int ggsw,
    float* gradOutput, int gosb, int gosc, int gosh, int gosw, cudaStream_t stream);

/* nms/src/nms_cuda_kernel.h:5-6.  All three pointers are DEVICE pointers (as the reference's
 * caller passes them).  Synchronous on return, like the reference.  keep_out[0..*num_out) =
 * kept indices ascending; entries beyond are left untouched. */
void nms_cuda_compute(int* keep_out, int* num_out, float* boxes_host, int boxes_num,
                      int boxes_dim, float nms_overlap_thresh);

/* ====================================================================================
 * Part 2 -- stream-ordered surface used by the Python host layer
 * ==================================================================================== */

/* Opt-in (environment D2T_NMS_PREFIX=1): d2t_nms_batched with max_keep > 0 first decides a prefix of the (sorted) lists --
 * greedy NMS never looks ahead -- and runs the full pass only for the lists that did not reach max_keep inside it.
 * Returns the prefix length that would be used (0: no prefix pass). */
int d2t_nms_prefix(int N, int max_keep);
/* Lists capped at 0 < max_keep <= 2048 (the proposal step: proposal_layer.py:149-156 keeps the first post_nms_topN survivors)
 * can take a mask-free kernel: one CTA per list keeps the survivors in shared memory and tests each candidate only against
 * them (6000 x 300 IoUs instead of the 18 M of the mask's upper triangle, one launch, same keep-set bit for bit).  Faster
 * when the cap is reached early (6.7 against 37 us per image on spread boxes), slower on heavily clustered lists where one
 * SM meets a long kept list for every chunk -- so it is opt-in: d2t_nms_set_mode(1) or D2T_NMS_GREEDY=1; (0) forces the
 * mask + sweep pair, (-1) restores the environment default.  d2t_nms_launch_count: kernels d2t_nms_batched launches for
 * (N, max_keep) under the current mode. */
int d2t_nms_set_mode(int greedy);
int d2t_nms_launch_count(int N, int max_keep);
/* ---- NMS: B independent, caller-sorted box lists in one launch pair ----
 * boxes   [B, N, box_dim] fp32 (x1,y1,x2,y2,...), n_valid [B] int32 or NULL (= N each)
 * keep    [B, keep_stride] int32, num_keep [B] int32; at most max_keep (<= keep_stride)
 *         indices are produced per list (max_keep <= 0: no cap, needs keep_stride >= N).
 * workspace: d2t_nms_workspace_bytes(B, N) bytes, 16-byte aligned. */
size_t d2t_nms_workspace_bytes(int B, int N);
int d2t_nms_batched(const float* boxes, const int* n_valid, int B, int N, int box_dim,
                    float thresh, int max_keep, int* keep, int keep_stride, int* num_keep,
                    void* workspace, size_t workspace_bytes, cudaStream_t stream);

/* Deterministic backward of the three RoI operators (the reference kernels -- roi_align_kernel.cu:94-143,
 * roi_pooling_kernel.cu:128-203, roi_crop_cuda_kernel.cu:111-194 -- and the launchers of Part 1 add fp32 terms with
 * atomicAdd in arrival order: last bits change from run to run).  Here every term is accumulated as a 64-bit fixed-point
 * integer with the L2's native integer atomic (order-independent) and the finished sums are converted to float once:
 * bit-identical from run to run, each term resolved to < 2^-37 of max |top_diff| for up to 2^24 terms.  A NaN / Inf
 * gradient falls back to the float atomics.  scratch: d2t_roi_backward_scratch_bytes(elements of the gradient tensor) bytes,
 * 8-byte aligned.  align / crop ADD to bottom_diff / grad_images (the caller zeroes, as for the reference launchers); pool
 * assigns. */
size_t d2t_roi_backward_scratch_bytes(size_t grad_elems);
int d2t_roi_align_backward_det(const float* top_diff, float spatial_scale, int batch_size, int num_rois, int height,
                               int width, int channels, int aligned_height, int aligned_width, const float* bottom_rois,
                               float* bottom_diff, void* scratch, size_t scratch_bytes, cudaStream_t stream);
int d2t_roi_pool_backward_det(const float* top_diff, int batch_size, int num_rois, int height, int width, int channels,
                              int pooled_height, int pooled_width, float* bottom_diff, const int* argmax_data,
                              void* scratch, size_t scratch_bytes, cudaStream_t stream);
int d2t_roi_crop_backward_det(const float* grids, const float* grad_output, float* grad_images, int batch_size,
                              int channels, int height, int width, int num_rois, int grid_h, int grid_w, void* scratch,
                              size_t scratch_bytes, cudaStream_t stream);

/* Kernel variant of d2t_psroi_forward / _backward, process-wide (tests and A/B measurements; the defaults are the product):
 * forward_mode -1 = chosen by the geometry alone (integer tables when the plane fits, never by batch size or SM count),
 * 0 = exactly-rounded fp64 tables, 1..4 = development variants of the integer tables; backward_mode 0 = two-limb integer
 * difference tables (exact sums, native shared-memory atomics), 2 = fp64 difference
 * tables, 1 = one-limb integer tables (experiment). */
int d2t_psroi_set_mode(int forward_mode, int backward_mode);
/* Fused PSRoI pooling + 7x7 vote (+ softmax) -- rfcn.py:62-64, 133-140, 194-196: vote[n][d] = the mean of the 49
 * pooled bins of (roi n, output channel d), optionally followed by the softmax over d, without the [R, D, 7, 7] tensor
 * reaching memory.  Same windows / tables as d2t_psroi_forward; pooled size 7x7, planes at most 64 wide (returns 0
 * otherwise: use d2t_psroi_forward + a mean). */
size_t d2t_psroi_vote_workspace_bytes(int num_rois, int batch, int pooled_h, int pooled_w, int out_dim);
int d2t_psroi_vote_forward(const float* bottom, int batch, int channels, int height, int width,
                           const float* rois, int num_rois, float scale, int pooled_h, int pooled_w, int group,
                           int out_dim, int softmax, float* vote, void* workspace, size_t workspace_bytes,
                           cudaStream_t stream);
/* ---- PSRoI with explicit workspace; mapping may be NULL; accumulate=0 overwrites the
 * touched planes of bottom_diff (no pre-zeroing needed for channels < D*G*G).
 * Replaces PSROIPoolForwardLauncher / PSROIPoolBackwardLauncher (psroi_pooling/src/psroi_pooling_kernel.h:8-14) for
 * callers that know the batch size.  Forward, 7x7 bins: bins are the reference's bit for bit; values come from 2-D
 * prefix-sum tables -- per-plane fixed-point int32 tables (|error| <= 2^-30 * L1 norm of the plane) when
 * batch * out_dim * 7 exceeds the SM count and width <= 64, exactly rounded fp64 tables otherwise or when the
 * environment has D2T_PSROI_INT=0.  The workspace must hold d2t_psroi_workspace_bytes() bytes, 4-byte aligned. ---- */
size_t d2t_psroi_workspace_bytes(int num_rois, int batch, int pooled_h, int pooled_w);
int d2t_psroi_forward(const float* bottom, int batch, int channels, int height, int width,
                      const float* rois, int num_rois, float scale, int pooled_h, int pooled_w,
                      int group, int out_dim, float* top, int* mapping,
                      void* workspace, size_t workspace_bytes, cudaStream_t stream);
int d2t_psroi_backward(const float* top_diff, int batch, int channels, int height, int width,
                       const float* rois, int num_rois, float scale, int pooled_h, int pooled_w,
                       int group, int out_dim, float* bottom_diff, int accumulate,
                       void* workspace, size_t workspace_bytes, cudaStream_t stream);
/* Integer bin windows [num_rois, PH, PW, 4] = (hstart, hend, wstart, wend), for the
 * bit-exact RoI->bin parity test. */
int d2t_psroi_bins(const float* rois, int num_rois, float scale, int pooled_h, int pooled_w,
                   int height, int width, int* bins, cudaStream_t stream);

/* ---- Correlation without the legacy stride/scratch arguments ---- */
int d2t_correlation_shape(int H, int W, int pad, int k, int md, int s1, int s2, int* out3);
int d2t_correlation_forward(const float* in1, const float* in2, int B, int C, int H, int W,
                            int pad, int k, int md, int s1, int s2, float* out,
                            cudaStream_t stream);
int d2t_correlation_backward(const float* in1, const float* in2, const float* grad_out,
                             int B, int C, int H, int W, int pad, int k, int md, int s1,
                             int s2, float* grad1, float* grad2, cudaStream_t stream);

/* ---- RPN proposal step (rpn/proposal_layer.py:67-159) ----
 * decode+clip: anchors [A,4], deltas [B,4A,H,W], scores = fg half of cls_prob [B,2A,H,W],
 * im_info [B,3] -> boxes [B, H*W*A, 4] and scores_out [B, H*W*A] in (y, x, a) order. */
int d2t_proposal_decode(const float* anchors, int A, const float* deltas, const float* cls_prob,
                        const float* im_info, int B, int H, int W, int feat_stride,
                        float* boxes, float* scores_out, cudaStream_t stream);
/* gather the top `n_take` boxes of each image through `order` [B, order_stride] (int64)
 * into dets [B, n_take, 5] = (x1,y1,x2,y2,score). */
int d2t_proposal_gather(const float* boxes, const float* scores, const int64_t* order,
                        int B, int n_total, int order_stride, int n_take, float* dets,
                        cudaStream_t stream);
/* rois [B, post, 5]: col 0 = image index, rows >= num_keep[b] zero (proposal_layer.py:127,158-159) */
/* Top-N selection + stable descending sort + gather in one launch: dets[b][j] = (box, score) of the j-th highest score of
 * image b -- exactly what a stable descending sort of all scores (ties in index order, NaN first: torch.sort) followed by
 * d2t_proposal_gather of the first n_take produces (proposal_layer.py:115-137), without sorting the other ~80 % of the list.
 * One CTA per image: radix select (keys in registers), compaction, bitonic sort in shared memory.  n_take <= n_total <= 32768,
 * n_take <= 16384 (d2t_proposal_topk_supported). */
int d2t_proposal_topk_supported(int n_total, int n_take);
int d2t_proposal_topk_gather(const float* boxes, const float* scores, int B, int n_total, int n_take, float* dets,
                             cudaStream_t stream);
/* The same result in six short launches that use the whole device -- two 16-bit histogram levels in global memory with their
 * scans (the n_take-th key), compaction, then a rank sort of the n_take distinct (score, index) composites fused with the
 * gather -- for any n_take <= n_total <= 32768.  scratch:
 * d2t_proposal_topk_scratch_bytes(B, n_take) bytes, 8-byte aligned. */
size_t d2t_proposal_topk_scratch_bytes(int B, int n_take);
int d2t_proposal_topk_gather_split(const float* boxes, const float* scores, int B, int n_total, int n_take, float* dets,
                                   void* scratch, size_t scratch_bytes, cudaStream_t stream);
int d2t_proposal_write_rois(const float* dets, const int* keep, int keep_stride,
                            const int* num_keep, int B, int n_take, int post, float* rois,
                            cudaStream_t stream);


/* ---- Convolution engine (tcgen05 / TMEM / TMA implicit GEMM) ----
 * Replaces the cuDNN convolutions + eval-mode BatchNorm + ReLU the reference reaches through
 * torch.nn for the ResNet-101 trunk and the heads (faster_rcnn/resnet.py:66-129, 258-312, 333-344;
 * faster_rcnn/rfcn.py:49-53; rpn/rpn.py:28-36, 62-71).
 * Activations are plain fp32 NHWC tensors [N, H, W, cstride].  Weights are packed once by
 * d2t_conv_pack_weights into the pair (w, w_lo), w_lo = w - trunc13(w) (trunc13 = low 13 mantissa
 * bits cleared).  kind::tf32 reads an fp32 operand as its truncation, so passes = 3 evaluates
 * hi*hi + hi*lo + lo*hi on the tensor cores (fp32-level accuracy; the activation lo tile is derived
 * in shared memory); passes = 1 is a plain single TF32 pass.
 * passes = 16 ("3xFP16") evaluates the same three products on kind::f16 -- twice the TF32 rate --
 * with fp16 (hi, lo) operand pairs.  fp16 lacks fp32's range, so operands carry power-of-two scales:
 * weights are packed by d2t_conv_pack_weights_f16 with 2^w_exp; every activation tensor owns one
 * float in device memory holding a running max |x| ("amax": written by the producing plan's epilogue
 * with atomicMax, zeroed by the caller before the producer runs), from which the consuming kernel
 * derives its scale.  The fp32 activation tile is split in shared memory; HBM tensors stay fp32. */
typedef struct d2t_conv_desc {
    int N, H, W;              /* input [N, H, W, in_cstride] */
    int Cin;                  /* input channels read, a multiple of 32 (zero padded) */
    int in_cstride;           /* channels per pixel of the input buffer, a multiple of 4 */
    int Cout, R, S;           /* filter [Cout, R, S, Cin] packed by d2t_conv_pack_weights */
    int stride, pad, dil;
    int passes;               /* 3, 1 or 16 (see above) */
    int relu;
    int out_cstride;          /* channels per pixel of the NHWC output buffer */
    int out_coffset;          /* this conv writes channels [out_coffset, out_coffset + Cout) */
    int res_cstride;          /* channels per pixel of the residual buffer (0: = Cout) */
    int w_exp;                /* passes = 16: the packed weights are w * 2^w_exp (d2t_conv_pack_weights_f16) */
} d2t_conv_desc;
typedef struct d2t_conv_plan d2t_conv_plan;

/* out = relu?( scale[c] * conv(in, w) + shift[c] + res ).  scale / shift / res / out / out_nchw may
 * be NULL (at least one output is required; passes = 16: scale must be NULL -- fold it into the packed weights); out_nchw is a plain fp32 [N, Cout, OH, OW] copy for
 * consumers that keep the reference's layout.  The plan captures the pointers (TMA tensor maps);
 * buffers must outlive it.  Plans of one device share a small stream-K scratch unless given their own
 * (d2t_conv_plan_set_scratch); d2t_conv_plan_run keeps launches on the shared scratch stream-ordered
 * (a launch arriving on another stream first waits for the previous one).  Returns NULL on error. */
d2t_conv_plan* d2t_conv_plan_create(const d2t_conv_desc* desc, const float* in, const void* w_hi,
                                    const void* w_lo, const float* scale, const float* shift,
                                    const float* res, float* out, float* out_nchw);
/* Attach the per-tensor running max |x| scalars (device pointers, one float each): amax_in belongs to
 * the plan's input tensor (required for passes = 16), amax_out to its NHWC output (optional: the
 * epilogue folds max |out| into it).  Works for stem and correlation plans too (amax_out). */
int d2t_conv_plan_set_amax(d2t_conv_plan* plan, const float* amax_in, float* amax_out);
/* Stream-K scratch.  By default the plans of a device share one scratch (partial tiles + flags), which is why they must
 * not run concurrently.  A caller that runs several chains of plans on different streams gives each chain its own
 * zero-initialised device buffer of d2t_conv_scratch_bytes() bytes. */
size_t d2t_conv_scratch_bytes(void);
int d2t_conv_plan_set_scratch(d2t_conv_plan* plan, void* scratch, size_t bytes);
/* Optional completion hand-shake for CONSECUTIVE launches of one chain on one stream (experiment, off unless called):
 * every CTA of `plan` adds 1 to *self_counter once its outputs are globally visible; if prev_counter is given (the
 * self_counter of `prev`, the plan launched immediately before on the same stream), the kernel polls it up to prev's grid
 * size instead of executing griddepcontrol.wait.  The caller zeroes the counters before each pass over the chain (the
 * engine keeps them in the arena its per-forward memset clears).  Null pointers restore the default. */
int d2t_conv_plan_set_done(d2t_conv_plan* plan, const d2t_conv_plan* prev, const int* prev_counter, int* self_counter);
/* Let a stand-alone launch of the plan fetch its first weight tiles BEFORE it waits for the previous launch on the stream
 * (programmatic dependent launch): the first touch of a layer's weights is a DRAM miss that otherwise sits in front of the
 * first MMA.  Only for weights that no kernel up to and including the launch immediately before this one writes (an
 * inference engine with weights packed once; NOT a training loop that re-packs them right before the layer).  Off by default. */
int d2t_conv_plan_set_early_weights(d2t_conv_plan* plan, int on);
void d2t_conv_plan_destroy(d2t_conv_plan* plan);
/* out8 = {OH, OW, tile_h, tile_w, BN, m_tiles, n_tiles, grid*10 + pair_mode} */
int d2t_conv_plan_info(const d2t_conv_plan* plan, int* out8);
int d2t_conv_plan_run(const d2t_conv_plan* plan, cudaStream_t stream);

/* ---- Layer chains: a LIST of 3xFP16 convolution plans run by ONE persistent launch ----
 * The reference's trunk is ~110 dependent cuDNN calls (faster_rcnn/resnet.py:66-109, 258-312; rfcn.py:49-53; rpn/rpn.py:62-71);
 * launched one by one, every kernel boundary idles the SMs for several microseconds.  A chain keeps one CTA per SM
 * resident over the whole list (tensor memory and barriers stay allocated, the tensor maps of layer i are read from a
 * descriptor array) and separates DEPENDENT layers by a grid-wide barrier instead of a kernel boundary.
 *   d2t_conv_plan_chainable: 1 if the plan can be a chain layer (plain 3xFP16 convolution with its input amax set: no
 *                            stem / correlation / weight-gradient / mask / completion-counter plans).
 *   d2t_conv_chain_create:   plans[i] becomes layer i.  sync_before[i] != 0 (or sync_before == NULL): layer i reads -- as
 *                            input or residual -- what a layer since the last such mark writes, so the CTAs meet at a
 *                            barrier before it; independent neighbours run back to back.  dev_buf: d2t_conv_chain_bytes(n)
 *                            bytes of device memory, 256-byte aligned, owned by the caller for the life of the chain.
 *                            The plans' tensor maps and arguments are COPIED: create the chain after every
 *                            d2t_conv_plan_set_* call; all plans must share one stream-K scratch.
 *   d2t_conv_chain_run:      one cooperative launch on `stream`; results are bit-identical to running the plans one by one. */
typedef struct d2t_conv_chain d2t_conv_chain;
int d2t_conv_plan_chainable(const d2t_conv_plan* plan);
size_t d2t_conv_chain_bytes(int n_layers);
d2t_conv_chain* d2t_conv_chain_create(const d2t_conv_plan* const* plans, const int* sync_before, int n_layers,
                                      void* dev_buf, size_t bytes);
int d2t_conv_chain_run(const d2t_conv_chain* chain, cudaStream_t stream);
void d2t_conv_chain_destroy(d2t_conv_chain* chain);

/* The 7x7 stride-2 pad-3 stem conv (faster_rcnn/resnet.py:116) on the same kernel: the image is
 * first packed by d2t_stem_pack_input into a zero-bordered NHWC4 buffer [N, (H+7)&~1, W+8, 4] and
 * the filter by d2t_stem_pack_weights into [Cout][7][32] (w, w_lo); output NHWC [N, OH, OW, out_cstride]. */
d2t_conv_plan* d2t_conv_stem_plan_create(int N, int H, int W, int Cout, int passes, const float* in,
                                         const float* w_hi, const float* w_lo, const float* scale,
                                         const float* shift, int relu, float* out, int out_cstride);
int d2t_stem_pack_input(const float* x_nchw, int N, int C, int H, int W, float* packed, cudaStream_t stream);
/* the same with max |x| folded into *amax (device float, zeroed by the caller): the operand scale of a passes = 16 stem plan,
 * whose weights are fp16 (hi, lo) [Cout][4][64] -- the [Cout][7][32] rows of d2t_stem_pack_weights padded to 8 rows, times the
 * per-channel scale and 2^k -- with max |w * scale| in a device scalar (d2t_conv_plan_set_weight_amax).  pairs > 0: x is the
 * reference's [pairs][2 legs] frame batch (im_data) and the packed batch is leg-major, frame n = leg * pairs + pair (N = 2 pairs);
 * pairs = 0: frames in order. */
int d2t_stem_pack_input_amax(const float* x_nchw, int N, int C, int H, int W, float* packed, float* amax, int pairs, cudaStream_t stream);
int d2t_stem_pack_weights(const float* w_oihw, int Cout, int Cin, float* w_hi, float* w_lo,
                          cudaStream_t stream);

/* Cross-frame correlation (correlation/src/correlation_cuda_kernel.cu:34-106) for kernel_size 1,
 * stride1 == stride2 = stride, 1 <= max_displacement/stride <= 8, on the same tensor-core pipeline:
 * in1 / in2 are NHWC [N, H, W, in_cstride] (C % 32 == 0 channels read, c_real of them meaningful:
 * the divisor); the (2r+1)^2-channel result goes to channels [out_coffset, ...) of an NHWC buffer
 * and/or to a plain fp32 NCHW tensor. */
d2t_conv_plan* d2t_corr_plan_create(int N, int C, int c_real, int H, int W, int in_cstride, int pad,
                                    int max_displacement, int stride, int passes, const float* in1,
                                    const float* in2, float* out, int out_cstride, int out_coffset,
                                    float* out_nchw);

/* Re-pack MANY operands in one launch (a training engine after every optimizer step: ~630 operands of a Res-101 D&T net).
 * items: device array of n_items records of d2t_conv_repack_item_bytes() bytes each, laid out as
 *   { const float* w_oihw; const float* scale_or_null; const float* amax; void* hi; void* lo;
 *     int O, I, R, S, pad (cin_pad | cout_pad), rows (backward-data: rows of the operand), dgrad (0 | 1), first_block; }
 * = the arguments of d2t_conv_pack_weights_f16_dev (dgrad = 0) / _dgrad (dgrad = 1); item i owns blocks
 * [first_block_i, first_block_{i+1}) with d2t_conv_repack_item_blocks(...) blocks (one per tile of 32 x 32 channels; 0: filter
 * with more than 9 taps, not supported), first_block_0 = 0, total_blocks their sum.
 * block_item (device, [total_blocks] int32, may be NULL): the item every block belongs to -- saves each block a binary search
 * over the items.  Same values, bit for bit, as the per-operand entry points. */
size_t d2t_conv_repack_item_bytes(void);
int d2t_conv_repack_item_blocks(int O, int I, int R, int S, int pad, int rows, int dgrad);
int d2t_conv_repack_many(const void* items, int n_items, int total_blocks, const int* block_item, cudaStream_t stream);
/* ---- Training path of the convolution engine: backward-data and weight-gradient on the same tcgen05 kernel ----
 * Replaces the cuDNN dgrad / wgrad calls autograd makes for the trainable convolutions of the reference's training step
 * (trainval_net.py:365-373 over faster_rcnn/resnet.py:66-109, 279-295, rfcn.py:49-53, rpn/rpn.py:28-36); 3xFP16 like
 * the forward pass (fp32-accurate).
 *
 * Backward-data of  y = scale * conv(x, w)  (stride 1) is itself a convolution of dy with the transposed, flipped
 * filter  wt[ci][r'][s'][co] = w[co][ci][R-1-r'][S-1-s'] * scale[co]  and padding dil*(R-1) - pad:
 * d2t_conv_pack_weights_f16_dgrad builds wt in the engine's packed format and d2t_conv_plan_create runs it (the
 * residual input adds the gradient arriving over the skip connection, in place if res == out).
 * d2t_conv_plan_set_mask fuses the ReLU backward of the tensor whose gradient the plan produces.  A stride-2 1x1
 * convolution's backward-data runs at the output resolution and is scattered by d2t_upsample2_add_mask.
 * Packed weights whose scale changes every optimizer step take it from a device scalar (max |w|):
 * d2t_conv_pack_weights_f16_dev / _dgrad read it, d2t_conv_plan_set_weight_amax makes the kernel read it. */
int d2t_conv_plan_set_mask(d2t_conv_plan* plan, const float* mask_nhwc, int mask_cstride);
int d2t_conv_plan_set_weight_amax(d2t_conv_plan* plan, const float* amax_w);
int d2t_conv_pack_weights_f16_dev(const float* w_oihw, const float* scale, int Cout, int Cin, int R, int S, int cin_pad,
                                  const float* amax_w, void* w_hi, void* w_lo, cudaStream_t stream);
/* (rows >= Cin: rows [Cin, rows) of wt are zero -- the backward-data output may be wider than the forward input) */
int d2t_conv_pack_weights_f16_dgrad(const float* w_oihw, const float* scale, int Cout, int Cin, int rows, int R, int S,
                                    int cout_pad, const float* amax_wt, void* wt_hi, void* wt_lo, cudaStream_t stream);
/* out[n,y,x,:] = mask > 0 ? (y, x even ? low[n,y/2,x/2,:] : 0) + extra[n,y,x,:] : 0  (NHWC, C % 4 == 0, one channel
 * stride = C; extra / mask / amax_out may be NULL); max |out| is folded into *amax_out */
int d2t_upsample2_add_mask(const float* low, int LH, int LW, const float* extra, const float* mask, int N, int H,
                           int W, int C, float* out, float* amax_out, cudaStream_t stream);
/* Weight gradient  dw[co][ci][r][s] = scale[co] * sum_{n,oy,ox} x[n, ci, oy + r*dil - pad, ox + s*dil - pad] * g[n, co, oy, ox]
 * as a GEMM with K over the output pixels: both operands are read from channel-major PLANES so that K is contiguous.
 * d2t_wgrad_pack_input: channels [0, C) of the NHWC input sampled at (oy*stride, ox*stride) -> fp32 planes
 * [N][C][OH][pitch] (a stride-2 1x1 convolution is a stride-1 one on the sampled positions); d2t_wgrad_pack_grad: the
 * NHWC gradient -> S copies [S][N][C][OH][pitch] of fp16 planes (hi, lo) of g * 2^k (k from *amax_g), copy s shifted
 * right by s*dil - pad columns: a TMA box must start on a 16-byte boundary of the contiguous axis, so the filter column's
 * offset is materialised by the packer (the row offset is a box coordinate).  Pitches in elements (% 4 / % 8). */
int d2t_wgrad_pack_input(const float* x, int N, int H, int W, int c_stride, int C, int stride, int OH, int OW,
                         int pitch, float* xt, cudaStream_t stream);
int d2t_wgrad_pack_grad(const float* g, int N, int OH, int OW, int c_stride, int C, int pitch, int S, int dil, int pad,
                        const float* amax_g, void* g_hi, void* g_lo, cudaStream_t stream);
d2t_conv_plan* d2t_wgrad_plan_create(int N, int Cin, int Cout, int xh, int xw, int xt_pitch, int OH, int OW,
                                     int g_pitch, int R, int S, int pad, int dil, const float* xt,
                                     const void* g_hi, const void* g_lo, const float* amax_x,
                                     const float* amax_g, const float* scale, float* dw);
/* Split-K reduction of a weight-gradient plan as a separate device-wide launch.  The GEMM has few tiles and K = every output
 * pixel, so each tile is shared by ~9 CTAs; with a partials buffer (d2t_wgrad_partials_bytes() bytes of device memory, may be
 * shared by all weight-gradient plans of a stream) every CTA parks its partial tiles there and d2t_conv_plan_run launches
 * wgrad_reduce right after the GEMM, which adds them in CTA order (deterministic) -- instead of one finisher CTA per tile
 * reading 8 partial tiles one after the other.  NULL restores the in-kernel finisher. */
size_t d2t_wgrad_partials_bytes(void);
int d2t_wgrad_plan_set_partials(d2t_conv_plan* plan, void* partials, size_t bytes);

/* Correlation backward on the tensor cores (replaces Correlation_backward_input1 / _input2,
 * correlation/src/correlation_cuda_kernel.cu:108-290, for kernel_size 1, stride1 == stride2, pad == max_displacement):
 *   g1[n, y, x, c] = (1/C) sum_t gO[n, y, x, t] * in2[n, (y, x) + t, c]          (lattice coordinates; t = (tj, ti), |.| <= r)
 *   g2[n, y, x, c] = (1/C) sum_t gO[n, (y, x) - t, t] * in1[n, (y, x) - t, c]
 * as the GEMM  [positions] x [halo positions] x [channels]  with a banded A operand: d2t_corrb_pack_band expands gO
 * (channels [c_offset, c_offset + D*D) of an NHWC buffer; flipped = 1 builds the band of g2) to fp16 (hi, lo)
 * [N][H][W][D][64]; d2t_corrb_pack_other writes the other frame's features, sampled on the correlation lattice, as
 * fp16 (hi, lo) planes [N][C][OH][pitch]; the plan's output is NHWC [N, OH, OW, out_cstride] channels
 * [out_coffset, out_coffset + C), every element written, multiplied by scale[c] (pass 1/C). */
int d2t_corrb_pack_band(const float* g, int N, int H, int W, int c_stride, int c_offset, int r, int flipped,
                        const float* amax_g, void* e_hi, void* e_lo, cudaStream_t stream);
int d2t_corrb_pack_other(const float* x, int N, int H, int W, int c_stride, int C, int stride, int OH, int OW,
                         int pitch, const float* amax_x, void* o_hi, void* o_lo, cudaStream_t stream);
d2t_conv_plan* d2t_corrb_plan_create(int N, int C, int H, int W, int r, const void* e_hi, const void* e_lo,
                                     const void* o_hi, const void* o_lo, int o_pitch, const float* amax_e,
                                     const float* amax_o, const float* scale, float* out, int out_cstride,
                                     int out_coffset);

/* OIHW fp32 -> [Cout][R*S][cin_pad] (w, w_lo); w_lo may be NULL */
int d2t_conv_pack_weights(const float* w_oihw, int Cout, int Cin, int R, int S, int cin_pad,
                          float* w_hi, float* w_lo, cudaStream_t stream);
/* OIHW fp32 -> [Cout][R*S][cin_pad] fp16 pair (hi, lo) of w * scale[cout] * 2^w_exp, cin_pad a multiple of 64:
 * hi = fp16(.), lo = fp16(. - hi).  scale (may be NULL) is the per-output-channel factor of the convolution -- the folded
 * BatchNorm scale: a 3xFP16 plan has no epilogue scale, it lives in the packed weights.  Choose w_exp so that
 * max |w * scale| * 2^w_exp < 2^15. */
int d2t_conv_pack_weights_f16(const float* w_oihw, const float* scale, int Cout, int Cin, int R, int S, int cin_pad,
                              int w_exp, void* w_hi, void* w_lo, cudaStream_t stream);
/* plain fp32 NCHW -> channels [c_offset, c_offset + c_width) of an NHWC tensor with c_stride channels
 * per pixel (the C source channels, then zeros); and back */
int d2t_nchw_to_nhwc(const float* x, int N, int C, int H, int W, int c_stride, int c_offset, int c_width,
                     float* out, cudaStream_t stream);
/* d2t_nchw_to_nhwc with the tensor's max |x| folded into *amax (device float, zeroed by the caller; NaN-free inputs): the
 * operand scale of a 3xFP16 consumer without a separate reduction pass. */
int d2t_nchw_to_nhwc_amax(const float* x, int N, int C, int H, int W, int c_stride, int c_offset, int c_width,
                          float* out, float* amax, cudaStream_t stream);
int d2t_nhwc_to_nchw(const float* x, int N, int C, int H, int W, int c_stride, int c_offset, float* out,
                     cudaStream_t stream);
/* MaxPool2d(3, stride 2, padding 0, ceil_mode=True) on NHWC (faster_rcnn/resnet.py:120) */
int d2t_maxpool3x3s2_nhwc(const float* in, int N, int H, int W, int C, float* out, cudaStream_t stream);

/* ---- Frame preparation on the device (the data format ahead of the path) ----
 * Replaces the reference's host chain for one batch of equally sized uint8 BGR frames (what cv2.imread returns):
 *   prep_im_for_blob: float32 cast, PIXEL_MEANS subtraction, cv2.resize(fx = fy = im_scale, INTER_LINEAR)
 *                                                                          (lib/model/utils/blob.py:35-52),
 *   the horizontal flip of flipped roidb entries                           (lib/roi_data_layer/minibatch.py:77-78),
 *   im_list_to_blob's zero padding                                         (lib/model/utils/blob.py:20-33),
 *   the loader's permute(0, 3, 1, 2)                                       (lib/roi_data_layer/roibatchLoader.py:183).
 * d2t_frames_resized_shape: im_scale = target_size / min(h, w), capped so that round(im_scale * max(h, w)) <= max_size
 * when cap != 0 (demo.py:270-274, online_tubes.py:656-659; minibatch.py passes through blob.py, where the cap is
 * commented out: cap = 0); out_hw = (round(h * im_scale), round(w * im_scale)), half to even as cv::resize rounds.
 * d2t_frames_prep: frames [n, src_h, src_w, 3] uint8 (device) -> blob float32 (device), [n, 3, blob_h, blob_w] (nhwc = 0)
 * or [n, blob_h, blob_w, 3] (nhwc = 1); rows >= dst_h and columns >= dst_w are written as zeros (the blob need not be
 * cleared); pixel_means = 3 host doubles (B, G, R; config.py:257).  Arithmetic: OpenCV's own float32 bilinear code
 * (coordinates in double, two fp32 passes) -- bit-identical to oracle/frames.py.  Returns 1, or 0 + d2t_last_error(). */
int d2t_frames_resized_shape(int src_h, int src_w, int target_size, int max_size, int cap, int* out_hw,
                             double* im_scale);
int d2t_frames_prep(const uint8_t* frames, int n, int src_h, int src_w, const double* pixel_means, double im_scale,
                    int flipped, int dst_h, int dst_w, float* blob, int blob_h, int blob_w, int nhwc,
                    cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#endif /* D2T_B200_H */
