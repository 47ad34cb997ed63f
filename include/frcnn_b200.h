/*
 * frcnn_b200.h -- C ABI of libfrcnn_b200.so: the B200 (sm_100a) implementation of the
 * region-proposal / RoI hot path of antoineBarbez/2D_object_detection.
 *
 * The reference has no FFI of its own (it is pure Python on TensorFlow 2); the entry points
 * below are what a binding for this path replaces, one per reference callable.  Each cites the
 * reference interface (file:line under the reference repository) whose results it reproduces.
 * INTEGRATION.md shows the reference-side stubs.
 *
 * Conventions
 *  - Plain pointers and sizes only.  Unless a parameter is marked HOST, every pointer is a DEVICE
 *    pointer owned by the caller, fp32 C-contiguous, 16-byte aligned.
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Functions only
 *    enqueue work on it: no allocation, no device synchronisation, re-entrant per stream.
 *    (Exception: the frcnn_pipeline_* host-buffer API owns its device buffers and synchronises.)
 *  - Scratch memory comes from the caller: frcnn_<op>_workspace_bytes() then `workspace`.
 *  - Return value: 0 on success; a positive cudaError_t, or a negative FRCNN_E* code.
 *    frcnn_last_error() returns a thread-local description of the last failure.
 *  - Box layout everywhere: [x_min, y_min, x_max, y_max]; "abs" = pixels, "rel" = divided by W/H.
 *  - All arithmetic is fp32 in the reference's operation order, compiled without FMA contraction.
 */
#ifndef FRCNN_B200_H_
#define FRCNN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FRCNN_VERSION 100 /* 0.1.0 */

#define FRCNN_EINVAL (-1)    /* bad argument (shape, alignment, limit) */
#define FRCNN_EWORKSPACE (-2) /* workspace too small */
#define FRCNN_ELIMIT (-3)    /* size beyond a documented limit of this build */

typedef void* frcnn_stream_t;

#if defined(__GNUC__)
#define FRCNN_API __attribute__((visibility("default")))
#else
#define FRCNN_API
#endif

FRCNN_API int frcnn_version(void);
FRCNN_API const char* frcnn_last_error(void);

/* ---- utils/boxes.py ---------------------------------------------------------------------- */

/* clip_to_window(boxes, window)  utils/boxes.py:4-17; window = x_min,y_min,x_max,y_max (:13). */
FRCNN_API int frcnn_clip_to_window(const float* boxes, int64_t n_boxes, float win_x_min, float win_y_min, float win_x_max,
                         float win_y_max, float* out, frcnn_stream_t stream);

/* decode(boxes, reference_boxes)  utils/boxes.py:20-41.  Reference box of element i is
 * refs[(i / ref_inner) % ref_period] (covers [N,4] regions tiled over batch and classes). */
FRCNN_API int frcnn_decode(const float* deltas, const float* refs, int64_t n_boxes, int64_t ref_inner, int64_t ref_period,
                 float* out, frcnn_stream_t stream);

/* encode(boxes, reference_boxes)  utils/boxes.py:44-73. Same reference indexing as frcnn_decode. */
FRCNN_API int frcnn_encode(const float* boxes, const float* refs, int64_t n_boxes, int64_t ref_inner, int64_t ref_period,
                 float* out, frcnn_stream_t stream);

/* to_absolute (divide = 0, multiplies) / to_relative (divide = 1, true division) utils/boxes.py:76-93. */
FRCNN_API int frcnn_scale_boxes(const float* boxes, int64_t n_boxes, float image_height, float image_width, int divide,
                      float* out, frcnn_stream_t stream);

/* ---- models/detectors/rpn_detector.py ------------------------------------------------------ */

/* RPNDetector._generate_anchors  rpn_detector.py:162-199.  scales/aspect_ratios are HOST arrays
 * (ns * nr <= 64).  anchors_out: [Hf*Wf*ns*nr, 4] abs, index (i*Wf + j)*A + (r*ns + s). */
FRCNN_API int frcnn_generate_anchors(int grid_h, int grid_w, const float* scales /*HOST*/, int ns,
                           const float* aspect_ratios /*HOST*/, int nr, float base_h, float base_w, float stride_h,
                           float stride_w, float* anchors_out, frcnn_stream_t stream);

/* RPNDetector._remove_invalid_anchors_and_predictions index list  rpn_detector.py:215-224:
 * ascending indices of anchors with x0>=0, y0>=0, x1<=W, y1<=H.  idx_out [n] int64 (first *count_out
 * entries valid), count_out [1] int32, both device. */
FRCNN_API int frcnn_valid_anchor_indices(const float* anchors, int64_t n, float image_height, float image_width,
                               int64_t* idx_out, int32_t* count_out, frcnn_stream_t stream);

/* gather rows: out[b, m, :] = in[b, idx[m], :] (tf.gather(axis=1) at rpn_detector.py:226-228);
 * row_floats floats per row; idx int64 [m] device.  batch = 1 gathers the anchors. */
FRCNN_API int frcnn_gather_rows(const float* in, int64_t batch, int64_t n, int64_t row_floats, const int64_t* idx, int64_t m,
                      float* out, frcnn_stream_t stream);

/* ---- utils/metrics.py:136-208 ------------------------------------------------------------ */
FRCNN_API int frcnn_area(const float* boxes, int64_t n, float* out, frcnn_stream_t stream);
FRCNN_API int frcnn_intersection_pairwise(const float* b1, int64_t n, const float* b2, int64_t m, float* out,
                                frcnn_stream_t stream);
FRCNN_API int frcnn_iou_pairwise(const float* b1, int64_t n, const float* b2, int64_t m, float* out, frcnn_stream_t stream);
FRCNN_API int frcnn_iou_elementwise(const float* b1, const float* b2, int64_t n, float* out, frcnn_stream_t stream);
FRCNN_API int frcnn_intersection_elementwise(const float* b1, const float* b2, int64_t n, float* out, frcnn_stream_t stream);

/* ---- utils/post_processing.py:6-63 postprocess_output ---------------------------------------
 * decode -> to_relative -> drop background column -> tf.image.combined_non_max_suppression
 * (pad_per_class=False, clip_boxes=True) -> zero-padded outputs.
 *   regions      [N,4] (regions_batched=0) or [B,N,4] (=1), abs
 *   pred_scores  [B,N,nc+1]      pred_boxes [B,N,nc,4] encoded
 *   out_boxes [B,R,4] rel clipped to [0,1]; out_scores [B,R]; out_classes [B,R] int32;
 *   out_valid [B] int32; out_idx [B,R] int32 region index of each detection, -1 padded (nullable);
 *   out_exhausted [B] int32 (nullable): 1 if pre_nms_top_n cut the candidate list before
 *   max_output_size_per_class boxes were kept in some class (result then differs from the reference,
 *   which has no pre-NMS cap: SURVEY trap T1).  pre_nms_top_n <= 0 means "all" = reference behaviour.
 * Score ties are visited lower index first; merged detections tie-break by class then rank.
 * Limits: min(max_per_class, N) <= 4096, nc * min(max_per_class, N) <= 16384, N < 2^22. */
FRCNN_API size_t frcnn_postprocess_workspace_bytes(int B, int N, int nc, int max_per_class, int max_total);
FRCNN_API int frcnn_postprocess(float image_height, float image_width, const float* regions, int regions_batched,
                      const float* pred_scores, const float* pred_boxes, int B, int N, int nc, float score_threshold,
                      float iou_threshold, int max_per_class, int max_total, int pre_nms_top_n, float* out_boxes,
                      float* out_scores, int32_t* out_classes, int32_t* out_valid, int32_t* out_idx,
                      int32_t* out_exhausted, void* workspace, size_t workspace_bytes, frcnn_stream_t stream);

/* The same with the classification head's Softmax activation fused in (models/detectors/rpn_detector.py:81-84:
 * 1x1 conv -> Reshape((-1, 2)) -> Softmax; fast_rcnn_detector.py:60-61 likewise): pred_logits [B,N,nc+1] are the
 * reshaped pre-activation outputs, p_j = exp(l_j - max) / sum_k exp(l_k - max) (sum left to right, IEEE division)
 * is computed in the decode kernel, so the probabilities are thresholded without a round trip through memory.
 * softmax_out [B,N,nc+1] (nullable) receives the probabilities (the "pred_scores" the detector returns and the
 * losses read).  Selection is bit-identical to frcnn_postprocess called on softmax_out. */
FRCNN_API int frcnn_postprocess_logits(float image_height, float image_width, const float* regions, int regions_batched,
                             const float* pred_logits, float* softmax_out, const float* pred_boxes, int B, int N, int nc,
                             float score_threshold, float iou_threshold, int max_per_class, int max_total,
                             int pre_nms_top_n, float* out_boxes, float* out_scores, int32_t* out_classes,
                             int32_t* out_valid, int32_t* out_idx, int32_t* out_exhausted, void* workspace,
                             size_t workspace_bytes, frcnn_stream_t stream);

/* Stage 1 of the above (exposed for stage-level parity and profiling): fused decode + to_relative
 * + background-column drop + score threshold.  boxes_rel [B,nc,N,4] (class-major, unclipped: trap T2),
 * keys [B,nc,N] uint32 = order-preserving score key, 0 for score <= threshold / NaN. */
FRCNN_API int frcnn_decode_filter(float image_height, float image_width, const float* regions, int regions_batched,
                        const float* pred_scores, const float* pred_boxes, int B, int N, int nc,
                        float score_threshold, float* boxes_rel, uint32_t* keys, frcnn_stream_t stream);

/* Stage 2: segmented greedy NMS in TF order on explicit boxes.  segments = independent (image,class)
 * problems; boxes [segments,N,4], scores [segments,N]; kept_idx [segments,cap] int32 in keep order
 * (-1 padded), kept_count [segments] int32, exhausted [segments] int32 (nullable).
 * cap = min(max_output, N). */
FRCNN_API size_t frcnn_nms_workspace_bytes(int segments, int N);
FRCNN_API int frcnn_nms(const float* boxes, const float* scores, int segments, int N, float score_threshold,
              float iou_threshold, int max_output, int pre_nms_top_n, int32_t* kept_idx, int32_t* kept_count,
              int32_t* exhausted, void* workspace, size_t workspace_bytes, frcnn_stream_t stream);

/* ---- models/detectors/fast_rcnn_detector.py:133-177 ROIPooling ------------------------------
 * tf.image.crop_and_resize to (P*k)x(P*k) (bilinear, (H-1) scaling, extrapolation 0) fused with
 * MaxPool2D(k) (stride k, VALID).  feature_maps [B,Hf,Wf,C] NHWC; rois [B,R,4] rel x1,y1,x2,y2;
 * out [B,R,P,P,C] (== [B,R,P*P*C] flattened in (py,px,c) order).  P*k <= 32.
 * The reference configuration (pooled_size 7, kernel_size 2, C % 512 == 0) runs the row-streaming TMA kernel
 * (roi_pool_stream.cu); it builds CUtensorMap descriptors of feature_maps on the host (cached per pointer and shape,
 * cuTensorMapEncodeTiled through cudaGetDriverEntryPoint), so feature_maps must stay valid until the stream has run.
 * Other shapes run the bin-row walk / generic kernels (roi_pool.cu).  All variants give identical bits.
 * PRECONDITION: feature_maps (and dy in the backward) hold FINITE values.  At an integer sample position TF reads
 * ceil(in) == floor(in) for the right / bottom neighbour; the walk and streaming kernels read the pixel after it with
 * weight exactly 0 -- (next - a) * 0 == 0 for finite data, NaN if that extra pixel is Inf / NaN
 * (tests: test_roi_pool_integer_positions_read_one_pixel_further_finite_inputs). */
/* Scratch for the row-streaming kernels: a work counter and one 512-byte record per RoI (built by a pre-kernel on the same
 * stream: sample offsets / weights, the feature rows of the walk, ring geometry).  0 for shapes that do not use them.
 * Launches that may overlap (different streams) need different workspaces. */
FRCNN_API size_t frcnn_roi_pool_workspace_bytes(int B, int R, int C, int pooled_size, int kernel_size);
/* The row-streaming kernels are persistent: one 16-warp CTA per SM, on every SM of the device by default.  A caller that
 * overlaps them with kernels of other streams (e.g. the proposal path of the next batch, whose 1024-thread NMS CTAs need
 * whole SMs) caps the SMs they occupy here; max_sms <= 0 = all SMs.  Process-wide setting, read at launch. */
FRCNN_API int frcnn_roi_pool_set_max_sms(int max_sms);
FRCNN_API int frcnn_roi_pool_forward(const float* feature_maps, int B, int Hf, int Wf, int C, const float* rois, int R,
                           int pooled_size, int kernel_size, float* out, void* workspace, size_t workspace_bytes,
                           frcnn_stream_t stream);

/* The same call in two parts, for callers that pipeline batches (engine.ProposalRoIEngine): frcnn_roi_pool_prepare builds
 * the per-RoI records of `rois` in the workspace (the small pre-kernel; it belongs with the proposal path, which produces
 * the rois -- fast_rcnn_detector.py:154-158 reads them once per call) and arms the workspace for ONE forward (the backward
 * re-arms its own work counter, so it may run any number of times on the same records); frcnn_roi_pool_forward_prepared / _backward_prepared then launch only the pooling kernel (argmax_codes of the
 * forward may be NULL: inference).  The caller orders the streams (prepare before the pooling launch) and keeps the
 * workspace untouched in between.  Row-streaming configuration only (frcnn_roi_pool_workspace_bytes(...) > 0), else
 * FRCNN_EINVAL.  Same bits as frcnn_roi_pool_forward / _forward_argmax / _backward_argmax. */
FRCNN_API int frcnn_roi_pool_prepare(const float* rois, int B, int R, int Hf, int Wf, int C, int pooled_size, int kernel_size,
                           void* workspace, size_t workspace_bytes, frcnn_stream_t stream);
FRCNN_API int frcnn_roi_pool_forward_prepared(const float* feature_maps, int B, int Hf, int Wf, int C, int R, int pooled_size,
                                    int kernel_size, float* out, unsigned char* argmax_codes, void* workspace,
                                    size_t workspace_bytes, frcnn_stream_t stream);
FRCNN_API int frcnn_roi_pool_backward_prepared(int B, int Hf, int Wf, int C, int R, int pooled_size, int kernel_size,
                                     const float* dy, const unsigned char* argmax_codes, float* dfeature, void* workspace,
                                     size_t workspace_bytes, frcnn_stream_t stream);

/* Gradient w.r.t. feature_maps (models/faster_rcnn.py:73,103 through :168 and :160): MaxPoolGrad to
 * the first maximum in row-major window order, then CropAndResizeGradImage.  dfeature [B,Hf,Wf,C] is
 * overwritten (zero-initialised here).  No box gradient (post_processing.py:58 stop_gradient). */
FRCNN_API int frcnn_roi_pool_backward(const float* feature_maps, int B, int Hf, int Wf, int C, const float* rois, int R,
                            int pooled_size, int kernel_size, const float* dy, float* dfeature,
                            frcnn_stream_t stream);

/* Training-mode pair: the forward additionally writes the per-channel arg-max of every 2x2 window as 2-bit codes
 * (argmax_codes [B,R,P,P,C/4] bytes, four codes per float4 of channels), the backward consumes them and needs neither
 * the feature maps nor a second interpolation pass.  Same results as the two functions above.
 * Supported when frcnn_roi_pool_argmax_supported returns 1: C % 4 == 0, kernel_size == 2, pooled_size <= 8.
 * In the reference configuration both run on the row-streaming walk; the backward adds whole feature-row segments to
 * dfeature with TMA reduce-add (cp.reduce.async.bulk.tensor), so dfeature needs the usual 16-byte alignment only. */
FRCNN_API int frcnn_roi_pool_argmax_supported(int C, int pooled_size, int kernel_size);
FRCNN_API int frcnn_roi_pool_forward_argmax(const float* feature_maps, int B, int Hf, int Wf, int C, const float* rois, int R,
                                  int pooled_size, int kernel_size, float* out, unsigned char* argmax_codes,
                                  void* workspace, size_t workspace_bytes, frcnn_stream_t stream);
FRCNN_API int frcnn_roi_pool_backward_argmax(int B, int Hf, int Wf, int C, const float* rois, int R, int pooled_size,
                                   int kernel_size, const float* dy, const unsigned char* argmax_codes, float* dfeature,
                                   void* workspace, size_t workspace_bytes, frcnn_stream_t stream);

/* ---- utils/training.py:7-77 generate_targets (+ :123-143), batched over images -------------
 *   gt_labels [B,G,nc+1] one-hot, zero rows = padding; gt_boxes [B,G,4] rel;
 *   regions [N,4] (regions_batched=0) or [B,N,4], abs.
 *   target_labels [B,N,nc+1]; target_boxes [B,N,nc,4]; max_iou [B,N], argmax [B,N] int32 (nullable;
 *   argmax indexes the gt list after padding removal, as the reference's tf.gather sees it).
 * An image with no gt after padding removal (undefined in the reference) gets max_iou 0,
 * no forced foreground and zero box targets.  G <= 1024. */
FRCNN_API size_t frcnn_generate_targets_workspace_bytes(int B, int N);
FRCNN_API int frcnn_generate_targets(const float* gt_labels, const float* gt_boxes, int B, int G, int nc, const float* regions,
                           int regions_batched, int N, float image_height, float image_width, float fg_lo,
                           float fg_hi, float bg_lo, float bg_hi, float* target_labels, float* target_boxes,
                           float* max_iou, int32_t* argmax, void* workspace, size_t workspace_bytes,
                           frcnn_stream_t stream);

/* The RPN's form (models/detectors/rpn_detector.py:141-149): the objectness labels
 *   one_hot(cast(reduce_sum(gt_labels, -1), int32), depth=2)
 * are computed inside the matching kernels from gt_labels [B,G,num_classes_plus_1], then generate_targets runs on them:
 * target_labels [B,N,2], target_boxes [B,N,1,4].  Padding rows become [1, 0] and are NOT removed (trap T11), exactly as in
 * the reference.  frcnn_objectness_labels writes those labels on their own ([rows,num_classes_plus_1] -> [rows,2]). */
FRCNN_API int frcnn_generate_targets_objectness(const float* gt_labels, const float* gt_boxes, int B, int G,
                                      int num_classes_plus_1, const float* regions, int regions_batched, int N,
                                      float image_height, float image_width, float fg_lo, float fg_hi, float bg_lo,
                                      float bg_hi, float* target_labels, float* target_boxes, float* max_iou,
                                      int32_t* argmax, void* workspace, size_t workspace_bytes, frcnn_stream_t stream);
FRCNN_API int frcnn_objectness_labels(const float* gt_labels, int64_t rows, int num_classes_plus_1, float* out,
                            frcnn_stream_t stream);

/* ---- utils/metrics.py:57-84 AveragePrecision.update_state_single_image, batched ------------------------------------
 *   gt_boxes [B,G,4] (rows summing to 0 = padding), pred_boxes [B,M,4]; true_positives [B,M] int32: 1 where the
 *   prediction is the highest-IoU prediction (first maximum) of at least one ground-truth box and that IoU exceeds
 *   iou_threshold; num_gt [B] int32 = ground-truth boxes after padding removal (nullable).  IoU as in frcnn_iou_*. */
FRCNN_API int frcnn_ap_match(const float* gt_boxes, int B, int G, const float* pred_boxes, int M, float iou_threshold,
                   int32_t* true_positives, int32_t* num_gt, frcnn_stream_t stream);

/* ---- utils/losses.py:4-43 on the sampled rows (models/faster_rcnn.py:87-101), forward + gradients ------------------
 *   ClassificationLoss: Keras CategoricalCrossentropy on probabilities (normalise by the row sum, clip to
 *   [1e-7, 1 - 1e-7], -sum t log c, mean over the `rows` rows); target_labels / pred_scores [rows, nc+1].
 *   RegressionLoss: Huber(delta 1) per coordinate, mean over the 4 coordinates, SUMMED over the rows whose target box is
 *   non-zero; target_boxes / pred_boxes [box_rows, 4].
 *   losses[0] = classification, losses[1] = regression, losses[2] = number of regression rows kept (device, 3 floats).
 *   d_pred_scores [rows, nc+1] / d_pred_boxes [box_rows, 4] (nullable) receive grad_cls * dL_cls/dp and
 *   grad_reg * dL_reg/dp.  One CTA, fixed summation order (deterministic); 1e-5 relative vs the reference's reductions. */
FRCNN_API int frcnn_sampled_losses(const float* target_labels, const float* pred_scores, int64_t rows, int num_classes_plus_1,
                         const float* target_boxes, const float* pred_boxes, int64_t box_rows, float grad_cls,
                         float grad_reg, float* losses, float* d_pred_scores, float* d_pred_boxes,
                         frcnn_stream_t stream);

/* ---- utils/training.py:80-120 get_sample_indices, batched, fused with the tf.gather(batch_dims=1) calls that follow
 * it (models/detectors/rpn_detector.py:151-160, fast_rcnn_detector.py:121-130) --------------------------------------
 *   target_labels [B,N,nc+1].  foreground = non-zero label row with background column 0, background = non-zero row
 *   with background column 1 (both in ascending index order, as tf.where lists them).
 *   n_fg = min(#foreground, round(num_samples * foreground_proportion)); n_fg foregrounds without replacement
 *   (shuffle + prefix), num_samples - n_fg backgrounds with replacement; indices [B,num_samples] int64 = fg then bg.
 *   The reference draws from TensorFlow's unseeded generator: parity is distributional.  Here every draw is a pure
 *   function of (seed, image, position) through Philox4x32-10, so a call is reproducible.
 *   counts [B,2] int32 = (#foreground, #background) per image (nullable).  An image that needs a background draw but has
 *   no background region gets index -1 there (the reference raises); callers check counts.
 *   Optional gathers (any of the four outputs may be NULL): out_labels / out_scores [B,num_samples,nc+1] from
 *   target_labels / pred_scores, out_target_boxes / out_pred_boxes [B,num_samples,nc,4] from target_boxes / pred_boxes. */
FRCNN_API size_t frcnn_sample_indices_workspace_bytes(int B, int N);
FRCNN_API int frcnn_sample_and_gather(const float* target_labels, int B, int N, int num_classes_plus_1, int num_samples,
                            double foreground_proportion, uint64_t seed, int64_t* indices, int32_t* counts,
                            const float* pred_scores, const float* target_boxes, const float* pred_boxes,
                            float* out_labels, float* out_scores, float* out_target_boxes, float* out_pred_boxes,
                            void* workspace, size_t workspace_bytes, frcnn_stream_t stream);

/* ---- host-buffer pipeline: FasterRCNN.call inference slice, models/faster_rcnn.py:52-55 ------
 * postprocess_output(rpn) -> ROIPooling on the proposals, taking and returning HOST buffers
 * (pinned memory recommended).  Owns its device buffers and streams; run() copies inputs host->device,
 * runs the kernels, copies the results device->host and returns after they have landed. */
typedef struct frcnn_pipeline frcnn_pipeline_t;

FRCNN_API int frcnn_pipeline_create(frcnn_pipeline_t** out, int B, int Hf, int Wf, int C, int N, int max_total,
                          int pooled_size, int kernel_size);
FRCNN_API void frcnn_pipeline_destroy(frcnn_pipeline_t* p);
/* regions [N,4] abs HOST; pred_scores [B,N,2] HOST; pred_boxes [B,N,1,4] HOST; feature_maps [B,Hf,Wf,C] HOST.
 * Outputs (HOST): rois [B,R,4], roi_scores [B,R], num_valid [B] int32, pooled [B,R,P,P,C] (nullable: skip copy). */
FRCNN_API int frcnn_pipeline_run(frcnn_pipeline_t* p, float image_height, float image_width, const float* regions,
                       const float* pred_scores, const float* pred_boxes, const float* feature_maps,
                       float score_threshold, float iou_threshold, int pre_nms_top_n, float* rois, float* roi_scores,
                       int32_t* num_valid, float* pooled);
/* The same, split: submit enqueues the copies and kernels of one batch and returns a ticket without waiting; wait blocks
 * until that batch's outputs have landed in the host buffers.  The device buffers exist twice, so batch i + 1 may be
 * submitted before batch i is waited for (at most two batches in flight; a third submit waits for the oldest): the
 * upstream copies of the next batch then overlap the downstream copies of the current one (PCIe is full duplex).  Input
 * and output host buffers must stay untouched between a submit and its wait.  run == submit + wait. */
FRCNN_API int frcnn_pipeline_submit(frcnn_pipeline_t* p, float image_height, float image_width, const float* regions,
                          const float* pred_scores, const float* pred_boxes, const float* feature_maps,
                          float score_threshold, float iou_threshold, int pre_nms_top_n, float* rois, float* roi_scores,
                          int32_t* num_valid, float* pooled, int32_t* ticket);
FRCNN_API int frcnn_pipeline_wait(frcnn_pipeline_t* p, int32_t ticket);

#ifdef __cplusplus
}
#endif
#endif /* FRCNN_B200_H_ */
