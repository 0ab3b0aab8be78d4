/*
 * ssdk.h -- C ABI of the B200-native per-anchor detection hot path ("single-shot-detector kernels").
 *
 * Drop-in boundary for TropComplique/single-shot-detector.  The reference has no FFI of its own
 * (it is pure TensorFlow-1.x Python); each entry point below names the reference function it
 * replaces (paths relative to the reference root).  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no C++/torch types.
 *   - Every function returns 0 (SSDK_OK) or a negative ssdk_status; ssdk_last_error() returns a
 *     thread-local, human readable message for the last failure on the calling thread.
 *   - All tensor pointers are DEVICE pointers on the context's device unless the function name
 *     ends in _host (then they are host pointers; pinned memory gives full PCIe speed).
 *     The caller owns every input and output buffer; the library keeps nothing after return
 *     except its private workspace inside the context.
 *   - Work is enqueued on the context's CUDA stream and is asynchronous w.r.t. the host unless
 *     noted ("synchronous").  One context per host thread per GPU; contexts are independent.
 *   - float = IEEE binary32, int = int32.  Boxes are [ymin, xmin, ymax, xmax], normalised.
 *   - `matches` values: >=0 index of the matched ground-truth box, -1 background, -2 ignore.
 *   - There is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef SSDK_H_
#define SSDK_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSDK_VERSION 100  /* 0.1.0 */

#if defined(__GNUC__)
#define SSDK_API __attribute__((visibility("default")))
#else
#define SSDK_API
#endif

typedef enum ssdk_status {
    SSDK_OK = 0,
    SSDK_ERR_ARG = -1,     /* bad argument value (also the reference's two asserts) */
    SSDK_ERR_SHAPE = -2,   /* shape / size / alignment not supported */
    SSDK_ERR_CUDA = -3,    /* CUDA runtime error */
    SSDK_ERR_NCCL = -4,    /* reserved: collective error */
    SSDK_ERR_NOMEM = -5    /* workspace allocation failed */
} ssdk_status;

typedef struct ssdk_ctx ssdk_ctx;

/* ---- library / context ------------------------------------------------------------------ */
SSDK_API int ssdk_version(void);
SSDK_API const char* ssdk_last_error(void);
/* stream: a cudaStream_t (NULL = legacy default stream).  Creates the context on `device`. */
SSDK_API int ssdk_ctx_create(int device, void* stream, ssdk_ctx** out);
SSDK_API int ssdk_ctx_set_stream(ssdk_ctx* ctx, void* stream);
/* Options of the fused training step (ssdk_ssd_loss_step, ssdk_ssd_targets_and_loss and their ssdk_head_* forms; see
 * csrc/train_step.cu): one persistent kernel whose CTAs take roles -- `SSDK_OPT_MATCH_CTAS_PER_SM` CTAs per SM (1..7, capped by
 * the kernel's occupancy) start with the ALU-bound target assignment while the others stream the logits, and join the
 * streaming afterwards for `SSDK_OPT_MATCH_FLAT_SHARE_PCT` percent (0..100) of a streaming CTA's share of the chunks.  The
 * defaults (0 and -1) choose both from the number of ground-truth boxes and classes (the ratio of matching to streaming work).
 * SSDK_OPT_FUSED_TRAIN_STEP = 0 runs the same computation as separate launches (matcher, flat pass, matched-anchor pass,
 * finalisation) -- for comparison; results agree to rounding of the double sums.  The same knobs are read once at context
 * creation from SSDK_MATCH_CTAS / SSDK_MATCH_FLAT_SHARE (values are clamped to their valid ranges). */
#define SSDK_OPT_FUSED_TRAIN_STEP 1
#define SSDK_OPT_MATCH_CTAS_PER_SM 2
#define SSDK_OPT_MATCH_FLAT_SHARE_PCT 3
/* SSDK_OPT_TRAIN_CTAS_PER_SM (default 0 = as many as fit, six): resident CTAs per SM of the fused training-step kernel.  It is a
 * persistent kernel that takes every CTA slot of the GPU; a caller that runs another sub-path on a second stream at the same time
 * (graph.concurrent) can leave room for it with a smaller value. */
#define SSDK_OPT_TRAIN_CTAS_PER_SM 4
/* SSDK_OPT_PROGRAMMATIC_LAUNCH (default 1; environment SSDK_PDL): the kernels of the post-processing chain (dense-image filter,
 * the NMS kernels, pack) are launched with programmatic dependent launch, so that each one's launch and prologue overlap its
 * predecessor's tail; 0 = plain stream order.  Same results either way. */
#define SSDK_OPT_PROGRAMMATIC_LAUNCH 5
/* SSDK_OPT_TRAIN_DYNAMIC_CHUNKS (default 1; environment SSDK_TRAIN_DYNAMIC): the streaming part of the fused training-step kernel
 * hands out its chunks dynamically (one atomic per four chunks) and adds the flat sum in 2^-32 fixed point, so that the result is
 * bit-identical from run to run whichever CTA took which chunk; the matcher CTAs join the streaming the moment they are done
 * (SSDK_OPT_MATCH_FLAT_SHARE_PCT is then unused), and CTAs that start late -- their SM was busy with another sub-path's kernel --
 * simply take fewer chunks.  0 = the static split (double accumulation in a fixed order); the two modes agree to ~1e-7 relative. */
#define SSDK_OPT_TRAIN_DYNAMIC_CHUNKS 6
SSDK_API int ssdk_ctx_set_option(ssdk_ctx* ctx, int option, int value);
SSDK_API int ssdk_ctx_destroy(ssdk_ctx* ctx);
/* Bytes of private workspace currently held (grows on demand, never shrinks). */
SSDK_API int64_t ssdk_ctx_workspace_bytes(const ssdk_ctx* ctx);
/* Counts kernel launches issued through this context since creation (bench.py's gpu_launches). */
SSDK_API int64_t ssdk_ctx_launch_count(const ssdk_ctx* ctx);
/* Optional per-kernel timing with CUDA events on the context's stream (for bench.py's roofline line; adds two
 * event records per kernel while enabled).  ssdk_ctx_profile_read synchronises the stream and returns, per kernel
 * id, the accumulated milliseconds and launch counts since the last reset.  Ids: 0 anchors, 1 match,
 * 2 force_match, 3 ssd_loss, 4 loss_reduce (unused), 5 filter, 6 filter_dense (images with dense scores, redone with per-tile aggregation),
 * 7 nms, 8 pack, 9 other, 10 ssd_loss_backward,
 * 11 head_flat (flat focal pass over the per-level head tensors), 12 head_rows (matched / ignored anchors), 13 head_concat,
 * 14 comm (peer-memory all-reduce), 15 train_step (the fused training step), 16 nms_rounds (dense / overflowing segments). */
#define SSDK_NUM_KERNEL_IDS 17
SSDK_API int ssdk_ctx_set_profiling(ssdk_ctx* ctx, int enable);
SSDK_API int ssdk_ctx_profile_read(ssdk_ctx* ctx, double* out_ms, int64_t* out_calls, int n, int reset);
/* Sticky asynchronous error word of the context (synchronises the stream first): 0, or the code of a condition a kernel met
 * that no status return could report.  SSDK_ASYNC_ROUNDS_TIMEOUT: the dense-segment stage of the post-processing
 * (nms_rounds_kernel, a persistent grid with grid-wide barriers) gave up after ~10 s because part of its grid was never
 * scheduled (the GPU was oversubscribed by other work for that long); the detections of that call are incomplete. */
#define SSDK_ASYNC_ROUNDS_TIMEOUT 1
SSDK_API int ssdk_ctx_async_error(ssdk_ctx* ctx, int* out_code);
/* Diagnostics of the dense-segment stage of the LAST post-processing call (synchronises): out[0] = rounds run (0 = no segment
 * overflowed its candidate region), out[1..n) = %globaltimer nanoseconds at kernel start, after initialisation, and per round at
 * its start and after the histogram pass, the planning step, the collect pass and the NMS phase (first rounds only). */
SSDK_API int ssdk_ctx_round_times(ssdk_ctx* ctx, int64_t* out, int n);
/* cudaStreamSynchronize on the context's stream (synchronous). */
SSDK_API int ssdk_ctx_synchronize(ssdk_ctx* ctx);

/* ---- anchors: detector/anchor_generator.py:40-120 (AnchorGenerator.__call__), :123-170 ---- */
/* Host-only arithmetic on shapes (no GPU needed): per-level counts h*w*per_location
 * (anchor_generator.py:59-62) and their sum. */
SSDK_API int ssdk_num_anchors(int image_height, int image_width, const int* strides, int num_levels,
                     int anchors_per_location, int64_t* out_total, int32_t* out_per_level);
/* scales: HOST float[num_levels * per_location] (= float32(multiplier * scale), :75);
 * ratios: HOST float[per_location] (:71); strides: HOST int[num_levels].
 * out_anchors: DEVICE float[A,4] normalised, unclipped (:110-118);
 * out_raw (may be NULL): DEVICE float[A,4] absolute pixel anchors, levels concatenated (:105). */
SSDK_API int ssdk_anchors(ssdk_ctx* ctx, int image_height, int image_width, const int* strides,
                 const float* scales, const float* ratios, int num_levels,
                 int anchors_per_location, float* out_anchors, float* out_raw);

/* ---- box utilities: detector/utils/box_utils.py ------------------------------------------ */
SSDK_API int ssdk_area(ssdk_ctx* ctx, const float* boxes, int64_t n, float* out);                 /* :53-61  */
SSDK_API int ssdk_intersection(ssdk_ctx* ctx, const float* boxes1, int64_t n, const float* boxes2,
                      int64_t m, float* out /*[n,m]*/);                                   /* :30-50  */
SSDK_API int ssdk_iou(ssdk_ctx* ctx, const float* boxes1, int64_t n, const float* boxes2, int64_t m,
             float* out /*[n,m]*/);                                                       /* :14-27  */
SSDK_API int ssdk_encode(ssdk_ctx* ctx, const float* boxes, const float* anchors, int64_t n,
                float* out /*[n,4]*/);                                                    /* :80-111 */
SSDK_API int ssdk_decode(ssdk_ctx* ctx, const float* codes, const float* anchors, int64_t n,
                float* out /*[n,4]*/);                                                    /* :114-142 */
/* codes [B,A,4], anchors [A,4] -> clip(decode, 0, 1) [B,A,4] */
SSDK_API int ssdk_batch_decode(ssdk_ctx* ctx, const float* codes, const float* anchors, int64_t B,
                      int64_t A, float* out);                                             /* :145-173 */

/* ---- target assignment: detector/training_target_creation.py ------------------------------ */
/* Batched match_boxes (:48-130).  gt_boxes [B,Gmax,4]; num_boxes DEVICE int[B] (NULL = every
 * image has Gmax boxes); images with 0 boxes get all -1 (get_training_targets :24-37).
 * Thresholds are doubles because the reference compares them as Python floats (:94).
 * Gmax <= 4096. */
SSDK_API int ssdk_match_boxes(ssdk_ctx* ctx, const float* anchors, int64_t A, const float* gt_boxes,
                     const int32_t* num_boxes, int B, int Gmax, double positives_threshold,
                     double negatives_threshold, int force_match_groundtruth,
                     int32_t* out_matches /*[B,A]*/);
/* create_targets (:133-176): reg [B,A,4], cls [B,A] (label+1, 0 = background). */
SSDK_API int ssdk_create_targets(ssdk_ctx* ctx, const float* anchors, int64_t A, const float* gt_boxes,
                        const int32_t* gt_labels, int B, int Gmax, const int32_t* matches,
                        float* out_reg, int32_t* out_cls);
/* get_training_targets (:5-45) for a batch == SSD._create_targets (detector/ssd.py:165-199):
 * matching (force match on) + targets in one pass. */
SSDK_API int ssdk_training_targets(ssdk_ctx* ctx, const float* anchors, int64_t A, const float* gt_boxes,
                          const int32_t* gt_labels, const int32_t* num_boxes, int B, int Gmax,
                          double positives_threshold, double negatives_threshold,
                          float* out_reg, int32_t* out_cls, int32_t* out_matches);

/* ssdk_training_targets that also delivers the number of matched anchors of this shard (matches >= 0; ssd.py:89,121-122)
 * in out_count (DEVICE double[1]) -- the loss normaliser's input, needed BEFORE the fused forward + backward pass; counted
 * inside the matching kernel, so no extra pass over `matches` (ssdk_count_matches) is needed. */
SSDK_API int ssdk_training_targets_count(ssdk_ctx* ctx, const float* anchors, int64_t A, const float* gt_boxes,
                                const int32_t* gt_labels, const int32_t* num_boxes, int B, int Gmax,
                                double positives_threshold, double negatives_threshold, float* out_reg,
                                int32_t* out_cls, int32_t* out_matches, double* out_count);

/* ---- losses: detector/losses.py ----------------------------------------------------------- */
/* localization_loss (:4-19): predictions/targets [B,A,4], weights [B,A] -> out [B,A]. */
SSDK_API int ssdk_localization_loss(ssdk_ctx* ctx, const float* predictions, const float* targets,
                           const float* weights, int64_t B, int64_t A, float* out);
/* focal_loss (:22-50) with dense one-hot float targets [B,A,C] exactly as the reference API. */
SSDK_API int ssdk_focal_loss(ssdk_ctx* ctx, const float* logits, const float* targets,
                    const float* weights, int64_t B, int64_t A, int C, double gamma, double alpha,
                    float* out /*[B,A]*/);

/* SSD.loss (detector/ssd.py:71-133) given the targets: focal loss with weights (matches >= -1)
 * on one_hot(cls)[..., 1:], smooth-L1 with weights (matches >= 0), matched count.
 *   logits [B,A,C] (16-byte aligned), codes [B,A,4], reg_targets [B,A,4], cls_targets [B,A],
 *   matches [B,A].
 *   out_sums: DEVICE double[3] = { sum(loc_losses), sum(cls_losses), num_matches } for THIS
 *             shard (un-normalised, so that ranks can all-reduce them; ssd.py:121-122,131-132).
 *   out_cls_losses / out_loc_losses: optional DEVICE float[B,A] per-anchor vectors (the tensors
 *             the reference's summaries consume, ssd.py:127-128); NULL to skip. */
SSDK_API int ssdk_ssd_loss(ssdk_ctx* ctx, const float* logits, const float* codes, const float* reg_targets,
                  const int32_t* cls_targets, const int32_t* matches, int64_t B, int64_t A, int C,
                  double gamma, double alpha, double* out_sums, float* out_cls_losses,
                  float* out_loc_losses);
/* Backward of SSD.loss (what the reference gets from TF autodiff at model.py:115-118 through ssd.py:89-133 and
 * losses.py:4-50; targets and weights are constants, ssd.py:197): gradients of
 *     upstream[0] * localization_loss + upstream[1] * classification_loss
 * w.r.t. logits [B,A,C] and codes [B,A,4].
 *   sums:     DEVICE double[3] as produced by ssdk_ssd_loss and, on several GPUs, all-reduced: only sums[2]
 *             (the GLOBAL num_matches) is read; normalizer = max(sums[2], 1) (ssd.py:123).
 *   upstream: DEVICE float[2] = { dL/d localization_loss, dL/d classification_loss }, e.g. the config's
 *             localization_loss_weight / classification_loss_weight (model.py:86-87); NULL = {1, 1}.
 *   grad_logits [B,A,C], grad_codes [B,A,4]: DEVICE outputs, every element is written. */
SSDK_API int ssdk_ssd_loss_backward(ssdk_ctx* ctx, const float* logits, const float* codes, const float* reg_targets,
                           const int32_t* cls_targets, const int32_t* matches, int64_t B, int64_t A, int C,
                           double gamma, double alpha, const double* sums, const float* upstream,
                           float* grad_logits, float* grad_codes);
/* Forward and backward of SSD.loss in ONE pass over the logits (a training step reads class_predictions once and
 * writes its gradient once).  The normaliser must therefore be known before the pass:
 *   num_matches: DEVICE double[1], the GLOBAL matched-anchor count (ssdk_count_matches on this shard's `matches`,
 *                all-reduced by the caller on several GPUs); normalizer = max(*num_matches, 1) (ssd.py:123).
 *   out_sums:    DEVICE double[3] = this shard's un-normalised { sum loc_losses, sum cls_losses, num_matches }, as
 *                ssdk_ssd_loss produces (all-reduce + ssdk_loss_finalize give the reported losses).
 * Other arguments as ssdk_ssd_loss_backward. */
SSDK_API int ssdk_ssd_loss_forward_backward(ssdk_ctx* ctx, const float* logits, const float* codes, const float* reg_targets,
                                   const int32_t* cls_targets, const int32_t* matches, int64_t B, int64_t A, int C,
                                   double gamma, double alpha, const double* num_matches, const float* upstream,
                                   double* out_sums, float* grad_logits, float* grad_codes);
/* out_count: DEVICE double[1] = number of entries of matches[n] that are >= 0 (ssd.py:89,121-122). */
SSDK_API int ssdk_count_matches(ssdk_ctx* ctx, const int32_t* matches, int64_t n, double* out_count);
/* normalizer = max(num_matches, 1) (ssd.py:123); out_losses: DEVICE float[2] =
 * { localization_loss, classification_loss } (ssd.py:133).  `sums` are the (all-reduced) sums. */
SSDK_API int ssdk_loss_finalize(ssdk_ctx* ctx, const double* sums, float* out_losses);

/* Whole training-side hot path for a batch of images resident in HBM:
 * targets (ssd.py:84) + losses (ssd.py:89-133).  Any of out_reg/out_cls/out_matches may be NULL,
 * in which case context workspace is used for them.  Without per-anchor outputs (out_cls_losses == out_loc_losses == NULL)
 * the loss is computed as a flat pass over the logits plus corrections for the matched / ignored anchors, by the fused
 * training-step kernel (one launch: matching CTAs and streaming CTAs side by side). */
SSDK_API int ssdk_ssd_targets_and_loss(ssdk_ctx* ctx, const float* anchors, const float* logits,
                              const float* codes, const float* gt_boxes, const int32_t* gt_labels,
                              const int32_t* num_boxes, int B, int64_t A, int C, int Gmax,
                              double positives_threshold, double negatives_threshold,
                              double gamma, double alpha, double* out_sums, float* out_reg,
                              int32_t* out_cls, int32_t* out_matches, float* out_cls_losses,
                              float* out_loc_losses);
/* SSD.loss (detector/ssd.py:71-133) for a batch in ONE launch: ssdk_ssd_targets_and_loss without per-anchor outputs, plus the
 * normalisation -- out_losses: DEVICE float[2] = { localization_loss, classification_loss } = sums / max(num_matches, 1)
 * (ssd.py:121-133).  flags & SSDK_STEP_ALL_REDUCE (the context must be connected, ssdk_comm_connect): the three sums are
 * all-reduced over the image shards of all ranks by the kernel's last CTA (NVLink peer memory), out_sums then holds the
 * GLOBAL sums and out_losses the global losses on every rank; a collective like ssdk_comm_all_reduce_sum. */
#define SSDK_STEP_ALL_REDUCE 1
SSDK_API int ssdk_ssd_loss_step(ssdk_ctx* ctx, const float* anchors, const float* logits, const float* codes,
                       const float* gt_boxes, const int32_t* gt_labels, const int32_t* num_boxes, int B, int64_t A, int C,
                       int Gmax, double positives_threshold, double negatives_threshold, double gamma, double alpha,
                       int flags, double* out_sums, float* out_losses, float* out_reg, int32_t* out_cls,
                       int32_t* out_matches);
/* Same call with HOST buffers (synchronous): copies inputs H2D, runs, copies
 * out_sums (double[3]) and out_losses (float[2], normalised with the LOCAL count) back. */
SSDK_API int ssdk_ssd_targets_and_loss_host(ssdk_ctx* ctx, const float* anchors, const float* logits,
                                   const float* codes, const float* gt_boxes,
                                   const int32_t* gt_labels, const int32_t* num_boxes, int B,
                                   int64_t A, int C, int Gmax, double positives_threshold,
                                   double negatives_threshold, double gamma, double alpha,
                                   double* out_sums, float* out_losses);

/* ---- post-processing: detector/utils/nms.py, detector/ssd.py:42-69 ------------------------ */
#define SSDK_INPUT_SCORES 0        /* `scores` are probabilities (batch_multiclass_nms, nms.py:48) */
#define SSDK_INPUT_LOGITS 1        /* `scores` are logits; sigmoid is fused (SSD.get_predictions, ssd.py:60) */
#define SSDK_BOXES_ENCODED 0       /* `codes` are box codes, decoded against `anchors` and clipped (nms.py:76-77) */
#define SSDK_BOXES_DECODED 2       /* `codes` are final boxes [B,A,4]; anchors ignored (multiclass_nms, nms.py:6) */
/* Split-phase post-processing (optional): a call with SSDK_POST_SCAN_ONLY enqueues only the HBM-bound part (counters zeroed, the
 * score scan, the dense-image filter) and writes no output; a following call with SSDK_POST_FINISH_ONLY and otherwise IDENTICAL
 * arguments, on the same context, with no other post-processing call in between, enqueues the rest (NMS stages, pack).  The two
 * calls may be issued on different streams (ssdk_ctx_set_stream) when the second stream waits for the first call's work: the
 * latency-bound second phase can then run next to another sub-path's streaming kernel.  Both flags clear = the whole chain. */
#define SSDK_POST_SCAN_ONLY 16
#define SSDK_POST_FINISH_ONLY 32
/* batch_multiclass_non_max_suppression (nms.py:48-102) with tf.image.non_max_suppression
 * (TF 1.12 NonMaxSuppressionV3) semantics per class: candidates score > score_threshold,
 * greedy in descending score (ties: lower anchor index), suppress iff IoU > iou_threshold,
 * at most K per class; class-major concatenation, zero padding to C*K, num_boxes per image.
 *   codes [B,A,4], anchors [A,4], scores [B,A,C];
 *   out_boxes [B,C*K,4], out_scores [B,C*K], out_classes int[B,C*K], out_num int[B].
 *   out_anchor_idx (may be NULL): int[B,C*K] anchor index of each kept box, -1 padding. */
SSDK_API int ssdk_postprocess(ssdk_ctx* ctx, const float* codes, const float* anchors, const float* scores,
                     int flags, int B, int64_t A, int C, double score_threshold,
                     double iou_threshold, int K, float* out_boxes, float* out_scores,
                     int32_t* out_classes, int32_t* out_num, int32_t* out_anchor_idx);
/* ssdk_postprocess plus the two consumers that directly follow it in the reference:
 *   box_scaler: DEVICE float[B,4] (16-byte aligned) or NULL -- boxes /= box_scaler[b] (model.py:67-68: undoes the
 *               resize / padding of the input pipeline), an IEEE division per coordinate;
 *   final_score_threshold: keeps only detections with score > it, order preserved, outputs re-packed and
 *               out_num updated (inference/detector.py:54-58: `to_keep = scores > score_threshold`); pass
 *               -INFINITY to keep everything. */
SSDK_API int ssdk_detect(ssdk_ctx* ctx, const float* codes, const float* anchors, const float* scores, int flags, int B,
                int64_t A, int C, double score_threshold, double iou_threshold, int K, const float* box_scaler,
                double final_score_threshold, float* out_boxes, float* out_scores, int32_t* out_classes,
                int32_t* out_num);
/* ssdk_detect plus the rows of the COCO results file that inference/evaluate_on_COCO.ipynb (cell 10) builds from the detector's
 * output, written by the same pack kernel: for detection i of image b (same order and padding as out_boxes)
 *     box * [height, width, height, width] (float32)  ->  out_bbox_xywh = [int(xmin), int(ymin), int(xmax - xmin), int(ymax - ymin)]
 * (Python int(): truncation towards zero), out_category_id = category_ids[class] (the notebook's integer_to_coco_id; NULL =
 * the class index), out_image_id = image_ids[b] (NULL = b); padding rows are 0 / -1 / -1.
 *   image_sizes: DEVICE float[B,2] = (height, width) in pixels; out_bbox_xywh: DEVICE int32[B,C*K,4] (16-byte aligned);
 *   out_category_id, out_image_id: DEVICE int32[B,C*K].  The notebook calls the detector with score_threshold 0.15: pass it
 *   as final_score_threshold. */
SSDK_API int ssdk_detect_coco(ssdk_ctx* ctx, const float* codes, const float* anchors, const float* scores, int flags, int B,
                     int64_t A, int C, double score_threshold, double iou_threshold, int K, const float* box_scaler,
                     double final_score_threshold, const float* image_sizes, const int32_t* image_ids,
                     const int32_t* category_ids, float* out_boxes, float* out_scores, int32_t* out_classes,
                     int32_t* out_num, int32_t* out_bbox_xywh, int32_t* out_category_id, int32_t* out_image_id);
/* The per-label detection lists the reference's evaluator accumulates (metrics.py:113-123: add_detections appends
 * get_box(box, image_name, score) to self.detections[label] for every detection of every image, in evaluation order): label-major
 * packing of the same NMS results.  For label c: out_counts[c] records; record j is out_boxes[c, j] / out_scores[c, j] /
 * out_image[c, j] (= image_ids[b], NULL = b): image 0's boxes of that class in descending score, then image 1's, ...
 *   out_boxes DEVICE float[C, B*K, 4] (16-byte aligned), out_scores float[C, B*K], out_image int32[C, B*K], out_counts int32[C];
 *   entries beyond out_counts[c] are not written. */
SSDK_API int ssdk_detect_by_label(ssdk_ctx* ctx, const float* codes, const float* anchors, const float* scores, int flags, int B,
                         int64_t A, int C, double score_threshold, double iou_threshold, int K, const float* box_scaler,
                         double final_score_threshold, const int32_t* image_ids, float* out_boxes, float* out_scores,
                         int32_t* out_image, int32_t* out_counts);
/* Same as ssdk_postprocess with HOST buffers (synchronous). */
SSDK_API int ssdk_postprocess_host(ssdk_ctx* ctx, const float* codes, const float* anchors,
                          const float* scores, int flags, int B, int64_t A, int C,
                          double score_threshold, double iou_threshold, int K, float* out_boxes,
                          float* out_scores, int32_t* out_classes, int32_t* out_num);

/* ---- head-layout fusion: detector/box_predictor.py:67-104 (reshape_and_concatenate) ---------- */
/* The box / class towers emit one tensor per FPN level in the network's data format (channels_first in the reference,
 * constants.py:9): class_predictions[l] is [B, n*C, h_l, w_l] and encoded_boxes[l] is [B, n*4, h_l, w_l], n =
 * num_anchors_per_location, channel = anchor_in_cell * C + class (resp. * 4 + coordinate).  The reference transposes,
 * reshapes and concatenates them into [B,A,C] / [B,A,4] (one full read + write of every logit) before the loss and
 * before post-processing.  The ssdk_head_* entry points consume the per-level tensors as they are; anchor index
 * a = level_offset_l + (y * w_l + x) * n + anchor_in_cell, exactly the order of AnchorGenerator (anchor_generator.py:
 * 99-103) and of reshape_and_concatenate.  SSDK_CHANNELS_LAST describes [B, h_l, w_l, n*C] towers. */
#define SSDK_MAX_LEVELS 8
#define SSDK_CHANNELS_LAST 0
#define SSDK_CHANNELS_FIRST 1
typedef struct ssdk_head {
    int32_t num_levels;                 /* 1..SSDK_MAX_LEVELS */
    int32_t anchors_per_location;       /* n */
    int32_t data_format;                /* SSDK_CHANNELS_FIRST or SSDK_CHANNELS_LAST */
    int32_t reserved;
    int32_t height[SSDK_MAX_LEVELS];    /* h_l */
    int32_t width[SSDK_MAX_LEVELS];     /* w_l */
    const float* class_predictions[SSDK_MAX_LEVELS];   /* DEVICE, 16-byte aligned */
    const float* encoded_boxes[SSDK_MAX_LEVELS];       /* DEVICE, 16-byte aligned */
} ssdk_head;
/* Gradient (output) tensors of the same shapes and data format as the head tensors. */
typedef struct ssdk_head_grads {
    float* class_predictions[SSDK_MAX_LEVELS];
    float* encoded_boxes[SSDK_MAX_LEVELS];
} ssdk_head_grads;

/* reshape_and_concatenate (box_predictor.py:67-104) materialised: out_encoded_boxes [B,A,4] and/or
 * out_class_predictions [B,A,C] (either may be NULL).  Not needed by the fused entry points below; provided for callers
 * that want the reference's tensors, and as the un-fused baseline. */
SSDK_API int ssdk_head_concat(ssdk_ctx* ctx, const ssdk_head* head, int B, int C, float* out_encoded_boxes,
                     float* out_class_predictions);
/* ssdk_ssd_loss on the per-level head tensors (no per-anchor outputs).  reg_targets [B,A,4], cls_targets [B,A],
 * matches [B,A] are in anchor order (as ssdk_training_targets writes them); out_sums as ssdk_ssd_loss. */
SSDK_API int ssdk_head_ssd_loss(ssdk_ctx* ctx, const ssdk_head* head, const float* reg_targets, const int32_t* cls_targets,
                       const int32_t* matches, int B, int64_t A, int C, double gamma, double alpha, double* out_sums);
/* ssdk_ssd_targets_and_loss on the per-level head tensors: target assignment (ssd.py:84) + losses (ssd.py:89-133).  The flat
 * pass over the logits needs no targets, so the matching runs inside the same kernel, on CTAs of its own (ALU-bound matching in
 * the issue slots the HBM-bound stream leaves idle); any of out_reg / out_cls / out_matches may be NULL (workspace is used). */
SSDK_API int ssdk_head_ssd_targets_and_loss(ssdk_ctx* ctx, const ssdk_head* head, const float* anchors, const float* gt_boxes,
                                   const int32_t* gt_labels, const int32_t* num_boxes, int B, int64_t A, int C, int Gmax,
                                   double positives_threshold, double negatives_threshold, double gamma, double alpha,
                                   double* out_sums, float* out_reg, int32_t* out_cls, int32_t* out_matches);
/* ssdk_ssd_loss_step on the per-level head tensors. */
SSDK_API int ssdk_head_ssd_loss_step(ssdk_ctx* ctx, const ssdk_head* head, const float* anchors, const float* gt_boxes,
                            const int32_t* gt_labels, const int32_t* num_boxes, int B, int64_t A, int C, int Gmax,
                            double positives_threshold, double negatives_threshold, double gamma, double alpha, int flags,
                            double* out_sums, float* out_losses, float* out_reg, int32_t* out_cls, int32_t* out_matches);
/* ssdk_ssd_loss_forward_backward on the per-level head tensors: one pass over the logits produces the loss sums and
 * the gradients, written per level in the head's own layout (every element of `grads` is written).
 * out_sums may be NULL (backward only).  num_matches, upstream as ssdk_ssd_loss_forward_backward. */
SSDK_API int ssdk_head_ssd_loss_forward_backward(ssdk_ctx* ctx, const ssdk_head* head, const float* reg_targets,
                                        const int32_t* cls_targets, const int32_t* matches, int B, int64_t A, int C,
                                        double gamma, double alpha, const double* num_matches, const float* upstream,
                                        double* out_sums, const ssdk_head_grads* grads);
/* ssdk_detect on the per-level head tensors (flags: SSDK_INPUT_LOGITS or SSDK_INPUT_SCORES; boxes are always encoded).
 * box_scaler may be NULL, final_score_threshold -INFINITY, out_anchor_idx NULL. */
SSDK_API int ssdk_head_detect(ssdk_ctx* ctx, const ssdk_head* head, const float* anchors, int flags, int B, int64_t A, int C,
                     double score_threshold, double iou_threshold, int K, const float* box_scaler,
                     double final_score_threshold, float* out_boxes, float* out_scores, int32_t* out_classes,
                     int32_t* out_num, int32_t* out_anchor_idx);

/* ---- ground-truth side of the random-crop augmentation: detector/input_pipeline/random_image_crop.py --- */
/* The box arithmetic that follows the choice of a crop window (the window itself comes from TensorFlow's
 * sample_distorted_bounding_box; JPEG decoding and the sampler are input-pipeline work and out of scope).  Boxes are
 * DEVICE float[.,4], 16-byte aligned; kept indices are ascending (tf.where + tf.gather order); outputs are caller-owned. */
/* ioa (:190-209): out[i,j] = clip(intersection(boxes1_i, boxes2_j) / (area(boxes2_j) + 1e-8), 0, 1); not symmetric. */
SSDK_API int ssdk_ioa(ssdk_ctx* ctx, const float* boxes1, int64_t n, const float* boxes2, int64_t m, float* out /*[n,m]*/);
/* change_coordinate_frame (:162-187): window DEVICE float[4]; out[i] = clip((box - origin) / size, 0, 1). */
SSDK_API int ssdk_change_coordinate_frame(ssdk_ctx* ctx, const float* boxes, int64_t n, const float* window, float* out);
/* prune_completely_outside_window (:102-131): out_boxes float[n,4] / out_indices int[n] hold the out_num[0] kept boxes
 * (not clipped) first, zero / -1 padding after. */
SSDK_API int ssdk_prune_completely_outside_window(ssdk_ctx* ctx, const float* boxes, int64_t n, const float* window,
                                         float* out_boxes, int32_t* out_indices, int32_t* out_num);
/* prune_non_overlapping_boxes (:134-159): keeps boxes1_i with max_j ioa(boxes2_j, boxes1_i) >= min_overlap. */
SSDK_API int ssdk_prune_non_overlapping_boxes(ssdk_ctx* ctx, const float* boxes1, int64_t n, const float* boxes2, int64_t m,
                                     double min_overlap, float* out_boxes, int32_t* out_indices, int32_t* out_num);
/* randomly_crop_image's box part (:86-99) for a batch in the pipeline's padded format: boxes [B,Gmax,4], num_boxes [B]
 * (NULL = Gmax each), windows [B,4] -> per image: prune outside, prune ioa(window, box) < overlap_thresh, move to the
 * window's frame; out_boxes [B,Gmax,4] zero padded, out_keep_indices int[B,Gmax] (-1 padded; indexes the image's input
 * boxes, use it to gather the labels, :36), out_num int[B]. */
SSDK_API int ssdk_crop_boxes(ssdk_ctx* ctx, const float* boxes, const int32_t* num_boxes, const float* windows, int B, int Gmax,
                    double overlap_thresh, float* out_boxes, int32_t* out_keep_indices, int32_t* out_num);

/* ---- multi-GPU: the one exchange of the path, over NVLink peer memory ------------------------ */
/* Images are sharded over the GPUs of one box, one process per GPU; the only coupling is the loss normaliser and the loss
 * sums (detector/ssd.py:121-123,131-133): an all-reduce(sum) of three doubles.  These entry points do it with peer-memory
 * stores inside one small kernel (see csrc/comm.cu) instead of a library collective; torch.distributed / NCCL remains the
 * plumbing that carries the 64-byte handles (and the fallback when peers cannot map each other's memory).
 *   ssdk_comm_local_handle : allocates this context's mailbox and returns its CUDA IPC handle (HOST buffer, 64 bytes).
 *   ssdk_comm_connect      : `handles` = HOST world*64 bytes, handle of rank r at offset 64*r (all-gathered by the caller).
 *                            Maps every peer's mailbox; SSDK_ERR_NCCL if a peer cannot be mapped (not the same box / no P2P).
 *   ssdk_comm_all_reduce_sum : values (DEVICE double[n], n <= 8) summed over all ranks IN PLACE, in rank order (bit-identical
 *                            on every rank); asynchronous on the context's stream, CUDA-graph capturable.  A collective:
 *                            every rank must issue the same sequence of ssdk_comm_* exchanges.
 *   ssdk_comm_loss_finalize : all-reduce of sums[3] (in place) fused with ssdk_loss_finalize, one launch.
 *   ssdk_comm_error        : synchronises and reports the exchange number at which a peer failed to show up within ~15 s
 *                            (0 = no error; the affected outputs are NaN).
 *   ssdk_comm_world        : number of connected ranks, 0 when not connected. */
#define SSDK_COMM_HANDLE_BYTES 64
SSDK_API int ssdk_comm_local_handle(ssdk_ctx* ctx, void* out_handle);
SSDK_API int ssdk_comm_connect(ssdk_ctx* ctx, int rank, int world, const void* handles);
SSDK_API int ssdk_comm_world(const ssdk_ctx* ctx);
SSDK_API int ssdk_comm_all_reduce_sum(ssdk_ctx* ctx, double* values, int n);
SSDK_API int ssdk_comm_loss_finalize(ssdk_ctx* ctx, double* sums, float* out_losses);
SSDK_API int ssdk_comm_error(ssdk_ctx* ctx, int64_t* out_epoch_of_timeout);
SSDK_API int ssdk_comm_disconnect(ssdk_ctx* ctx);

/* ---- observability: detector/ssd.py:125-129,135-163 (loss summaries) ------------------------ */
/* Per-image, per-level statistics behind the reference's TensorBoard summaries, from the per-anchor vectors:
 *   _add_scalewise_matches_summaries (ssd.py:152-163) / total_mean_matches_per_image (ssd.py:129):
 *       out_matches[b,l] = number of matched anchors (matches >= 0) of image b on level l; the summaries are the
 *       means over b of one column, resp. of the row sums.
 *   _add_scalewise_summaries (ssd.py:135-150), for a per-anchor loss vector `values` [B,A] (all values >= 0):
 *       k_l = ceil(n_l * top_fraction) (ssd.py:146; 0.20 in the reference), tf.nn.top_k(values[:, level l], k_l,
 *       sorted=False) selects the k_l biggest values of every image.  The order inside that selection is unspecified
 *       in TF, so what is exported is what does not depend on it: out_topk_mean[b,l] = mean of the selected values,
 *       out_topk_kth[b,l] = the smallest selected value (the k_l-th biggest; selected set = values >= it, ties cut).
 * values / matches may be NULL (the corresponding outputs are then not written and may be NULL too).
 *   per_level: HOST int32[num_levels] = num_anchors_per_feature_map (anchor_generator.py:65), sum == A.
 * Outputs are DEVICE float[B, num_levels]. */
SSDK_API int ssdk_level_summaries(ssdk_ctx* ctx, const float* values, const int32_t* matches, int B, int64_t A,
                         const int32_t* per_level, int num_levels, double top_fraction, float* out_topk_mean,
                         float* out_topk_kth, float* out_matches);

#ifdef __cplusplus
}
#endif
#endif /* SSDK_H_ */
