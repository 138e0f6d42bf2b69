/*
 * b200ssl.h -- C ABI of the B200-native semi-supervised loss-and-mixing path.
 *
 * One shared library (libb200ssl.so, sm_100a) replaces the ATen call chains behind
 * four Python modules of the reference (cowmix.py, lovasz.py, mean_teacher.py,
 * metrics.py).  Every entry point cites the reference lines it stands in for.
 *
 * Conventions (all entry points)
 *   - return 0 on success, a positive cudaError_t if a launch failed, a negative
 *     B200SSL_E* code for argument errors; b200ssl_last_error() returns a
 *     thread-local message for the last non-zero return.
 *   - all data pointers are DEVICE pointers unless the name ends in `_host`.
 *   - tensors are dense, row-major NCHW fp32 unless stated; labels are int64 by
 *     default (torch.argmax convention), see b200ssl_label_dtype.
 *   - the library allocates no device memory, frees nothing, never synchronises and
 *     never touches the legacy default stream: the caller owns every buffer and
 *     passes a workspace (size from the matching *_workspace_bytes query) and a
 *     stream (cudaStream_t, e.g. torch.cuda.current_stream().cuda_stream).  The only
 *     resources it creates are two internal non-blocking streams and three events
 *     per device for the fork/join inside b200ssl_loss_path_step, and the CUDA events
 *     of the optional profiler.  The one exception to "allocates nothing" is the 2 MB peer mailbox
 *     of b200ssl_peer_create (it has to be exportable through CUDA IPC).
 *   - 16-byte aligned bases take the 128-bit path; other alignments fall back to
 *     scalar accesses inside the same kernel (never an error).
 */
#ifndef B200SSL_H_
#define B200SSL_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
#endif

#define B200SSL_VERSION 100 /* 0.1.0 */

#define B200SSL_EINVAL (-1)    /* bad argument */
#define B200SSL_EWORKSPACE (-2) /* workspace too small */
#define B200SSL_EUNSUPPORTED (-3)
#define B200SSL_ETIMEOUT (-4)    /* a peer exchange gave up waiting (b200ssl_peer_status) */

typedef void* b200ssl_stream_t; /* cudaStream_t */

typedef enum b200ssl_label_dtype {
  B200SSL_I64 = 0,
  B200SSL_I32 = 1,
  B200SSL_U8 = 2
} b200ssl_label_dtype;

int b200ssl_version(void);
const char* b200ssl_last_error(void);
/* number of kernel launches issued by this library in this process (bench.py's gpu_launches) */
long long b200ssl_launch_count(void);

/* Optional per-kernel timing for bench.py's roofline line.  While enabled, every kernel launch of
 * this library is bracketed by two CUDA events recorded on the launch stream.
 * b200ssl_prof_report waits for the recorded events (the only call of the library that
 * synchronises), writes one line per kernel name -- "<name> <launches> <total_ms> <min_ms>\n" --
 * into buf (NUL terminated, truncated to capacity), forgets the records and returns the untruncated
 * length. */
void b200ssl_prof_enable(int on);
long long b200ssl_prof_report(char* buf, size_t capacity);

/* ---------------------------------------------------------------------------------------------
 * Mean-teacher EMA over all parameter tensors in ONE launch.
 * Replaces mean_teacher.update_ema_variables, reference mean_teacher.py:10-11
 *     ema_param.data.mul_(alpha).add_(other=param.data, alpha=1. - alpha)
 * Bit contract (fp32): t = RN(e * (float)alpha);  e' = fmaf(p, (float)(1.0 - alpha), t).
 *
 * The caller describes the tensors once; b200ssl_ema_build_table_host turns that into a chunk
 * table (host memory), which the caller copies to the device and re-uses while the data
 * pointers stay the same.
 * --------------------------------------------------------------------------------------------- */
#define B200SSL_EMA_CHUNK 4096 /* elements per table entry */

typedef struct b200ssl_ema_chunk {
  float* ema;        /* device pointer to first element of the chunk (teacher) */
  const float* param; /* device pointer to first element of the chunk (student) */
  int32_t count;     /* elements in this chunk, 1..B200SSL_EMA_CHUNK */
  int32_t pad_;
} b200ssl_ema_chunk;

/* number of table entries needed for n_tensors tensors of the given element counts */
int64_t b200ssl_ema_table_entries(const int64_t* numels_host, int n_tensors);
/* fill table_host[0..entries); returns entries written or a negative error */
int64_t b200ssl_ema_build_table_host(void* const* ema_ptrs_host, void* const* param_ptrs_host,
                                     const int64_t* numels_host, int n_tensors,
                                     b200ssl_ema_chunk* table_host, int64_t table_capacity);
int b200ssl_ema_multi(const b200ssl_ema_chunk* table_dev, int64_t n_entries, double alpha,
                      b200ssl_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Fused CowMix mixing of up to two tensors that share one mask.
 * Replaces two calls of cowmix.mix_with_mask, reference cowmix.py:72-73 (train.py:82,84-86)
 *     tensor_a * mask + tensor_b * (1. - mask)
 * evaluated as RN(RN(a*m) + RN(b*RN(1-m))) -- bit-identical to the four ATen kernels for any
 * mask value (not only {0,1}), including inf/NaN propagation from the unselected operand.
 *   a0,b0,out0 : [n, c0, hw]   (images)        a1,b1,out1 : [n, c1, hw] (predictions) or NULL,c1=0
 *   mask       : [n, 1, hw]  if mask_channels==1, else [n, c, hw] (per-channel mask; requires c1==0)
 * --------------------------------------------------------------------------------------------- */
int b200ssl_mix2(const float* a0, const float* b0, float* out0, int c0, const float* a1,
                 const float* b1, float* out1, int c1, const float* mask, int mask_channels,
                 int64_t n, int64_t hw, b200ssl_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * CowMix mask generation from caller-supplied noise.
 * Replaces cowmix.dual_pass_gaussian_fileter2d + the statistics/threshold tail of
 * cowmix.generate_cowmix_masks_like, reference cowmix.py:27-37 and :56-68:
 *     V[n,y,x] = sum_i taps[n,i] * noise[n, y+i-k, x]      (zero padded, k = K/2, i ascending)
 *     S[n,y,x] = sum_j taps[n,j] * V[n, y, x+j-k]
 *     tau_n    = thr_factor[n] * std_n(S, unbiased) + mean_n(S)
 *     mask     = (S > tau_n) ? 1.f : 0.f
 * taps are the reference's (off-centre) normalised Gaussian weights, computed by the caller on
 * the host exactly as cowmix.py:6-24 does and uploaded ([n, K] fp32, K odd).
 * thr_factor[n] = erfinv(2p-1)*sqrt(2) (cowmix.py:64), also caller-computed.
 * field_out (optional, may be NULL) receives S for diagnostics/parity margins.
 * --------------------------------------------------------------------------------------------- */
size_t b200ssl_cowmix_workspace_bytes(int n, int h, int w);
int b200ssl_cowmix_mask(const float* noise, const float* taps, int K, const float* thr_factor,
                        int n, int h, int w, float* mask_out, float* field_out, void* workspace,
                        size_t workspace_bytes, b200ssl_stream_t stream);

/* The same pipeline split for fusion: b200ssl_cowmix_field stops after the smoothing and the
 * per-sample threshold (field_out = S [n,h,w], tau_out [n]); b200ssl_mix2_field forms the mask
 * S > tau[image] on the fly while mixing and writes it to mask_out -- bit-identical to
 * b200ssl_cowmix_mask followed by b200ssl_mix2, one pass over S and one launch fewer. */
int b200ssl_cowmix_field(const float* noise, const float* taps, int K, const float* thr_factor, int n, int h,
                         int w, float* field_out, float* tau_out, void* workspace, size_t workspace_bytes,
                         b200ssl_stream_t stream);
int b200ssl_mix2_field(const float* a0, const float* b0, float* out0, int c0, const float* a1,
                       const float* b1, float* out1, int c1, const float* field, const float* tau,
                       float* mask_out, int64_t n, int64_t hw, b200ssl_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Lovasz-softmax forward (+ unit gradients) and backward.
 * Replaces lovasz.lovasz_softmax / lovasz_softmax_flat / flatten_probas / lovasz_grad,
 * reference lovasz.py:155-170, :173-201, :204-220, :19-31, and autograd's backward of them.
 *
 * A *segment* is one (image-group, class) pair that the reference sorts on its own:
 *   per_image != 0 : n_groups = n_images, segment length L = hw
 *   per_image == 0 : n_groups = 1,        segment length L = n_images*hw
 * Classes summed: class_mode ALL/PRESENT -> 0..C-1, LIST -> class_list[0..n_list).
 * Segment order: seg = group * n_cls + class_slot.
 *
 * Forward writes, per segment s:
 *   seg_loss[s]    fp32  dot(errors_sorted, lovasz_grad(fg_sorted))             (lovasz.py:200)
 *   seg_fg[s]      int32 number of valid foreground pixels (0 => absent class)   (lovasz.py:188)
 *   seg_valid[s]   int32 number of non-ignored pixels that were sorted (all of them, except on the multi-class
 *                  probability path, where background pixels behind the last foreground element -- error below
 *                  min_fg |1 - p|, Jaccard delta exactly 0 -- are pruned before the sort)
 * and, for every pixel i of every summed class c, the unit gradient
 *   jgrad[n,c,i] = sign(p - fg) * (J[rank] - J[rank-1])      (0 for ignored pixels)
 * computed with the reference's fp32 operation sequence (integer counts -> IEEE div ->
 * 1-q -> adjacent difference), ties ordered by ascending pixel index (stable).
 * Planes of classes that are not summed are zero-filled.
 * loss_out (optional) receives lovasz_softmax's scalar: mean over groups of the mean over
 * counted classes (PRESENT skips seg_fg==0), sequential fp32 sums as lovasz.py:235-253.
 *
 * Backward: grad_probas[n,c,i] = seg_scale[seg(n,c)] * jgrad[n,c,i] (in place allowed).
 * b200ssl_lovasz_seg_scale derives seg_scale from the scalar upstream gradient for loss_out.
 * --------------------------------------------------------------------------------------------- */
#define B200SSL_LOVASZ_ALL 0
#define B200SSL_LOVASZ_PRESENT 1
#define B200SSL_LOVASZ_LIST 2
#define B200SSL_LOVASZ_MAX_LIST 64
/* error_mode: which per-pixel error is sorted.
 *   ABS   : |fg - class_pred|, `probas` are probabilities                    (lovasz.py:193-200, lovasz_softmax_flat)
 *   HINGE : 1 - logit * (2*label - 1), `probas` are raw logits [B,1,H,W]     (lovasz.py:96-111, lovasz_hinge_flat)
 *           loss = dot(relu(errors_sorted), lovasz_grad(gt_sorted)).  Needs n_channels == 1 and class_mode LIST
 *           with ONE entry, the foreground label (1); every other non-ignored label counts as background.
 *           Pixels with error <= 0 carry relu = 0 and sit behind every positive error in the descending order, so
 *           they are not sorted: gradient 0, and they are counted only in seg_fg / seg_valid.  The gradient is
 *           with respect to the logits. */
#define B200SSL_LOVASZ_ERR_ABS 0
#define B200SSL_LOVASZ_ERR_HINGE 1

typedef struct b200ssl_lovasz_desc {
  int32_t n_images;
  int32_t n_channels;   /* C of probas; 1 = sigmoid mode (class_pred = probas[:,0]) */
  int64_t hw;           /* H*W */
  int32_t per_image;
  int32_t class_mode;   /* B200SSL_LOVASZ_* */
  int32_t n_list;
  int32_t class_list[B200SSL_LOVASZ_MAX_LIST];
  int32_t has_ignore;
  int64_t ignore_index;
  int32_t label_dtype;  /* b200ssl_label_dtype */
  int32_t error_mode;   /* B200SSL_LOVASZ_ERR_* (0 for descriptors that predate the field) */
} b200ssl_lovasz_desc;

int32_t b200ssl_lovasz_num_segments(const b200ssl_lovasz_desc* d);
size_t b200ssl_lovasz_workspace_bytes(const b200ssl_lovasz_desc* d);
int b200ssl_lovasz_forward(const b200ssl_lovasz_desc* d, const float* probas, const void* labels,
                           float* loss_out, float* seg_loss, int32_t* seg_fg, int32_t* seg_valid,
                           float* jgrad, void* workspace, size_t workspace_bytes,
                           b200ssl_stream_t stream);
/* Forward and backward in one go when the upstream gradient of the scalar loss is already known
 * (device scalar grad_out, normally 1): the last radix pass writes the FINAL gradient
 *   grad_probas[n,c,i] = RN(scale_seg * delta) * sign      (the same values b200ssl_lovasz_backward
 * produces from jgrad and b200ssl_lovasz_seg_scale / b200ssl_binary_lovasz_scale), so neither the
 * unit-gradient buffer nor a separate backward launch is needed.
 *   binary_nonzero == NULL : loss_out = lovasz_softmax's scalar, scale as b200ssl_lovasz_seg_scale
 *   binary_nonzero != NULL : losses.binary_lovasz_loss_with_logits (needs per_image, one class):
 *       loss_out = sum_i w_i L_i / denom, denom_out = sum_i w_i + 0.001, w_i = binary_nonzero[i] > 0 */
int b200ssl_lovasz_forward_backward(const b200ssl_lovasz_desc* d, const float* probas, const void* labels,
                                    const float* grad_out, const int32_t* binary_nonzero, float* loss_out,
                                    float* denom_out, float* seg_loss, int32_t* seg_fg, int32_t* seg_valid,
                                    float* grad_probas, void* workspace, size_t workspace_bytes,
                                    b200ssl_stream_t stream);
int b200ssl_lovasz_seg_scale(const b200ssl_lovasz_desc* d, const float* grad_out,
                             const int32_t* seg_fg, const int32_t* seg_valid, float* seg_scale,
                             b200ssl_stream_t stream);
int b200ssl_lovasz_backward(const b200ssl_lovasz_desc* d, const float* seg_scale,
                            const float* jgrad, float* grad_probas, b200ssl_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * losses.binary_lovasz_loss_with_logits glue, reference losses.py:239-250:
 *     int_target = argmax(target, 1);  per sample i:  w_i = (int_target_i.sum() > 0)
 *     loss = sum_i w_i * L_i / (sum_i w_i + 0.001)         (python left-to-right fp32 sums)
 * b200ssl_argmax_channels turns the soft one-hot target [n, C, hw] into integer labels
 * (first maximum wins, as torch.argmax; NaN counts as maximum) and counts the non-zero labels
 * per image (nonzero_out: int32 [n], ACCUMULATED; may be NULL).  out_dtype: B200SSL_I64 or
 * B200SSL_U8 (C <= 256).
 * b200ssl_binary_lovasz_reduce evaluates the weighted mean from the per-image Lovasz losses
 * L_i (seg_loss of a per_image, single-class forward); *_scale produces the per-segment upstream
 * gradients RN(grad_out / denom) * w_i for b200ssl_lovasz_backward.
 * --------------------------------------------------------------------------------------------- */
/* The whole of losses.binary_lovasz_loss_with_logits, forward and backward, with a fused front end:
 * ONE pass over target and scores yields labels_out = argmax_c target (uint8 [n,hw]), nonzero[n],
 * the sort words of class `cls` and -- if cm != NULL -- the confusion matrix cm[label*C + argmax_c
 * scores] (accumulated, int64 [C,C]); then the radix passes write grad = dLoss/dscores for the
 * upstream gradient *grad_out.  Replaces b200ssl_argmax_channels + b200ssl_lovasz_forward_backward +
 * b200ssl_confusion_from_logits.  Returns B200SSL_EUNSUPPORTED (nothing launched) unless
 * 2 <= n_channels <= 16, hw % 4 == 0 and the planes are 16-byte aligned; workspace as for a
 * per_image, single-class b200ssl_lovasz_desc. */
int b200ssl_binary_lovasz_fused(const float* scores, const float* target, int n_images, int n_channels,
                                int64_t hw, int cls, const float* grad_out, unsigned char* labels_out,
                                int32_t* nonzero, float* loss_out, float* denom_out, float* seg_loss,
                                int32_t* seg_fg, int32_t* seg_valid, float* grad, long long* cm,
                                int cm_has_ignore, int64_t cm_ignore_index, void* workspace,
                                size_t workspace_bytes, b200ssl_stream_t stream);
int b200ssl_argmax_channels(const float* x, int n_images, int n_channels, int64_t hw,
                            void* labels_out, int out_dtype, int32_t* nonzero_out,
                            b200ssl_stream_t stream);
int b200ssl_binary_lovasz_reduce(const float* seg_loss, const int32_t* nonzero, int n,
                                 float* loss_out, float* denom_out, b200ssl_stream_t stream);
int b200ssl_binary_lovasz_scale(const float* grad_out, const int32_t* nonzero,
                                const float* denom, int n, float* seg_scale,
                                b200ssl_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Confusion matrix (new component; the reference only has its derived quantities:
 * metrics.dice_metric metrics.py:1-7 and lovasz.iou lovasz.py:54-73).
 *     cm[l*D + p] += 1   for every pixel with label != ignore_index
 * cm is int64 [D*D] and is ACCUMULATED into (zero it first for a fresh matrix).
 *   other_bucket == 0: D = C; pixels whose label (other than ignore_index) or prediction is
 *       outside [0,C) are not counted; their number is added to *dropped (optional int64 scalar).
 *   other_bucket != 0: D = C+1; out-of-range labels / predictions are counted in row / column C
 *       (needed to reproduce lovasz.iou's union exactly when predictions contain void classes).
 * If per_image != 0, cm is [n_pixels/hw, D*D] and pixel i goes to image i / hw.
 * --------------------------------------------------------------------------------------------- */
int b200ssl_confusion_matrix(const void* labels, const void* preds, int64_t n_pixels,
                             int num_classes, int other_bucket, int has_ignore,
                             int64_t ignore_index, int label_dtype, int per_image, int64_t hw,
                             long long* cm, long long* dropped, b200ssl_stream_t stream);

/* Fused argmax + confusion matrix straight from logits ("next" row N3 of SURVEY 8f):
 *   pred = argmax_c logits[n,c,i] (first maximum wins, as torch.argmax), then as above (D = C). */
int b200ssl_confusion_from_logits(const float* logits, const void* labels, int n_images,
                                  int num_classes, int64_t hw, int has_ignore, int64_t ignore_index,
                                  int label_dtype, int per_image, long long* cm, long long* dropped,
                                  b200ssl_stream_t stream);

/* metrics.dice_metric, reference metrics.py:1-7, for input/target [n, chw] fp32:
 *   dice[i] = (2*sum(x*y) + 1) / (sum(x+y) + 1); sums accumulated in fp64 then rounded to fp32
 *   (exact, hence identical to the reference, for {0,1} inputs below 2^24 elements). */
int b200ssl_dice_metric(const float* input, const float* target, int n, int64_t chw,
                        float* dice_out, void* workspace, size_t workspace_bytes,
                        b200ssl_stream_t stream);
size_t b200ssl_dice_workspace_bytes(int n, int64_t chw);

/* The validation metric of train.py:171-175 in one pass: argmax over the channels of logits [n,C,h,w] (the
 * network's resolution), one_hot, `nearest` resize to the mask size (src = min(floor(dst * in/out), in-1), in/out
 * in fp32 like ATen), mask[:, fg] > threshold.  Accumulates the per-image 2x2 matrices cm_per_image [n][4] =
 * {TN, FP, FN, TP} (int64, caller zeroes them; streaming evaluation may accumulate); b200ssl_dice_from_cm then
 * gives metrics.dice_metric (metrics.py:1-7) of (pred_map_binary[:, 1:], mask_binary[:, 1:]) exactly. */
int b200ssl_validation_cm(const float* logits, int n, int n_channels, int h, int w, const float* mask,
                          int mask_channels, int H, int W, float threshold, int fg_class, long long* cm_per_image,
                          b200ssl_stream_t stream);

/* the same quantity from per-image 2x2 confusion matrices [n][4] = {TN, FP, FN, TP}:
 *   dice[n] = (2*TP + 1) / (2*TP + FP + FN + 1)  evaluated in fp32 like the reference. */
int b200ssl_dice_from_cm(const long long* cm_per_image, int n_images, float* dice_out,
                         b200ssl_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Confidence-masked consistency loss of the semi-supervised branch ("next" row N1), reference
 * train.py:98-107 (inline code) and autograd's backward w.r.t. the student logits:
 *     t = sigmoid(teacher), s = sigmoid(student)                      [n, c, hw] fp32 logits
 *     conf = (max_c t > threshold) ? 1 : 0
 *     loss = sum_pixels(sum_c (s-t)^2 * conf) / sum(conf)             (NaN if no pixel is confident)
 * forward : stats_out[0] = loss, [1] = sum(conf), [2] = mean(conf)    (one pass, 8 B/element)
 * backward: grad_student = grad_out * 2 (s-t) conf / sum(conf) * s (1-s)   (12 B/element);
 *           `stats` is the forward's stats_out, grad_out a device scalar.
 * --------------------------------------------------------------------------------------------- */
size_t b200ssl_consistency_workspace_bytes(int n, int64_t hw);
int b200ssl_consistency_forward(const float* student, const float* teacher, int n, int c, int64_t hw,
                                float threshold, float* stats_out, void* workspace, size_t workspace_bytes,
                                b200ssl_stream_t stream);
int b200ssl_consistency_backward(const float* student, const float* teacher, int n, int c, int64_t hw,
                                 float threshold, const float* stats, const float* grad_out,
                                 float* grad_student, b200ssl_stream_t stream);

/* The same loss with the teacher formed on the fly (SURVEY 8f N1 + N2): train.py:69-82 up-samples the two teacher
 * predictions (F.interpolate bilinear, align_corners=False), mixes them with the CowMix mask
 *     mixed_ema_pred = ema_pred_a * mask + ema_pred_b * (1 - mask)          (cowmix.py:72-73)
 * and train.py:98-107 is that tensor's only consumer.  These entry points evaluate it in registers from
 *     teacher_a / teacher_b  [n, c, th, tw]  (th x tw <= h x w; th == h and tw == w: read as they are),
 *     mask [n, 1, h, w] ({0,1} fp32, the output of b200ssl_cowmix_mask / b200ssl_mix2_field),
 * with the arithmetic of b200ssl_mix2_upsampled, so mixed_ema_pred is never written or read
 * (4C + 4 + 8C/s^2 bytes per pixel instead of 12C + 4 + 8C/s^2 for mix + loss) and the results equal the
 * two-step route: stats to the last bits of the fp64 partial sums, gradients bit for bit.
 * conf_out / conf (optional, [n, h, w] bytes): the forward's per-pixel confidence decision; handing it to the
 * backward saves it one evaluation of the teacher (without it the backward recomputes the decision). */
size_t b200ssl_consistency_mixed_workspace_bytes(int n, int h, int w);
int b200ssl_consistency_mixed_forward(const float* student, const float* teacher_a, const float* teacher_b,
                                      const float* mask, int n, int c, int h, int w, int th, int tw, float threshold,
                                      float* stats_out, unsigned char* conf_out, void* workspace,
                                      size_t workspace_bytes, b200ssl_stream_t stream);
int b200ssl_consistency_mixed_backward(const float* student, const float* teacher_a, const float* teacher_b,
                                       const float* mask, int n, int c, int h, int w, int th, int tw, float threshold,
                                       const float* stats, const unsigned char* conf, const float* grad_out,
                                       float* grad_student, b200ssl_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Row N3: lovasz_softmax straight from logits.  lovasz.py:155-160 expects F.softmax(logits, 1); the
 * probabilities are never written here:
 *   b200ssl_softmax_stats          per-pixel max_c x and sum_c exp(x - max)  ([n,hw] fp32 planes each)
 *   b200ssl_lovasz_forward_logits  b200ssl_lovasz_forward with p = exp(x - max) / sum formed inside the
 *                                  key-build; jgrad = unit gradients w.r.t. the PROBABILITIES (or the scaled
 *                                  ones when grad_out != NULL, as b200ssl_lovasz_forward_backward)
 *   b200ssl_softmax_backward       in place on the gradient buffer: dL/dz = (g - sum_c g_c p_c) * p
 * --------------------------------------------------------------------------------------------- */
int b200ssl_softmax_stats(const float* logits, int n, int c, int64_t hw, float* softmax_max, float* softmax_sum,
                          b200ssl_stream_t stream);
int b200ssl_lovasz_forward_logits(const b200ssl_lovasz_desc* d, const float* logits, const float* softmax_max,
                                  const float* softmax_sum, const void* labels, const float* grad_out,
                                  float* loss_out, float* seg_loss, int32_t* seg_fg, int32_t* seg_valid,
                                  float* jgrad, void* workspace, size_t workspace_bytes,
                                  b200ssl_stream_t stream);
int b200ssl_softmax_backward(const float* logits, const float* softmax_max, const float* softmax_sum, float* grad,
                             int n, int c, int64_t hw, b200ssl_stream_t stream);

/* Round 2: the soft-max written out once, probas[n,C,hw] = exp(x - max) / sum (the arithmetic of softmax_stats + the
 * logits key-build), and its backward from the stored probabilities, grad <- (grad - sum_c grad_c p_c) * p in place.
 * With the exact tail pruning on the probability path, soft-max -> b200ssl_lovasz_forward -> this backward is faster
 * than the never-materialising b200ssl_lovasz_forward_logits route (0.58 vs 0.94 ms at 4x21x512x512) at the price of
 * one [n,C,hw] tensor kept for the backward pass. */
int b200ssl_softmax_forward(const float* logits, int n, int c, int64_t hw, float* probas, b200ssl_stream_t stream);
int b200ssl_softmax_backward_probas(const float* probas, float* grad, int n, int c, int64_t hw, b200ssl_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Row N2: bilinear up-sampling (F.interpolate(..., mode='bilinear', align_corners=False), train.py:
 * 71-75,93-94, losses.py:18-19) fused into the mix.  b200ssl_mix2_upsampled is b200ssl_mix2 /
 * b200ssl_mix2_field with the SECOND tensor pair given at low resolution [n,c1,h_in,w_in]: the
 * full-resolution teacher predictions are never materialised.  tau == NULL: `mask` is the {0,1} mask
 * [n,1,h,w]; tau != NULL: `mask` is the smoothed field and the mask is formed and written to mask_out.
 * b200ssl_upsample_bilinear is the stand-alone interpolation of `planes` = n*c planes.
 * Bit-identical to ATen's CPU kernel for out >= in (the only direction the reference uses).
 * --------------------------------------------------------------------------------------------- */
int b200ssl_mix2_upsampled(const float* a0, const float* b0, float* out0, int c0, const float* a1_lo,
                           const float* b1_lo, float* out1, int c1, int h_in, int w_in, const float* mask,
                           const float* tau, float* mask_out, int64_t n, int h, int w,
                           b200ssl_stream_t stream);
int b200ssl_upsample_bilinear(const float* in, int64_t planes, int h_in, int w_in, float* out, int h, int w,
                              b200ssl_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Row N4: multi-tensor gradient clip + SGD step + EMA epilogue.
 *   train.py:122 clip_grad_norm_(params, gradient_clip_value) -> b200ssl_grad_norm_multi (+ _grad_scale_multi
 *                when the scaled gradients themselves are wanted)
 *   train.py:123-124 optimizer.step(); optimizer.zero_grad()  (torch.optim.SGD, configs/default_config.py:151)
 *   train.py:130 / mean_teacher.py:10-11 EMA of the UPDATED parameters
 *                -> b200ssl_sgd_ema_multi: one launch, p/g/momentum/teacher read once.
 * The chunk table is built like the EMA table (chunks of B200SSL_EMA_CHUNK elements); momentum / ema
 * pointer arrays may be NULL (no momentum buffer / no teacher).  norm_and_coef_out: fp32[2] =
 * {total L2 norm of all gradients, min(max_norm / (norm + 1e-6), 1)} -- torch's clip coefficient, which
 * b200ssl_sgd_ema_multi applies to every gradient when coef_dev points at norm_and_coef_out + 1
 * (coef_dev NULL: no clipping).  Arithmetic is torch's, op for op (see csrc/optim.cu).
 * --------------------------------------------------------------------------------------------- */
typedef struct b200ssl_sgd_chunk {
  float* param;
  float* grad;
  float* momentum; /* may be NULL */
  float* ema;      /* may be NULL */
  int32_t count;
  int32_t tensor;
} b200ssl_sgd_chunk;

typedef struct b200ssl_sgd_hyper {
  double lr, momentum, dampening, weight_decay;
  double ema_alpha;   /* < 0: no EMA */
  int32_t nesterov;
  int32_t first_step; /* momentum buffers are uninitialised: buf = grad (torch clones the gradient) */
  int32_t zero_grad;  /* also write zeros to the gradients (optimizer.zero_grad(set_to_none=False)) */
  int32_t reserved_;
} b200ssl_sgd_hyper;

int64_t b200ssl_sgd_build_table_host(void* const* param_ptrs_host, void* const* grad_ptrs_host,
                                     void* const* momentum_ptrs_host, void* const* ema_ptrs_host,
                                     const int64_t* numels_host, int n_tensors,
                                     b200ssl_sgd_chunk* table_host, int64_t table_capacity);
size_t b200ssl_grad_norm_workspace_bytes(int64_t n_entries);
int b200ssl_grad_norm_multi(const b200ssl_sgd_chunk* table_dev, int64_t n_entries, double max_norm,
                            float* norm_and_coef_out, void* workspace, size_t workspace_bytes,
                            b200ssl_stream_t stream);
int b200ssl_grad_scale_multi(const b200ssl_sgd_chunk* table_dev, int64_t n_entries, const float* coef_dev,
                             b200ssl_stream_t stream);
int b200ssl_sgd_ema_multi(const b200ssl_sgd_chunk* table_dev, int64_t n_entries, const float* coef_dev,
                          const b200ssl_sgd_hyper* hyper, b200ssl_stream_t stream);

/* Row N2, student side: losses.CalculateLoss (losses.py:15-22: F.interpolate of every prediction to
 * the target size, bilinear, align_corners=False) + binary_lovasz_loss_with_logits (losses.py:239-250) computed
 * straight from LOW-RESOLUTION logits scores_low [n,C,low_h,low_w]: the fused front end interpolates channel
 * `cls` in registers (ATen's upsample_bilinear2d arithmetic) while it builds the sort words, the last radix pass
 * scatters dLoss/d(up-sampled logit) into the one-channel scratch plane grad_full [n,H,W], and the transposed
 * interpolation gathers it -- deterministically, no floating-point atomics -- into grad_low [n,C,low_h,low_w]
 * (zero for channels other than `cls`).  Workspace: b200ssl_lovasz_workspace_bytes of the descriptor
 * {n, 1 channel, H*W, per_image, class list [cls], ignore 255, uint8 labels}.  Returns B200SSL_EUNSUPPORTED
 * (nothing launched) unless W % 4 == 0, target is 16-byte aligned and the up-sampling ratio is <= 9.
 * b200ssl_upsample_bilinear_backward is the stand-alone transposed interpolation of dense planes. */
int b200ssl_binary_lovasz_lowres(const float* scores_low, const float* target, int n_images, int n_channels,
                                 int low_h, int low_w, int H, int W, int cls, const float* grad_out,
                                 unsigned char* labels_out, int32_t* nonzero, float* loss_out, float* denom_out,
                                 float* seg_loss, int32_t* seg_fg, int32_t* seg_valid, float* grad_full,
                                 float* grad_low, void* workspace, size_t workspace_bytes, b200ssl_stream_t stream);
int b200ssl_upsample_bilinear_backward(const float* grad_full, int64_t planes, int H, int W, float* grad_low,
                                       int low_h, int low_w, b200ssl_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Per-step all-reduce over NVLink peer memory.  Replaces utils/utils.py:43-54
 * reduce_tensor, which the reference calls 4-6 times per step on single scalars (train.py:53,58,109,
 * 113,178,181), by ONE exchange of [n_ints int64 counts || n_floats fp32 scalars]: every rank stores
 * its words straight into every rank's mailbox (one 8-byte store = 32 payload bits + the step's
 * sequence number, so data and ready-flag arrive together), and a one-block kernel on each rank adds
 * the world rows of its own mailbox in rank order (int64 adds / fp64 adds: bit-identical results on all
 * ranks, counts exact).  No NCCL call is on the data path.
 *
 *   b200ssl_peer_create   allocates and zeroes the local mailbox on the current device and returns its
 *                         CUDA IPC handle (handle_out, B200SSL_PEER_HANDLE_BYTES); synchronises once.
 *   b200ssl_peer_connect  maps the mailboxes of all ranks from their handles (world x HANDLE_BYTES, in
 *                         rank order; exchanged by the caller, e.g. torch.distributed.all_gather).
 *   b200ssl_peer_connect_ptrs  the same for communicators living in ONE process (plain device
 *                         pointers from b200ssl_peer_mailbox; peer access must already be enabled).
 *   b200ssl_peer_post     publishes the next step (stream-ordered; waits only if it is more than 3 steps
 *                         ahead of the slowest rank's collect).  floats_host: HOST array of n_floats DEVICE
 *                         pointers to fp32 scalars.  prev_ints_out / prev_floats_out (may be NULL): the same
 *                         launch also completes the PREVIOUS step's exchange into them -- by then every
 *                         peer posted it a whole step ago, so a steady-state step pays no separate collect.
 *   b200ssl_peer_collect  sums the LATEST posted step into ints_out [n_ints] int64 / floats_out [n_floats]
 *                         fp64 on `stream` (ordered after the post); does nothing if that step has already
 *                         been collected (idempotent: the flush after the last step of a run).
 *   b200ssl_peer_allreduce  post + collect of the same step in ONE launch on `stream`.
 * Sequence numbers live in device memory: no call advances host state, so the calls are CUDA-graph
 * capturable and a replayed graph exchanges a new step every time.  One payload shape per communicator.
 *   b200ssl_peer_status   (synchronous) B200SSL_ETIMEOUT if any wait gave up (default 20 s per wait,
 *                         env B200SSL_PEER_TIMEOUT_MS); results of such a step are undefined.
 * Posts must be issued in the same order on all ranks; every rank must collect (or fold-collect) at
 * least every 4th step.
 * Destroy only after all ranks have finished (barrier first).
 * --------------------------------------------------------------------------------------------- */
#define B200SSL_PEER_MAX_RANKS 16
#define B200SSL_PEER_MAX_WORDS 4096 /* 2*n_ints + n_floats */
#define B200SSL_PEER_MAX_FLOATS 8
#define B200SSL_PEER_HANDLE_BYTES 64
typedef struct b200ssl_peer_comm b200ssl_peer_comm;

int b200ssl_peer_create(int rank, int world, b200ssl_peer_comm** comm_out, unsigned char* handle_out);
void* b200ssl_peer_mailbox(b200ssl_peer_comm* comm);
int b200ssl_peer_connect(b200ssl_peer_comm* comm, const unsigned char* handles);
int b200ssl_peer_connect_ptrs(b200ssl_peer_comm* comm, void* const* mailboxes_host);
int b200ssl_peer_post(b200ssl_peer_comm* comm, const long long* ints, int n_ints,
                      const float* const* floats_host, int n_floats, long long* prev_ints_out,
                      double* prev_floats_out, b200ssl_stream_t stream);
int b200ssl_peer_collect(b200ssl_peer_comm* comm, long long* ints_out, double* floats_out,
                         b200ssl_stream_t stream);
int b200ssl_peer_allreduce(b200ssl_peer_comm* comm, const long long* ints, int n_ints,
                           const float* const* floats_host, int n_floats, long long* ints_out,
                           double* floats_out, b200ssl_stream_t stream);
int b200ssl_peer_status(b200ssl_peer_comm* comm);
int b200ssl_peer_destroy(b200ssl_peer_comm* comm);

/* ---------------------------------------------------------------------------------------------
 * The whole loss path of one semi-supervised step in one call (train.py:65-130 order): mask, fused
 * mix of images + teacher predictions, Lovasz forward/backward (unit upstream gradient), EMA,
 * confusion matrix of (labels, argmax scores).  Chains the entry points above; any stage whose input
 * pointer is NULL (noise / image_a / scores / ema_table / cm) is skipped.  The three independent
 * chains (mask+mix, Lovasz+matrix, EMA) are forked onto two internal side streams after the work
 * already queued on `stream` and joined back into it before the call returns, so the caller sees one
 * stream-ordered operation (flag B200SSL_STEP_SERIAL keeps everything on `stream`).
 *   mode BINARY : losses.binary_lovasz_loss_with_logits (losses.py:239-250); target = soft one-hot
 *                 fp32 [n,C,h,w]; labels_u8 [n,h,w] and nonzero [n] are scratch/outputs
 *   mode SOFTMAX: lovasz.lovasz_softmax with `lovasz` as given; target = integer labels [n,h,w]
 * small: fp32[4] = {loss (out), denom (out, binary mode), upstream gradient (in, normally 1), pad}.
 * --------------------------------------------------------------------------------------------- */
#define B200SSL_STEP_BINARY 0
#define B200SSL_STEP_SOFTMAX 1
#define B200SSL_STEP_SERIAL 1    /* flags: keep every kernel on `stream` (no internal fork/join) */
#define B200SSL_STEP_PREFORKED 2 /* flags: the fork point was already recorded by b200ssl_loss_path_fork */
/* A step issued in TWO calls, so that the chains that do not depend on the mask parameters are already
 * running while the host still draws p / sigma and builds the taps: the first call (ISSUE_SIDE) forks and
 * launches the Lovasz + confusion-matrix (+ peer exchange) and EMA chains on the internal side streams and
 * returns without joining; the second call (ISSUE_MAIN, same descriptor with the mask/mix fields filled in)
 * launches mask + mix on `stream` and joins the side streams back into it.  Every ISSUE_SIDE call must be
 * followed by exactly one ISSUE_MAIN call on the same stream and device. */
#define B200SSL_STEP_ISSUE_SIDE 4
#define B200SSL_STEP_ISSUE_MAIN 8

typedef struct b200ssl_step_desc {
  int32_t n, classes, h, w, image_channels, K, mode, cm_has_ignore;
  int64_t cm_ignore_index;
  int32_t cm_label_dtype;
  int32_t flags;  /* B200SSL_STEP_SERIAL | B200SSL_STEP_PREFORKED | B200SSL_STEP_ISSUE_SIDE / _MAIN */
  b200ssl_lovasz_desc lovasz;
  /* inputs */
  const float* noise;      /* [n,1,h,w] */
  const float* taps;       /* [n,K] */
  const float* thr_factor; /* [n] */
  const float* image_a;
  const float* image_b;
  const float* teacher_a;  /* may be NULL */
  const float* teacher_b;
  const float* scores;     /* [n,C,h,w] logits or probabilities */
  const void* target;
  const void* cm_labels;   /* NULL: the Lovasz labels */
  /* outputs */
  float* mask;
  float* mixed_images;
  float* mixed_teacher;
  float* grad;             /* dLoss/dscores */
  long long* cm;           /* [C,C], accumulated */
  /* scratch (caller-owned) */
  float* small;
  float* seg_loss;
  int32_t* seg_fg;
  int32_t* seg_valid;
  int32_t* nonzero;
  unsigned char* labels_u8;
  void* ws_cowmix;
  size_t ws_cowmix_bytes;
  void* ws_lovasz;
  size_t ws_lovasz_bytes;
  /* EMA */
  const b200ssl_ema_chunk* ema_table;
  int64_t ema_entries;
  double ema_alpha;
  /* multi-GPU (optional): when `peer` is set the block that finalises the loss posts [cm || loss] to
   * every rank and sums the PREVIOUS step's exchange into peer_cm_out [C*C] int64 / peer_loss_out [1]
   * fp64 (no extra launch; b200ssl_peer_collect flushes the last step of a run) */
  struct b200ssl_peer_comm* peer;
  long long* peer_cm_out;
  double* peer_loss_out;
  /* row N2: teacher_a / teacher_b are [n,classes,teacher_h,teacher_w] and are bilinearly up-sampled to
   * h x w inside the mix (0 = they are already h x w) */
  int32_t teacher_h, teacher_w;
  /* upstream gradient of the loss (device scalar); NULL = small[2].  A caller that keeps a constant 1.0 on
   * the device passes it here and then `small` needs no initialisation at all. */
  const float* grad_out;
} b200ssl_step_desc;

int b200ssl_loss_path_step(const b200ssl_step_desc* d, b200ssl_stream_t stream);

/* Optional early fork point.  The Lovasz and EMA chains depend neither on the mask parameters nor on
 * the noise field, so a caller that produces the noise on `stream` right before the step (torch.normal)
 * can record the fork point BEFORE doing so: call b200ssl_loss_path_fork(stream), enqueue the noise
 * generation, then call b200ssl_loss_path_step with B200SSL_STEP_PREFORKED.  The side chains then start
 * at the recorded point and overlap the noise generation.  Everything the Lovasz / confusion-matrix /
 * EMA stages read or write (scores, target, grad, labels, cm, small, parameters) must have been
 * allocated and initialised on `stream` before the fork point. */
int b200ssl_loss_path_fork(b200ssl_stream_t stream);

/* sizeof() of the ABI structs as compiled into the library, for binding self-checks:
 * which = 0: b200ssl_ema_chunk, 1: b200ssl_lovasz_desc, 2: b200ssl_step_desc; else 0. */
size_t b200ssl_sizeof(int which);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* B200SSL_H_ */
