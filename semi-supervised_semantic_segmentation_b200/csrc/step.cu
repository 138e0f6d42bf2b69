// One C call for the whole loss path of a semi-supervised step (train.py:65-130 order):
//   mask (cowmix.py:56-68) -> fused mix of images and teacher predictions (cowmix.py:72-73 x2)
//   -> Lovasz forward + backward (lovasz.py / losses.py:239-250) -> EMA (mean_teacher.py:10-11)
//   -> confusion matrix of (labels, argmax scores).
// It chains the per-stage entry points of this library.  Two things make it faster than calling them
// one by one:
//   * host cost: ~15 launches issued back to back from C (~2.5 us each) instead of as many
//     Python/ctypes round trips, so a 16x512x512 step stays GPU-bound;
//   * the three chains of the path are independent of each other -- (mask -> mix), (Lovasz +
//     confusion matrix) and (EMA) -- and bound by different units (fp32 FMA issue, latency/issue of
//     the radix passes, HBM), so they are forked onto two internal side streams and joined back
//     into the caller's stream: the caller still sees ONE stream-ordered operation.
#include <mutex>

#include "common.cuh"
#include "peer_device.cuh"

namespace b200ssl {

struct SideStreams {
  cudaStream_t s[2] = {nullptr, nullptr};
  cudaEvent_t fork = nullptr, join[2] = {nullptr, nullptr};
  bool ok = false;
};

// one set per device, created on first use (non-blocking streams, timing-less events)
static SideStreams* side_streams() {
  static std::mutex mu;
  static SideStreams per_dev[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lk(mu);
  SideStreams& ss = per_dev[dev];
  if (!ss.ok) {
    // s[0] carries the Lovasz chain, the longest one: it gets the highest stream priority so that its
    // blocks are dispatched first and the other chains fill the SMs it leaves free
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    bool good = cudaEventCreateWithFlags(&ss.fork, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < 2 && good; ++i)
      good = cudaStreamCreateWithPriority(&ss.s[i], cudaStreamNonBlocking, i == 0 ? prio_hi : prio_lo) == cudaSuccess &&
             cudaEventCreateWithFlags(&ss.join[i], cudaEventDisableTiming) == cudaSuccess;
    if (!good) return nullptr;
    ss.ok = true;
  }
  return &ss;
}

}  // namespace b200ssl

// the conditions under which b200ssl_binary_lovasz_fused takes the shape (it must not be asked to post
// and then decline): mirrors the check at the top of binary_lovasz_fused_impl
static bool fused_front_end_ok(const b200ssl_step_desc* d, int64_t hw) {
  using namespace b200ssl;
  return d->n > 0 && hw > 0 && d->classes >= 2 && d->classes <= 16 && hw % 4 == 0 && aligned16(d->scores) &&
         aligned16(d->target) && (reinterpret_cast<uintptr_t>(d->labels_u8) & 3u) == 0 && d->labels_u8 && d->nonzero &&
         d->grad && hw <= (1ll << 28) && d->n <= 65535;
}

// which: bit 0 = the mask+mix chain, bit 1 = the side chains (Lovasz + matrix + peer exchange, EMA)
static int loss_path_step_on(const b200ssl_step_desc* d, b200ssl_stream_t s_mix, b200ssl_stream_t s_lovasz,
                             b200ssl_stream_t s_ema, int which) {
  using namespace b200ssl;
  const int64_t hw = (int64_t)d->h * d->w;
  int rc;
  bool posted = false;
  b200ssl_stream_t stream = s_mix;
  const bool do_main = (which & 1) != 0, do_side = (which & 2) != 0;

  // 1.+2. mask and mix.  With images present the threshold pass is fused into the mix: the field S
  // and tau live in the cowmix workspace (S in its second region, tau behind the partials).
  if (!do_main) {
  } else if (d->noise && d->image_a) {
    const size_t plane_bytes = align_up((size_t)d->n * hw * sizeof(float), 256);
    float* field = reinterpret_cast<float*>(static_cast<char*>(d->ws_cowmix) + plane_bytes);
    const size_t need = b200ssl_cowmix_workspace_bytes(d->n, d->h, d->w);
    B200SSL_REQUIRE(d->ws_cowmix && d->ws_cowmix_bytes >= need, "loss_path_step: cowmix workspace too small");
    float* tau = reinterpret_cast<float*>(static_cast<char*>(d->ws_cowmix) + need - align_up((size_t)d->n * sizeof(float), 256));
    rc = b200ssl_cowmix_field(d->noise, d->taps, d->K, d->thr_factor, d->n, d->h, d->w, field, tau, d->ws_cowmix,
                              d->ws_cowmix_bytes, stream);
    if (rc) return rc;
    const bool lowres = d->teacher_a && d->teacher_h > 0 && d->teacher_w > 0 && (d->teacher_h != d->h || d->teacher_w != d->w);
    if (lowres)
      rc = b200ssl_mix2_upsampled(d->image_a, d->image_b, d->mixed_images, d->image_channels, d->teacher_a, d->teacher_b,
                                  d->mixed_teacher, d->classes, d->teacher_h, d->teacher_w, field, tau, d->mask, d->n,
                                  d->h, d->w, stream);
    else
      rc = b200ssl_mix2_field(d->image_a, d->image_b, d->mixed_images, d->image_channels, d->teacher_a, d->teacher_b,
                              d->mixed_teacher, d->teacher_a ? d->classes : 0, field, tau, d->mask, d->n, hw, stream);
    if (rc) return rc;
  } else {
    if (d->noise) {
      rc = b200ssl_cowmix_mask(d->noise, d->taps, d->K, d->thr_factor, d->n, d->h, d->w, d->mask, nullptr,
                               d->ws_cowmix, d->ws_cowmix_bytes, stream);
      if (rc) return rc;
    }
    if (d->image_a) {
      if (d->teacher_a && d->teacher_h > 0 && d->teacher_w > 0 && (d->teacher_h != d->h || d->teacher_w != d->w))
        rc = b200ssl_mix2_upsampled(d->image_a, d->image_b, d->mixed_images, d->image_channels, d->teacher_a,
                                    d->teacher_b, d->mixed_teacher, d->classes, d->teacher_h, d->teacher_w, d->mask,
                                    nullptr, nullptr, d->n, d->h, d->w, stream);
      else
        rc = b200ssl_mix2(d->image_a, d->image_b, d->mixed_images, d->image_channels, d->teacher_a, d->teacher_b,
                          d->mixed_teacher, d->teacher_a ? d->classes : 0, d->mask, 1, d->n, hw, stream);
      if (rc) return rc;
    }
  }
  if (!do_side) return 0;
  // 3. Lovasz forward + backward with the upstream gradient small[2], 5. confusion matrix
  stream = s_lovasz;
  if (d->scores) {
    float* loss = d->small;        // [0] loss  [1] denom  [2] upstream gradient (1.0) unless d->grad_out is given
    const float* grad_out = d->grad_out ? d->grad_out : d->small + 2;
    bool done = false;
    if (d->mode == B200SSL_STEP_BINARY) {
      // losses.py:240: int_target = argmax(target, 1); :246 w_i = (tgt.sum() > 0)
      B200SSL_REQUIRE(d->labels_u8 && d->nonzero, "loss_path_step: binary mode needs labels_u8 / nonzero scratch");
      // fused front end: labels, weights, sort words and (when it uses the same labels) the matrix in one pass
      // multi-GPU: when the matrix comes out of the same front end, the block that finalises the loss
      // also posts [cm || loss] to the peers (one kernel computes and communicates)
      PeerTail tail = {};
      const bool fuse_post = d->peer && (!d->cm || !d->cm_labels) && fused_front_end_ok(d, hw);
      if (fuse_post) {
        rc = peer_tail(d->peer, d->cm ? d->classes * d->classes : 0, 1, &tail);
        if (rc) return rc;
        tail.ints = d->cm;
        tail.n_ints = d->cm ? d->classes * d->classes : 0;
        tail.prev_ints_out = d->cm ? d->peer_cm_out : nullptr;
        tail.prev_floats_out = d->peer_loss_out;
      }
      rc = binary_lovasz_fused_impl(d->scores, static_cast<const float*>(d->target), d->n, d->classes, hw, 1,
                                    grad_out, d->labels_u8, d->nonzero, loss, d->small + 1, d->seg_loss,
                                    d->seg_fg, d->seg_valid, d->grad, d->cm_labels ? nullptr : d->cm,
                                    d->cm_has_ignore, d->cm_ignore_index, d->ws_lovasz, d->ws_lovasz_bytes,
                                    stream, fuse_post ? &tail : nullptr);
      if (fuse_post && rc == 0) posted = true;
      if (rc == 0) {
        done = true;
        if (d->cm && d->cm_labels) {
          rc = b200ssl_confusion_from_logits(d->scores, d->cm_labels, d->n, d->classes, hw, d->cm_has_ignore,
                                             d->cm_ignore_index, d->cm_label_dtype, 0, d->cm, nullptr, stream);
          if (rc) return rc;
        }
      } else if (rc != B200SSL_EUNSUPPORTED) {
        return rc;
      }
    }
    if (!done) {
      b200ssl_lovasz_desc ld = d->lovasz;
      const void* labels = d->target;
      if (d->mode == B200SSL_STEP_BINARY) {
        cudaMemsetAsync(d->nonzero, 0, (size_t)d->n * sizeof(int32_t), (cudaStream_t)stream);
        rc = b200ssl_argmax_channels(static_cast<const float*>(d->target), d->n, d->classes, hw, d->labels_u8,
                                     B200SSL_U8, d->nonzero, stream);
        if (rc) return rc;
        labels = d->labels_u8;
        ld.n_images = d->n; ld.n_channels = d->classes; ld.hw = hw; ld.per_image = 1;
        ld.class_mode = B200SSL_LOVASZ_LIST; ld.n_list = 1; ld.class_list[0] = 1;
        ld.has_ignore = 1; ld.ignore_index = 255; ld.label_dtype = B200SSL_U8;
      }
      // The matrix does not depend on the Lovasz chain: it goes FIRST, so that the chain's last kernel -- the
      // one-block finalisation -- can post [cm || loss] to the peers and complete the previous step's exchange
      // (no separate post kernel behind the critical chain: it cost 26 us per step at 21 classes).
      if (d->cm) {
        rc = b200ssl_confusion_from_logits(d->scores, d->cm_labels ? d->cm_labels : labels, d->n, d->classes, hw,
                                           d->cm_has_ignore, d->cm_ignore_index,
                                           d->cm_labels ? d->cm_label_dtype : ld.label_dtype, 0, d->cm, nullptr,
                                           stream);
        if (rc) return rc;
      }
      PeerTail tail = {};
      const bool fuse_post = d->peer && d->n > 0 && hw > 0;
      if (fuse_post) {
        rc = peer_tail(d->peer, d->cm ? d->classes * d->classes : 0, 1, &tail);
        if (rc) return rc;
        tail.ints = d->cm;
        tail.n_ints = d->cm ? d->classes * d->classes : 0;
        tail.prev_ints_out = d->cm ? d->peer_cm_out : nullptr;
        tail.prev_floats_out = d->peer_loss_out;
      }
      rc = lovasz_forward_backward_tail(&ld, d->scores, labels, grad_out,
                                        d->mode == B200SSL_STEP_BINARY ? d->nonzero : nullptr, loss, d->small + 1,
                                        d->seg_loss, d->seg_fg, d->seg_valid, d->grad, d->ws_lovasz, d->ws_lovasz_bytes,
                                        stream, fuse_post ? &tail : nullptr);
      if (rc) return rc;
      if (fuse_post) posted = true;
    }
  }
  // 6. multi-GPU: post [cm || loss] into every rank's mailbox and collect the PREVIOUS step into
  // peer_cm_out / peer_loss_out (unless the Lovasz finalising block already did both)
  if (d->peer) {
    B200SSL_REQUIRE(d->scores && d->small, "loss_path_step: the peer exchange needs the Lovasz stage");
    if (!posted) {
      const float* scalars[1] = {d->small};
      rc = peer_post_impl(d->peer, d->cm, d->cm ? d->classes * d->classes : 0, scalars, 1,
                          d->cm ? d->peer_cm_out : nullptr, d->peer_loss_out, (cudaStream_t)s_lovasz);
      if (rc) return rc;
    }
  }
  // 4. EMA over all parameters
  stream = s_ema;
  if (d->ema_table && d->ema_entries > 0) {
    rc = b200ssl_ema_multi(d->ema_table, d->ema_entries, d->ema_alpha, stream);
    if (rc) return rc;
  }
  return 0;
}

extern "C" int b200ssl_loss_path_step(const b200ssl_step_desc* d, b200ssl_stream_t stream) {
  using namespace b200ssl;
  B200SSL_REQUIRE(d != nullptr, "loss_path_step: null descriptor");
  B200SSL_REQUIRE(d->n >= 1 && d->classes >= 1 && d->h >= 1 && d->w >= 1, "loss_path_step: bad extents");
  B200SSL_REQUIRE(d->mode == B200SSL_STEP_BINARY || d->mode == B200SSL_STEP_SOFTMAX, "loss_path_step: bad mode");
  cudaStream_t main = (cudaStream_t)stream;
  const bool only_side = (d->flags & B200SSL_STEP_ISSUE_SIDE) != 0, only_main = (d->flags & B200SSL_STEP_ISSUE_MAIN) != 0;
  B200SSL_REQUIRE(!(only_side && only_main), "loss_path_step: ISSUE_SIDE and ISSUE_MAIN are two separate calls");
  // a split step always has a mask+mix chain coming (that is what the second call issues)
  const bool has_mix = d->noise || d->image_a || only_side, has_lov = d->scores != nullptr;
  const bool has_ema = d->ema_table && d->ema_entries > 0;
  const bool serial = (d->flags & B200SSL_STEP_SERIAL) != 0, preforked = (d->flags & B200SSL_STEP_PREFORKED) != 0;
  B200SSL_REQUIRE(!(serial && (only_side || only_main)), "loss_path_step: a split step cannot be serial");
  SideStreams* ss = (((int)has_mix + (int)has_lov + (int)has_ema >= 2 || only_main) && !serial) ? side_streams() : nullptr;
  if (!ss) {
    B200SSL_REQUIRE(!(only_side || only_main), "loss_path_step: side streams unavailable for a split step");
    return loss_path_step_on(d, stream, stream, stream, 3);
  }
  int rc = 0;
  if (!only_main) {
    // fork: the side streams start after everything already queued on the caller's stream (or at the
    // point recorded earlier by b200ssl_loss_path_fork)
    if ((!preforked && cudaEventRecord(ss->fork, main) != cudaSuccess) ||
        cudaStreamWaitEvent(ss->s[0], ss->fork, 0) != cudaSuccess ||
        cudaStreamWaitEvent(ss->s[1], ss->fork, 0) != cudaSuccess) {
      set_error("loss_path_step: fork failed: %s", cudaGetErrorString(cudaGetLastError()));
      return (int)cudaErrorUnknown;
    }
  }
  // Lovasz chain on the high-priority side stream, mask+mix on the caller's stream, EMA on the other
  if (only_side) {
    rc = loss_path_step_on(d, stream, ss->s[0], ss->s[1], 2);
    for (int i = 0; i < 2; ++i) cudaEventRecord(ss->join[i], ss->s[i]);   // joined by the ISSUE_MAIN call
    return rc;
  }
  rc = loss_path_step_on(d, stream, ss->s[0], ss->s[1], only_main ? 1 : 3);
  // join (also on error paths, so that the caller's stream never runs ahead of the side work)
  for (int i = 0; i < 2; ++i) {
    if (!only_main) cudaEventRecord(ss->join[i], ss->s[i]);
    cudaStreamWaitEvent(main, ss->join[i], 0);
  }
  return rc;
}

extern "C" int b200ssl_loss_path_fork(b200ssl_stream_t stream) {
  using namespace b200ssl;
  SideStreams* ss = side_streams();
  if (!ss || cudaEventRecord(ss->fork, (cudaStream_t)stream) != cudaSuccess) {
    set_error("loss_path_fork: %s", cudaGetErrorString(cudaGetLastError()));
    return (int)cudaErrorUnknown;
  }
  return 0;
}
