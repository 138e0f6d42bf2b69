"""The whole loss path of one semi-supervised step as a single call (what bench.py times).

Order (train.py:65-130):
    mask   = cowmix.generate_cowmix_masks_like(image_a, p_range, sigma_range)      train.py:77-80
    mixed  = mix_with_mask(teacher_a, teacher_b, mask), mix_with_mask(image_a, image_b, mask)  :82-86
    loss, dloss/dlogits = Lovasz (binary shim losses.py:239-250, or lovasz_softmax)          :51,:61
    update_ema_variables(model, ema_model, alpha)                                             :130
    confusion matrix of (labels, argmax logits)                                               (new)
Nothing in here synchronises with the host; results are device tensors.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import lib, check, stream_ptr
from . import cowmix, lovasz, losses, mean_teacher, metrics


class LossPathStep:
    def __init__(self, num_classes, mask_proportion_range=(0.45, 0.55), sigma_range=(8, 32),
                 ema_alpha=0.99, mode="binary", classes="present", per_image=False, ignore=255):
        if mode not in ("binary", "softmax"):
            raise ValueError("mode must be 'binary' (losses.binary_lovasz_loss_with_logits) or 'softmax'")
        self.num_classes = num_classes
        self.mask_proportion_range = mask_proportion_range
        self.sigma_range = sigma_range
        self.ema_alpha = ema_alpha
        self.mode = mode
        self.classes = classes
        self.per_image = per_image
        self.ignore = ignore
        self._ema = mean_teacher.EmaUpdater()

    # -- Lovasz loss and its gradient w.r.t. the logits/probabilities, without autograd bookkeeping
    def lovasz_loss_and_grad(self, scores, target):
        dev = scores.device
        scores = scores.contiguous()
        if self.mode == "binary":
            labels, nonzero = losses.argmax_channels(target)
            desc = lovasz._make_desc(scores, labels, [1], True, 255)
        else:
            labels = target.contiguous()
            desc = lovasz._make_desc(scores, labels, self.classes, self.per_image, self.ignore)
        n_seg = lib.b200ssl_lovasz_num_segments(C.byref(desc))
        if n_seg < 0:
            check(n_seg, "lovasz_num_segments")
        small = torch.empty(4 + 2 * n_seg, dtype=torch.float32, device=dev)   # loss, denom, one, pad, seg_loss, seg_scale
        seg_meta = torch.empty((2, max(n_seg, 1)), dtype=torch.int32, device=dev)
        jgrad = torch.empty_like(scores)
        grad = torch.empty_like(scores)
        ws = _lib.workspaces.get(dev, "lovasz", lib.b200ssl_lovasz_workspace_bytes(C.byref(desc)))
        s = stream_ptr(dev)
        base = small.data_ptr()
        seg_loss_p, seg_scale_p = base + 16, base + 16 + 4 * n_seg
        with torch.cuda.device(dev):
            small[2] = 1.0  # upstream gradient of the scalar loss
            check(lib.b200ssl_lovasz_forward(
                C.byref(desc), scores.data_ptr(), labels.data_ptr(), base, seg_loss_p,
                seg_meta[0].data_ptr(), seg_meta[1].data_ptr(), jgrad.data_ptr(), ws.data_ptr(),
                ws.numel(), s), "lovasz_forward")
            if self.mode == "binary":
                check(lib.b200ssl_binary_lovasz_reduce(seg_loss_p, nonzero.data_ptr(), n_seg, base, base + 4, s),
                      "binary_lovasz_reduce")
                check(lib.b200ssl_binary_lovasz_scale(base + 8, nonzero.data_ptr(), base + 4, n_seg, seg_scale_p, s),
                      "binary_lovasz_scale")
            else:
                check(lib.b200ssl_lovasz_seg_scale(C.byref(desc), base + 8, seg_meta[0].data_ptr(),
                                                   seg_meta[1].data_ptr(), seg_scale_p, s), "lovasz_seg_scale")
            check(lib.b200ssl_lovasz_backward(C.byref(desc), seg_scale_p, jgrad.data_ptr(), grad.data_ptr(), s),
                  "lovasz_backward")
        return small[0], grad, labels

    def __call__(self, image_a, image_b, teacher_a, teacher_b, scores, target, params, ema_params,
                 cm_labels=None, cm_out=None):
        """scores: student logits (binary mode) or probabilities (softmax mode) [N,C,H,W];
        target: soft one-hot [N,C,H,W] (binary mode) or integer labels [N,H,W] (softmax mode);
        params / ema_params: lists of student / teacher parameter tensors;
        cm_labels: integer labels for the confusion matrix (defaults to the Lovasz labels)."""
        with torch.no_grad():
            mask = cowmix.generate_cowmix_masks_like(image_a, self.mask_proportion_range, self.sigma_range)
            mixed_images, mixed_teacher = cowmix.mix2_with_mask(image_a, image_b, teacher_a, teacher_b, mask)
            loss, grad, labels = self.lovasz_loss_and_grad(scores, target)
            self._ema(ema_params, params, self.ema_alpha)
            cm = metrics.confusion_matrix_from_logits(
                scores, labels if cm_labels is None else cm_labels,
                ignore_index=self.ignore, out=cm_out)
        return {"mask": mask, "mixed_images": mixed_images, "mixed_teacher": mixed_teacher,
                "loss": loss, "grad": grad, "cm": cm}
