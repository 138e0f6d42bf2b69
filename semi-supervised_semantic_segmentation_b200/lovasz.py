"""Drop-in for the on-path part of the reference's lovasz.py.

    reference lovasz.py:155-170  lovasz_softmax(probas, labels, classes, per_image, ignore)
    reference lovasz.py:173-201  lovasz_softmax_flat   (folded into the CUDA forward)
    reference lovasz.py:204-220  flatten_probas        (no copy here: NCHW is read in place)
    reference lovasz.py:19-31    lovasz_grad           (fused into the last radix pass)
    reference lovasz.py:79-111   lovasz_hinge / lovasz_hinge_flat / flatten_binary_scores (same sort core, hinge errors)
    reference lovasz.py:54-73    iou / iou_binary      (derived from the confusion matrix)

Forward computes the loss AND the unit gradient (sort rank -> Jaccard delta, scattered back to
the pixel); backward is one elementwise scale.  Tie order is stable by pixel index.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import lib, check, stream_ptr, require_cuda, LovaszDesc


def _make_desc(probas, labels, classes, per_image, ignore):
    b, c = probas.shape[0], probas.shape[1]
    hw = probas[0, 0].numel() if probas.numel() else (probas.shape[2] * probas.shape[3])
    d = LovaszDesc()
    d.n_images, d.n_channels, d.hw = b, c, hw
    d.per_image = 1 if per_image else 0
    if isinstance(classes, str):
        if classes not in ("all", "present"):
            # the reference falls through to iterating the string and fails on `labels == 'x'`
            raise ValueError(f"classes must be 'all', 'present' or a list of class indices, got {classes!r}")
        if c == 1:
            # lovasz.py:190-192: len('present') > 1 -> ValueError as soon as a class is summed
            raise ValueError('Sigmoid output possible only with 1 class')
        d.class_mode = _lib.LOVASZ_ALL if classes == "all" else _lib.LOVASZ_PRESENT
        d.n_list = 0
    else:
        cls = [int(x) for x in classes]
        if c == 1 and len(cls) > 1:
            raise ValueError('Sigmoid output possible only with 1 class')
        if len(cls) == 0:
            d.class_mode, d.n_list = _lib.LOVASZ_LIST, 0
        else:
            if len(cls) > _lib.MAX_LIST:
                raise ValueError(f"at most {_lib.MAX_LIST} classes can be listed")
            if len(set(cls)) != len(cls):
                raise ValueError("b200ssl.lovasz: duplicate entries in `classes` are not supported")
            for x in cls:
                if c != 1 and not (0 <= x < c):
                    raise IndexError(f"index {x} is out of bounds for dimension 1 with size {c}")
            d.class_mode, d.n_list = _lib.LOVASZ_LIST, len(cls)
            for i, x in enumerate(cls):
                d.class_list[i] = x
    d.has_ignore = 0 if ignore is None else 1
    d.ignore_index = 0 if ignore is None else int(ignore)
    d.label_dtype = _lib.label_dtype_code(labels)
    return d


class _LovaszForward(torch.autograd.Function):
    """(probas, labels) -> (scalar loss, per-segment losses, per-segment fg counts)."""

    @staticmethod
    def forward(ctx, probas, labels, desc, want_scalar):
        ctx.set_materialize_grads(False)
        dev = probas.device
        n_seg = lib.b200ssl_lovasz_num_segments(C.byref(desc))
        if n_seg < 0:
            check(n_seg, "lovasz_num_segments")
        loss = torch.empty((), dtype=torch.float32, device=dev)
        seg_loss = torch.empty(max(n_seg, 1), dtype=torch.float32, device=dev)[:n_seg]
        seg_meta = torch.empty((2, max(n_seg, 1)), dtype=torch.int32, device=dev)
        jgrad = torch.empty_like(probas)
        ws_bytes = lib.b200ssl_lovasz_workspace_bytes(C.byref(desc))
        ws = _lib.workspaces.get(dev, "lovasz", ws_bytes)
        with torch.cuda.device(dev):
            check(lib.b200ssl_lovasz_forward(
                C.byref(desc), probas.data_ptr(), labels.data_ptr(), loss.data_ptr(),
                seg_loss.data_ptr(), seg_meta[0].data_ptr(), seg_meta[1].data_ptr(),
                jgrad.data_ptr(), ws.data_ptr(), ws.numel(), stream_ptr(dev)), "lovasz_forward")
        ctx.desc = desc
        ctx.n_seg = n_seg
        ctx.save_for_backward(jgrad, seg_meta)
        seg_fg = seg_meta[0, :n_seg]
        ctx.mark_non_differentiable(seg_fg)
        return loss, seg_loss, seg_fg

    @staticmethod
    def backward(ctx, g_loss, g_seg, _g_fg):
        jgrad, seg_meta = ctx.saved_tensors
        dev = jgrad.device
        desc, n_seg = ctx.desc, ctx.n_seg
        scale = None
        with torch.cuda.device(dev):
            if g_loss is not None:
                g_loss = g_loss.to(torch.float32).contiguous()
                scale = torch.empty(max(n_seg, 1), dtype=torch.float32, device=dev)
                check(lib.b200ssl_lovasz_seg_scale(
                    C.byref(desc), g_loss.data_ptr(), seg_meta[0].data_ptr(), seg_meta[1].data_ptr(),
                    scale.data_ptr(), stream_ptr(dev)), "lovasz_seg_scale")
            if g_seg is not None:
                # explicit per-segment upstream gradients (losses.binary_lovasz_loss_with_logits)
                g_seg = g_seg.to(torch.float32).reshape(-1)
                scale = g_seg.contiguous() if scale is None else (scale[:n_seg] + g_seg).contiguous()
            if scale is None:
                return None, None, None, None
            if scale.numel() == 0:
                scale = torch.zeros(1, dtype=torch.float32, device=dev)
            grad = torch.empty_like(jgrad)
            check(lib.b200ssl_lovasz_backward(
                C.byref(desc), scale.data_ptr(), jgrad.data_ptr(), grad.data_ptr(), stream_ptr(dev)),
                "lovasz_backward")
        return grad, None, None, None


def _prepare(probas, labels):
    require_cuda(probas, "probas", torch.float32)
    require_cuda(labels, "labels")
    if probas.dim() == 3:  # lovasz.py:208-211: output of a sigmoid layer
        probas = probas.unsqueeze(1)
    if probas.dim() != 4:
        raise ValueError("probas must be [B,C,H,W] or [B,H,W]")
    if labels.dim() != 3 or labels.shape[0] != probas.shape[0] or labels[0].numel() != probas[0, 0].numel():
        raise ValueError(f"labels {tuple(labels.shape)} do not match probas {tuple(probas.shape)}")
    return probas.contiguous(), labels.contiguous()


def lovasz_segment_losses(probas, labels, classes='present', per_image=False, ignore=None):
    """Per-(image-group, class) Lovasz losses `dot(errors_sorted, lovasz_grad(fg_sorted))`
    (lovasz.py:200) as a differentiable [groups, classes] tensor plus the fg pixel counts."""
    probas, labels = _prepare(probas, labels)
    desc = _make_desc(probas, labels, classes, per_image, ignore)
    _, seg_loss, seg_fg = _LovaszForward.apply(probas, labels, desc, False)
    groups = probas.shape[0] if per_image else 1
    return seg_loss.view(groups, -1), seg_fg.view(groups, -1)


def lovasz_softmax(probas, labels, classes='present', per_image=False, ignore=None):
    """
    Multi-class Lovasz-Softmax loss
      probas: [B, C, H, W] class probabilities at each prediction (between 0 and 1).
              Interpreted as binary (sigmoid) output with outputs of size [B, H, W].
      labels: [B, H, W] Tensor, ground truth labels (between 0 and C - 1)
      classes: 'all' for all, 'present' for classes present in labels, or a list of classes to average.
      per_image: compute the loss per image instead of per batch
      ignore: void class labels
    Deviations from the reference, all on degenerate inputs: an image whose pixels are all void
    contributes a zero loss instead of an empty tensor (lovasz.py:180-182 returns `probas * 0.`
    there, which poisons the per-image mean); C == 1 with a string `classes` always raises the
    reference's ValueError (the reference raises it only when class 0 is present).
    """
    probas, labels = _prepare(probas, labels)
    if not isinstance(classes, str) and len(classes) == 0:
        return 0  # mean([]) == 0 (lovasz.py:246)
    if probas.shape[0] == 0 or probas[0, 0].numel() == 0:
        if per_image and probas.shape[0] == 0:
            return 0
        return probas.permute(0, 2, 3, 1).reshape(-1, probas.shape[1]) * 0.  # lovasz.py:180-182
    desc = _make_desc(probas, labels, classes, per_image, ignore)
    loss, _, _ = _LovaszForward.apply(probas, labels, desc, True)
    return loss


# --------------------------- binary Lovasz hinge (lovasz.py:79-111) ------------------------------
def lovasz_hinge(logits, labels, per_image=True, ignore=None):
    """
    Binary Lovasz hinge loss
      logits: [B, H, W] logits at each pixel (between -infinity and +infinity)
      labels: [B, H, W] Tensor, binary ground truth masks (0 or 1)
      per_image: compute the loss per image instead of per batch
      ignore: void class id
    Same kernels as lovasz_softmax with the error 1 - logit * (2*label - 1) (one rounding, as the reference);
    pixels whose error is <= 0 have relu(error) = 0 and sort behind every positive error, so they are never
    sorted and get a zero gradient.  Labels other than 1 (and `ignore`) count as background.
    """
    require_cuda(logits, "logits", torch.float32)
    if torch.is_tensor(labels) and (labels.is_floating_point() or labels.dtype == torch.bool):
        labels = labels.to(torch.int64)   # the reference takes float / bool masks too (`labels.float()`, lovasz.py:105)
    require_cuda(labels, "labels")
    if logits.dim() != 3 or tuple(labels.shape) != tuple(logits.shape):
        raise ValueError(f"logits {tuple(logits.shape)} and labels {tuple(labels.shape)} must both be [B,H,W]")
    if logits.numel() == 0:
        if per_image and logits.shape[0] == 0:
            return 0                      # mean of no images (lovasz.py:246)
        return logits.sum() * 0.          # lovasz.py:103-105: only void pixels
    x, labels = logits.unsqueeze(1).contiguous(), labels.contiguous()
    desc = _make_desc(x, labels, [1], per_image, ignore)
    desc.error_mode = _lib.LOVASZ_ERR_HINGE
    loss, _, _ = _LovaszForward.apply(x, labels, desc, True)
    return loss


# --------------------------- row N3: lovasz_softmax straight from logits ------------------------
class _LovaszFromLogits(torch.autograd.Function):
    """logits -> lovasz_softmax(F.softmax(logits, 1), labels) without materialising the probabilities:
    per-pixel (max, sum) statistics, probabilities formed inside the key-build, soft-max backward
    applied in place on the gradient (b200ssl_softmax_stats / _lovasz_forward_logits / _softmax_backward)."""

    @staticmethod
    def forward(ctx, logits, labels, desc):
        dev = logits.device
        n, c = logits.shape[0], logits.shape[1]
        hw = logits[0, 0].numel()
        n_seg = lib.b200ssl_lovasz_num_segments(C.byref(desc))
        if n_seg < 0:
            check(n_seg, "lovasz_num_segments")
        loss = torch.empty((), dtype=torch.float32, device=dev)
        seg_loss = torch.empty(max(n_seg, 1), dtype=torch.float32, device=dev)
        seg_meta = torch.empty((2, max(n_seg, 1)), dtype=torch.int32, device=dev)
        stats = torch.empty((2, n, hw), dtype=torch.float32, device=dev)
        jgrad = torch.empty_like(logits)
        ws = _lib.workspaces.get(dev, "lovasz", lib.b200ssl_lovasz_workspace_bytes(C.byref(desc)))
        with torch.cuda.device(dev):
            st = stream_ptr(dev)
            check(lib.b200ssl_softmax_stats(logits.data_ptr(), n, c, hw, stats[0].data_ptr(), stats[1].data_ptr(), st),
                  "softmax_stats")
            check(lib.b200ssl_lovasz_forward_logits(
                C.byref(desc), logits.data_ptr(), stats[0].data_ptr(), stats[1].data_ptr(), labels.data_ptr(), None,
                loss.data_ptr(), seg_loss.data_ptr(), seg_meta[0].data_ptr(), seg_meta[1].data_ptr(),
                jgrad.data_ptr(), ws.data_ptr(), ws.numel(), st), "lovasz_forward_logits")
        ctx.desc, ctx.n_seg = desc, n_seg
        ctx.save_for_backward(logits, stats, jgrad, seg_meta)
        return loss

    @staticmethod
    def backward(ctx, g_loss):
        logits, stats, jgrad, seg_meta = ctx.saved_tensors
        dev = logits.device
        desc, n_seg = ctx.desc, ctx.n_seg
        n, c = logits.shape[0], logits.shape[1]
        hw = logits[0, 0].numel()
        with torch.cuda.device(dev):
            st = stream_ptr(dev)
            g_loss = g_loss.to(torch.float32).contiguous()
            scale = torch.empty(max(n_seg, 1), dtype=torch.float32, device=dev)
            check(lib.b200ssl_lovasz_seg_scale(C.byref(desc), g_loss.data_ptr(), seg_meta[0].data_ptr(),
                                               seg_meta[1].data_ptr(), scale.data_ptr(), st), "lovasz_seg_scale")
            grad = torch.empty_like(jgrad)
            check(lib.b200ssl_lovasz_backward(C.byref(desc), scale.data_ptr(), jgrad.data_ptr(), grad.data_ptr(), st),
                  "lovasz_backward")
            check(lib.b200ssl_softmax_backward(logits.data_ptr(), stats[0].data_ptr(), stats[1].data_ptr(),
                                               grad.data_ptr(), n, c, hw, st), "softmax_backward")
        return grad, None, None


class _LovaszViaProbas(torch.autograd.Function):
    """logits -> lovasz_softmax(F.softmax(logits, 1), labels) with the soft-max written out ONCE by
    b200ssl_softmax_forward and the loss taken on the probability path (exact tail pruning, one-byte label copy);
    backward = segment scales, then b200ssl_softmax_backward_probas in place.  Faster than the never-materialising
    route above (0.58 vs 0.94 ms at 4x21x512x512) for one [N,C,H,W] tensor kept until the backward pass."""

    @staticmethod
    def forward(ctx, logits, labels, desc):
        dev = logits.device
        n, c = logits.shape[0], logits.shape[1]
        hw = logits[0, 0].numel()
        n_seg = lib.b200ssl_lovasz_num_segments(C.byref(desc))
        if n_seg < 0:
            check(n_seg, "lovasz_num_segments")
        loss = torch.empty((), dtype=torch.float32, device=dev)
        seg_loss = torch.empty(max(n_seg, 1), dtype=torch.float32, device=dev)
        seg_meta = torch.empty((2, max(n_seg, 1)), dtype=torch.int32, device=dev)
        probas = torch.empty_like(logits)
        jgrad = torch.empty_like(logits)
        ws = _lib.workspaces.get(dev, "lovasz", lib.b200ssl_lovasz_workspace_bytes(C.byref(desc)))
        with torch.cuda.device(dev):
            st = stream_ptr(dev)
            check(lib.b200ssl_softmax_forward(logits.data_ptr(), n, c, hw, probas.data_ptr(), st), "softmax_forward")
            check(lib.b200ssl_lovasz_forward(
                C.byref(desc), probas.data_ptr(), labels.data_ptr(), loss.data_ptr(), seg_loss.data_ptr(),
                seg_meta[0].data_ptr(), seg_meta[1].data_ptr(), jgrad.data_ptr(), ws.data_ptr(), ws.numel(), st),
                "lovasz_forward")
        ctx.desc, ctx.n_seg = desc, n_seg
        ctx.save_for_backward(probas, jgrad, seg_meta)
        return loss

    @staticmethod
    def backward(ctx, g_loss):
        probas, jgrad, seg_meta = ctx.saved_tensors
        dev = probas.device
        desc, n_seg = ctx.desc, ctx.n_seg
        n, c = probas.shape[0], probas.shape[1]
        hw = probas[0, 0].numel()
        with torch.cuda.device(dev):
            st = stream_ptr(dev)
            g_loss = g_loss.to(torch.float32).contiguous()
            scale = torch.empty(max(n_seg, 1), dtype=torch.float32, device=dev)
            check(lib.b200ssl_lovasz_seg_scale(C.byref(desc), g_loss.data_ptr(), seg_meta[0].data_ptr(),
                                               seg_meta[1].data_ptr(), scale.data_ptr(), st), "lovasz_seg_scale")
            grad = torch.empty_like(jgrad)
            check(lib.b200ssl_lovasz_backward(C.byref(desc), scale.data_ptr(), jgrad.data_ptr(), grad.data_ptr(), st),
                  "lovasz_backward")
            check(lib.b200ssl_softmax_backward_probas(probas.data_ptr(), grad.data_ptr(), n, c, hw, st),
                  "softmax_backward_probas")
        return grad, None, None


def lovasz_softmax_with_logits(logits, labels, classes='present', per_image=False, ignore=None, materialize=True):
    """`lovasz_softmax(F.softmax(logits, dim=1), labels, classes, per_image, ignore)` (lovasz.py:155-160's
    contract: "probas ... typically the output of a softmax") computed from the logits; the gradient flows
    to the logits.  Same degenerate-input behaviour as `lovasz_softmax`.
    materialize=True (default, faster): the probabilities are written once into a scratch tensor that lives
    until the backward pass; materialize=False: they are formed in registers inside the key-build and never
    stored (two [N,H,W] statistics planes instead of one [N,C,H,W] tensor)."""
    logits, labels = _prepare(logits, labels)
    if logits.shape[1] < 2:
        raise ValueError("lovasz_softmax_with_logits needs at least 2 channels (use lovasz_softmax for sigmoid outputs)")
    if not isinstance(classes, str) and len(classes) == 0:
        return 0
    if logits.shape[0] == 0 or logits[0, 0].numel() == 0:
        if per_image and logits.shape[0] == 0:
            return 0
        return logits.permute(0, 2, 3, 1).reshape(-1, logits.shape[1]) * 0.
    desc = _make_desc(logits, labels, classes, per_image, ignore)
    return (_LovaszViaProbas if materialize else _LovaszFromLogits).apply(logits, labels, desc)


# --------------------------- IoU helpers (lovasz.py:34-73), from the confusion matrix -----------
def _mean(values, empty=0):
    values = list(values)
    if not values:
        return empty
    acc = values[0]
    for v in values[1:]:
        acc += v
    return acc if len(values) == 1 else acc / len(values)


def iou(preds, labels, C, EMPTY=1., ignore=None, per_image=False):
    """Array of IoU for each (non ignored) class -- lovasz.py:54-73, computed from one
    (C+1)x(C+1) confusion matrix per image instead of 2*C masked reductions."""
    from .metrics import confusion_matrix
    cm = confusion_matrix(labels, preds, C, ignore_index=ignore, per_image=True, other_bucket=True)
    if not per_image:
        cm = cm.sum(0, keepdim=True)
    cm = cm.cpu().numpy()
    ious = []
    for m in cm:
        row = []
        for i in range(C):
            if i != ignore:
                inter = int(m[i, i])
                union = int(m[i, :].sum() + m[:, i].sum() - m[i, i])
                row.append(EMPTY if not union else float(inter) / float(union))
        ious.append(row)
    ious = [_mean(x) for x in zip(*ious)]
    return 100 * np.array(ious)


def iou_binary(preds, labels, EMPTY=1., ignore=None, per_image=True):
    """IoU for foreground class (1 foreground, 0 background) -- lovasz.py:34-51."""
    from .metrics import confusion_matrix
    cm = confusion_matrix(labels, preds, 2, ignore_index=ignore, per_image=True, other_bucket=True)
    if not per_image:
        cm = cm.sum(0, keepdim=True)
    cm = cm.cpu().numpy()
    vals = []
    for m in cm:
        inter = int(m[1, 1])
        union = int(m[1, :].sum() + m[:, 1].sum() - m[1, 1])
        vals.append(EMPTY if not union else float(inter) / float(union))
    return 100 * _mean(vals)
