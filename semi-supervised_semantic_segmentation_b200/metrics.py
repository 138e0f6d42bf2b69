"""Drop-in for the reference's metrics.py plus the confusion-matrix mIoU the north star asks for.

    reference metrics.py:1-7    dice_metric(input, target)
    (new)                       confusion_matrix / confusion_matrix_from_logits / miou_from_cm

There is no confusion matrix in the reference; its definition here is
`bincount(label*C + pred, minlength=C*C)` over the non-ignored pixels, and the derived IoU / Dice
equal lovasz.iou / metrics.dice_metric exactly.
"""
import torch

from . import _lib
from ._lib import lib, check, stream_ptr, require_cuda


def dice_metric(input, target):
    """metrics.py:1-7: per-sample (2*sum(x*y)+1)/(sum(x+y)+1) over dims (1,2,3)."""
    require_cuda(input, "input", torch.float32)
    require_cuda(target, "target", torch.float32)
    if input.shape != target.shape:
        input, target = torch.broadcast_tensors(input, target)
    if input.dim() != 4:
        raise IndexError("Dimension out of range (dice_metric sums over dims (1, 2, 3))")
    input, target = input.contiguous(), target.contiguous()
    n = input.shape[0]
    chw = input[0].numel() if n else 0
    out = torch.empty(n, dtype=torch.float32, device=input.device)
    if n == 0:
        return out
    ws_bytes = lib.b200ssl_dice_workspace_bytes(n, chw)
    ws = _lib.workspaces.get(input.device, "dice", ws_bytes)
    with torch.cuda.device(input.device):
        check(lib.b200ssl_dice_metric(input.data_ptr(), target.data_ptr(), n, chw, out.data_ptr(),
                                      ws.data_ptr(), ws.numel(), stream_ptr(input.device)), "dice_metric")
    return out


def confusion_matrix(labels, preds, num_classes, ignore_index=None, per_image=False,
                     other_bucket=False, out=None, return_dropped=False):
    """int64 confusion matrix cm[label, pred] ([C,C], or [N,C,C] with per_image).

    labels/preds: integer tensors of identical shape ([N,H,W] or flat), int64 / int32 / uint8.
    `out` accumulates into an existing matrix (streaming evaluation).  With other_bucket the matrix
    is (C+1)x(C+1) and out-of-range labels/predictions are counted in the last row/column.
    """
    require_cuda(labels, "labels")
    require_cuda(preds, "preds")
    if labels.shape != preds.shape:
        raise ValueError(f"labels {tuple(labels.shape)} and preds {tuple(preds.shape)} differ in shape")
    if labels.dtype != preds.dtype:
        raise TypeError("labels and preds must have the same integer dtype")
    code = _lib.label_dtype_code(labels)
    labels, preds = labels.contiguous(), preds.contiguous()
    d = num_classes + (1 if other_bucket else 0)
    n_img = labels.shape[0] if (per_image and labels.dim() > 1) else 1
    hw = labels[0].numel() if (per_image and labels.dim() > 1 and n_img) else labels.numel()
    shape = (n_img, d, d) if per_image else (d, d)
    if out is None:
        out = torch.zeros(shape, dtype=torch.int64, device=labels.device)
    else:
        require_cuda(out, "out", torch.int64)
        if tuple(out.shape) != shape or not out.is_contiguous():
            raise ValueError(f"out must be a contiguous int64 tensor of shape {shape}")
    dropped = torch.zeros((), dtype=torch.int64, device=labels.device) if return_dropped else None
    with torch.cuda.device(labels.device):
        check(lib.b200ssl_confusion_matrix(
            labels.data_ptr(), preds.data_ptr(), labels.numel(), num_classes, 1 if other_bucket else 0,
            0 if ignore_index is None else 1, 0 if ignore_index is None else int(ignore_index), code,
            1 if per_image else 0, hw, out.data_ptr(),
            dropped.data_ptr() if return_dropped else None, stream_ptr(labels.device)), "confusion_matrix")
    return (out, dropped) if return_dropped else out


def confusion_matrix_from_logits(logits, labels, ignore_index=None, per_image=False, out=None):
    """argmax over the channel dim fused into the histogram: logits [N,C,H,W] fp32, labels [N,H,W]."""
    require_cuda(logits, "logits", torch.float32)
    require_cuda(labels, "labels")
    if logits.dim() != 4 or labels.dim() != 3 or labels.shape[0] != logits.shape[0] or \
            labels[0].numel() != logits[0, 0].numel():
        raise ValueError(f"logits {tuple(logits.shape)} / labels {tuple(labels.shape)} mismatch")
    code = _lib.label_dtype_code(labels)
    logits, labels = logits.contiguous(), labels.contiguous()
    n, c = logits.shape[0], logits.shape[1]
    hw = labels[0].numel() if n else 0
    shape = (n, c, c) if per_image else (c, c)
    if out is None:
        out = torch.zeros(shape, dtype=torch.int64, device=logits.device)
    elif tuple(out.shape) != shape or out.dtype != torch.int64 or not out.is_contiguous():
        raise ValueError(f"out must be a contiguous int64 tensor of shape {shape}")
    with torch.cuda.device(logits.device):
        check(lib.b200ssl_confusion_from_logits(
            logits.data_ptr(), labels.data_ptr(), n, c, hw, 0 if ignore_index is None else 1,
            0 if ignore_index is None else int(ignore_index), code, 1 if per_image else 0,
            out.data_ptr(), None, stream_ptr(logits.device)), "confusion_from_logits")
    return out


def iou_from_cm(cm, EMPTY=1.0):
    """Per-class IoU = TP / (row + col - TP) with lovasz.iou's EMPTY convention; cm [..., C, C]."""
    cm = cm.to(torch.float64)
    tp = torch.diagonal(cm, dim1=-2, dim2=-1)
    union = cm.sum(-1) + cm.sum(-2) - tp
    return torch.where(union > 0, tp / union.clamp(min=1), torch.full_like(tp, EMPTY))


def miou_from_cm(cm, EMPTY=1.0):
    return iou_from_cm(cm, EMPTY).mean(-1)


def dice_from_cm(cm_per_image):
    """Dice of metrics.py:1-7 from per-image 2x2 matrices [N,2,2] (foreground = class 1)."""
    require_cuda(cm_per_image, "cm_per_image", torch.int64)
    if cm_per_image.dim() != 3 or cm_per_image.shape[1:] != (2, 2):
        raise ValueError("cm_per_image must be [N,2,2]")
    cm_per_image = cm_per_image.contiguous()
    n = cm_per_image.shape[0]
    out = torch.empty(n, dtype=torch.float32, device=cm_per_image.device)
    with torch.cuda.device(cm_per_image.device):
        check(lib.b200ssl_dice_from_cm(cm_per_image.data_ptr(), n, out.data_ptr(),
                                       stream_ptr(cm_per_image.device)), "dice_from_cm")
    return out


def validation_dice(pred_logits, mask, threshold=0.5, fg_class=1, cm_out=None):
    """The validation metric of train.py:171-175 in one pass over the mask:

        one_hot = F.one_hot(argmax(pred_logits, 1), 2).permute(0, 3, 1, 2)
        pred_map_binary = F.interpolate(one_hot, size=mask.shape[2:], mode='nearest')
        metrics.dice_metric(pred_map_binary[:, 1:], (mask > 0.5)[:, 1:])

    pred_logits: [N,C,h,w] fp32 at the network's resolution; mask: [N,Cm,H,W] fp32 (soft) labels.
    Returns (dice [N] fp32 -- take .mean() for train.py:175 --, per-image 2x2 matrices [N,2,2] int64
    {TN, FP; FN, TP}); `cm_out` accumulates into existing matrices."""
    require_cuda(pred_logits, "pred_logits", torch.float32)
    require_cuda(mask, "mask", torch.float32)
    if pred_logits.dim() != 4 or mask.dim() != 4 or pred_logits.shape[0] != mask.shape[0]:
        raise ValueError(f"pred_logits {tuple(pred_logits.shape)} / mask {tuple(mask.shape)} mismatch")
    if not 0 <= fg_class < mask.shape[1]:
        raise IndexError(f"class {fg_class} is not a channel of the mask")
    pred_logits, mask = pred_logits.contiguous(), mask.contiguous()
    n, c, h, w = pred_logits.shape
    if cm_out is None:
        cm_out = torch.zeros((n, 2, 2), dtype=torch.int64, device=mask.device)
    elif tuple(cm_out.shape) != (n, 2, 2) or cm_out.dtype != torch.int64 or not cm_out.is_contiguous():
        raise ValueError("cm_out must be a contiguous int64 tensor of shape [N,2,2]")
    with torch.cuda.device(mask.device):
        check(lib.b200ssl_validation_cm(pred_logits.data_ptr(), n, c, h, w, mask.data_ptr(), mask.shape[1],
                                        mask.shape[2], mask.shape[3], float(threshold), int(fg_class),
                                        cm_out.data_ptr(), stream_ptr(mask.device)), "validation_cm")
    return dice_from_cm(cm_out), cm_out
