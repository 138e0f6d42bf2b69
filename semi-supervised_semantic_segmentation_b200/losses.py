"""The Lovasz call shim of the reference's losses.py, on the CUDA path.

    reference losses.py:8-22     CalculateLoss (bilinear resize each prediction, weighted sum)
    reference losses.py:239-250  binary_lovasz_loss_with_logits(input, target)
"""
import ctypes as C

import torch

from . import _lib
from ._lib import lib, check, stream_ptr, require_cuda
from . import lovasz


class CalculateLoss():
    def __init__(self, losses):
        # losses: list of dicts with
        # 'loss_fn': class instance or fn to run like loss_fn(prediction, target)
        # 'weight': per-prediction scalars to multiply the loss with
        self.losses = losses

    def __call__(self, predictions_list, target):
        loss = 0
        for prediction_idx, prediction in enumerate(predictions_list):
            prediction = torch.nn.functional.interpolate(prediction, size=(target.size(2), target.size(3)),
                                                         mode='bilinear', align_corners=False)
            for loss_spec in self.losses:
                loss += loss_spec['loss_fn'](prediction, target) * loss_spec['weight'][prediction_idx]
        return loss


def argmax_channels(target, out_dtype=torch.uint8, count_nonzero=True):
    """`torch.argmax(target, dim=1)` for a soft one-hot target [N,C,H,W], written as compact labels,
    plus the per-image number of non-zero labels (for losses.py:246 `tgt.sum() > 0`)."""
    require_cuda(target, "target", torch.float32)
    if target.dim() != 4:
        raise ValueError("target must be [N,C,H,W]")
    target = target.contiguous()
    n, c = target.shape[0], target.shape[1]
    hw = target[0, 0].numel() if target.numel() else 0
    if out_dtype == torch.uint8 and c > 256:
        out_dtype = torch.int64
    labels = torch.empty((n,) + tuple(target.shape[2:]), dtype=out_dtype, device=target.device)
    nonzero = torch.zeros(max(n, 1), dtype=torch.int32, device=target.device)[:n] if count_nonzero else None
    with torch.cuda.device(target.device):
        check(lib.b200ssl_argmax_channels(
            target.data_ptr(), n, c, hw, labels.data_ptr(),
            _lib.U8 if out_dtype == torch.uint8 else _lib.I64,
            nonzero.data_ptr() if count_nonzero and n else None, stream_ptr(target.device)), "argmax_channels")
    return labels, nonzero


class _BinaryReduce(torch.autograd.Function):
    """loss = sum_i w_i L_i / (sum_i w_i + 0.001) with python's left-to-right fp32 sums."""

    @staticmethod
    def forward(ctx, seg_loss, nonzero):
        dev = seg_loss.device
        n = seg_loss.numel()
        seg_loss = seg_loss.contiguous()
        out = torch.empty(2, dtype=torch.float32, device=dev)  # [loss, denom]
        with torch.cuda.device(dev):
            check(lib.b200ssl_binary_lovasz_reduce(seg_loss.data_ptr(), nonzero.data_ptr(), n,
                                                   out.data_ptr(), out.data_ptr() + 4, stream_ptr(dev)),
                  "binary_lovasz_reduce")
        ctx.save_for_backward(out, nonzero)
        ctx.n = n
        return out[0]

    @staticmethod
    def backward(ctx, g):
        out, nonzero = ctx.saved_tensors
        dev = out.device
        g = g.to(torch.float32).contiguous()
        scale = torch.empty(max(ctx.n, 1), dtype=torch.float32, device=dev)[:ctx.n]
        with torch.cuda.device(dev):
            check(lib.b200ssl_binary_lovasz_scale(g.data_ptr(), nonzero.data_ptr(), out.data_ptr() + 4,
                                                  ctx.n, scale.data_ptr(), stream_ptr(dev)),
                  "binary_lovasz_scale")
        return scale, None


def binary_lovasz_loss_with_logits(input, target):
    """losses.py:239-250.  `input` are raw logits (the reference's sigmoid is commented out),
    `target` a soft one-hot [N,C,H,W]; class 1 only, void label 255, one Lovasz problem per image,
    images without any non-zero label get weight 0."""
    require_cuda(input, "input", torch.float32)
    if input.shape[0] == 0:
        raise ValueError("binary_lovasz_loss_with_logits needs a non-empty batch")
    labels, nonzero = argmax_channels(target)                    # int_target, (tgt.sum() > 0)
    seg_loss, _ = lovasz.lovasz_segment_losses(input, labels, classes=[1], per_image=True, ignore=255)
    return _BinaryReduce.apply(seg_loss.reshape(-1), nonzero)
