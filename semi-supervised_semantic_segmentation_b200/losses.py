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
        """losses.py:15-22.  Row N2: a loss_fn that declares `accepts_lowres` (this module's
        binary_lovasz_loss_with_logits) is handed the prediction at ITS resolution and interpolates inside its
        own kernels; the full-resolution prediction is only materialised when some other loss_fn needs it."""
        loss = 0
        for prediction_idx, prediction in enumerate(predictions_list):
            size = (target.size(2), target.size(3))
            full = None
            for loss_spec in self.losses:
                fn = loss_spec['loss_fn']
                if getattr(fn, "accepts_lowres", False) and prediction.is_cuda and _lowres_ok(prediction, target):
                    value = fn(prediction, target)
                else:
                    if full is None:
                        full = torch.nn.functional.interpolate(prediction, size=size, mode='bilinear',
                                                               align_corners=False)
                    value = fn(full, target)
                loss += value * loss_spec['weight'][prediction_idx]
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


_MAX_UPSAMPLE_RATIO = 9.0   # csrc/lovasz.cu: kMaxBackWin


def _lowres_ok(prediction, target):
    """shapes the fused low-resolution front end takes (else the caller interpolates with torch)"""
    return (prediction.dim() == 4 and target.dim() == 4 and prediction.dtype == torch.float32 and
            prediction.shape[:2] == target.shape[:2] and prediction.shape[1] >= 2 and target.shape[3] % 4 == 0 and
            1 <= prediction.shape[2] <= target.shape[2] and 1 <= prediction.shape[3] <= target.shape[3] and
            target.shape[3] / prediction.shape[3] <= _MAX_UPSAMPLE_RATIO and
            (target.shape[2] != prediction.shape[2] or target.shape[3] != prediction.shape[3]))


class _BinaryLovaszLowres(torch.autograd.Function):
    """CalculateLoss's bilinear resize (losses.py:18-19) + binary_lovasz_loss_with_logits (:239-250) from
    low-resolution logits in one chain of kernels (b200ssl_binary_lovasz_lowres).  The loss is linear in its
    upstream gradient, so forward computes d loss / d input for an upstream 1 and backward scales it."""

    @staticmethod
    def forward(ctx, input_low, target):
        dev = input_low.device
        n, c, lh, lw = input_low.shape
        H, W = target.shape[2], target.shape[3]
        desc = _lib.LovaszDesc()
        desc.n_images, desc.n_channels, desc.hw = n, 1, H * W
        desc.per_image, desc.class_mode, desc.n_list = 1, _lib.LOVASZ_LIST, 1
        desc.class_list[0] = 1
        desc.has_ignore, desc.ignore_index, desc.label_dtype = 1, 255, _lib.U8
        ws = _lib.workspaces.get(dev, "lovasz", lib.b200ssl_lovasz_workspace_bytes(C.byref(desc)))
        small = torch.empty(3, dtype=torch.float32, device=dev)                 # loss, denom, upstream 1.0
        small[2] = 1.0
        segf = torch.empty(n, dtype=torch.float32, device=dev)
        segi = torch.empty(3 * n, dtype=torch.int32, device=dev)                # seg_fg | seg_valid | nonzero
        labels = torch.empty((n, H, W), dtype=torch.uint8, device=dev)
        grad_full = torch.empty((n, H, W), dtype=torch.float32, device=dev)     # d loss / d up-sampled logit of class 1
        grad_low = torch.empty_like(input_low)
        with torch.cuda.device(dev):
            check(lib.b200ssl_binary_lovasz_lowres(
                input_low.data_ptr(), target.data_ptr(), n, c, lh, lw, H, W, 1, small.data_ptr() + 8,
                labels.data_ptr(), segi.data_ptr() + 8 * n, small.data_ptr(), small.data_ptr() + 4, segf.data_ptr(),
                segi.data_ptr(), segi.data_ptr() + 4 * n, grad_full.data_ptr(), grad_low.data_ptr(), ws.data_ptr(),
                ws.numel(), stream_ptr(dev)), "binary_lovasz_lowres")
        ctx.save_for_backward(grad_low)
        return small[0]

    @staticmethod
    def backward(ctx, g):
        (grad_low,) = ctx.saved_tensors
        return grad_low * g.to(torch.float32), None


def binary_lovasz_loss_with_logits(input, target):
    """losses.py:239-250.  `input` are raw logits (the reference's sigmoid is commented out),
    `target` a soft one-hot [N,C,H,W]; class 1 only, void label 255, one Lovasz problem per image,
    images without any non-zero label get weight 0.
    Row N2: `input` may be at a LOWER resolution than `target` (the network's stride-4 logits); it is then
    bilinearly up-sampled (align_corners=False, as losses.py:18-19 does before calling the loss) inside the
    front end of the sort and the gradient comes back at the input's resolution."""
    require_cuda(input, "input", torch.float32)
    if input.shape[0] == 0:
        raise ValueError("binary_lovasz_loss_with_logits needs a non-empty batch")
    if input.dim() == 4 and target.dim() == 4 and input.shape[2:] != target.shape[2:]:
        require_cuda(target, "target", torch.float32)
        if _lowres_ok(input, target):
            return _BinaryLovaszLowres.apply(input.contiguous(), target.contiguous())
        input = torch.nn.functional.interpolate(input, size=(target.size(2), target.size(3)), mode='bilinear',
                                                align_corners=False)
    labels, nonzero = argmax_channels(target)                    # int_target, (tgt.sum() > 0)
    seg_loss, _ = lovasz.lovasz_segment_losses(input, labels, classes=[1], per_image=True, ignore=255)
    return _BinaryReduce.apply(seg_loss.reshape(-1), nonzero)


binary_lovasz_loss_with_logits.accepts_lowres = True
