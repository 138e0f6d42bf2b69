"""Confidence-masked consistency loss of the semi-supervised branch ("next" row N1).

The reference has no function for it -- the code is inline in train.py:98-107:

    mixed_ema_pred_prob     = torch.sigmoid(mixed_ema_pred)
    mixed_student_pred_prob = torch.sigmoid(mixed_student_pred)
    confidence_modulator = (mixed_ema_pred_prob.max(dim=1).values > confidence_threshold).to(...)
    consistency_loss = torch.pow(mixed_student_pred_prob - mixed_ema_pred_prob, exponent=2.0)
    consistency_loss = (consistency_loss.sum(dim=1) * confidence_modulator).sum() / confidence_modulator.sum()
    confidence_modulator = confidence_modulator.mean()

`confidence_masked_consistency` returns (consistency_loss, confidence_modulator.mean()) from one fused
pass over both logit tensors; the backward (w.r.t. the student logits only -- the teacher is computed
under no_grad in the reference) is one more pass.  Replaces ~12 ATen kernels and 6 temporaries.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import lib, check, stream_ptr, require_cuda


class _Consistency(torch.autograd.Function):
    @staticmethod
    def forward(ctx, student, teacher, threshold):
        dev = student.device
        n, c = student.shape[0], student.shape[1]
        hw = student[0, 0].numel()
        stats = torch.empty(3, dtype=torch.float32, device=dev)
        ws = _lib.workspaces.get(dev, "consistency", lib.b200ssl_consistency_workspace_bytes(n, hw))
        with torch.cuda.device(dev):
            check(lib.b200ssl_consistency_forward(student.data_ptr(), teacher.data_ptr(), n, c, hw, float(threshold),
                                                  stats.data_ptr(), ws.data_ptr(), ws.numel(), stream_ptr(dev)),
                  "consistency_forward")
        ctx.save_for_backward(student, teacher, stats)
        ctx.threshold = float(threshold)
        ctx.mark_non_differentiable(stats)
        return stats[0], stats

    @staticmethod
    def backward(ctx, g_loss, _g_stats):
        student, teacher, stats = ctx.saved_tensors
        if g_loss is None or not ctx.needs_input_grad[0]:
            return None, None, None
        dev = student.device
        n, c = student.shape[0], student.shape[1]
        hw = student[0, 0].numel()
        g = g_loss.to(torch.float32).contiguous()
        grad = torch.empty_like(student)
        with torch.cuda.device(dev):
            check(lib.b200ssl_consistency_backward(student.data_ptr(), teacher.data_ptr(), n, c, hw, ctx.threshold,
                                                   stats.data_ptr(), g.data_ptr(), grad.data_ptr(), stream_ptr(dev)),
                  "consistency_backward")
        return grad, None, None


def confidence_masked_consistency(mixed_student_pred, mixed_ema_pred, confidence_threshold):
    """train.py:98-107.  Both inputs are raw logits [N,C,H,W] (fp32, CUDA).  Returns
    (consistency_loss, confidence_modulator_mean) as 0-dim tensors; the loss is differentiable w.r.t.
    `mixed_student_pred`.  As in the reference the loss is NaN (0/0) when no pixel is confident."""
    require_cuda(mixed_student_pred, "mixed_student_pred", torch.float32)
    require_cuda(mixed_ema_pred, "mixed_ema_pred", torch.float32)
    if mixed_student_pred.dim() != 4 or mixed_student_pred.shape != mixed_ema_pred.shape:
        raise ValueError("student and teacher predictions must be [N,C,H,W] tensors of the same shape")
    if mixed_student_pred.numel() == 0:
        raise ValueError("empty predictions")
    student = mixed_student_pred.contiguous()
    teacher = mixed_ema_pred.detach().contiguous()
    loss, stats = _Consistency.apply(student, teacher, confidence_threshold)
    return loss, stats[2]


class _ConsistencyMixed(torch.autograd.Function):
    @staticmethod
    def forward(ctx, student, teacher_a, teacher_b, mask, threshold):
        dev = student.device
        n, c, h, w = student.shape
        th, tw = teacher_a.shape[2], teacher_a.shape[3]
        stats = torch.empty(3, dtype=torch.float32, device=dev)
        conf = torch.empty((n, h, w), dtype=torch.uint8, device=dev)     # the decision per pixel, for the backward
        ws = _lib.workspaces.get(dev, "consistency", lib.b200ssl_consistency_mixed_workspace_bytes(n, h, w))
        with torch.cuda.device(dev):
            check(lib.b200ssl_consistency_mixed_forward(
                student.data_ptr(), teacher_a.data_ptr(), teacher_b.data_ptr(), mask.data_ptr(), n, c, h, w, th, tw,
                float(threshold), stats.data_ptr(), conf.data_ptr(), ws.data_ptr(), ws.numel(), stream_ptr(dev)),
                "consistency_mixed_forward")
        ctx.save_for_backward(student, teacher_a, teacher_b, mask, stats, conf)
        ctx.threshold = float(threshold)
        ctx.mark_non_differentiable(stats)
        return stats[0], stats

    @staticmethod
    def backward(ctx, g_loss, _g_stats):
        student, teacher_a, teacher_b, mask, stats, conf = ctx.saved_tensors
        if g_loss is None or not ctx.needs_input_grad[0]:
            return None, None, None, None, None
        dev = student.device
        n, c, h, w = student.shape
        th, tw = teacher_a.shape[2], teacher_a.shape[3]
        g = g_loss.to(torch.float32).contiguous()
        grad = torch.empty_like(student)
        with torch.cuda.device(dev):
            check(lib.b200ssl_consistency_mixed_backward(
                student.data_ptr(), teacher_a.data_ptr(), teacher_b.data_ptr(), mask.data_ptr(), n, c, h, w, th, tw,
                ctx.threshold, stats.data_ptr(), conf.data_ptr(), g.data_ptr(), grad.data_ptr(), stream_ptr(dev)),
                "consistency_mixed_backward")
        return grad, None, None, None, None


def confidence_masked_consistency_mixed(mixed_student_pred, ema_pred_a, ema_pred_b, mask, confidence_threshold, fused=None):
    """train.py:69-82 + 98-107 with the teacher formed on the fly:

        ema_pred_x      = F.interpolate(ema_pred_x, size, mode='bilinear', align_corners=False)   # if below `size`
        mixed_ema_pred  = cowmix.mix_with_mask(ema_pred_a, ema_pred_b, mask)
        loss, conf_mean = confidence_masked_consistency(mixed_student_pred, mixed_ema_pred, confidence_threshold)

    `ema_pred_a` / `ema_pred_b`: the teacher's raw logits [N,C,h,w] at the network's resolution (h x w <= H x W of the
    student prediction; equal sizes are read as they are), `mask` the CowMix mask [N,1,H,W].  The mixed teacher
    prediction is never materialised; values and gradients equal the three-call route (gradients bit for bit).
    fused=None chooses by the channel count: forming the teacher in registers costs one more interpolation of both
    teacher tensors in the backward pass, which pays for few channels (1.22x over mix + loss at C = 2, the reference's
    configuration) and does not at 19 (0.84x: the interpolation is instruction-bound, benchmarks/consistency_mixed.py);
    above 4 channels the wrapper therefore materialises the mixed teacher with ONE mix2_upsampled launch and calls the
    two-tensor loss.  fused=True / False force either route (same results)."""
    require_cuda(mixed_student_pred, "mixed_student_pred", torch.float32)
    require_cuda(ema_pred_a, "ema_pred_a", torch.float32)
    require_cuda(ema_pred_b, "ema_pred_b", torch.float32)
    require_cuda(mask, "mask", torch.float32)
    s = mixed_student_pred
    if s.dim() != 4 or ema_pred_a.dim() != 4 or ema_pred_a.shape != ema_pred_b.shape or \
            ema_pred_a.shape[:2] != s.shape[:2]:
        raise ValueError("student [N,C,H,W] and teacher predictions [N,C,h,w] (both teachers alike) expected")
    if tuple(mask.shape) != (s.shape[0], 1, s.shape[2], s.shape[3]):
        raise ValueError(f"mask must be [N,1,H,W] = {(s.shape[0], 1, s.shape[2], s.shape[3])}, got {tuple(mask.shape)}")
    if s.numel() == 0 or ema_pred_a.numel() == 0:
        raise ValueError("empty predictions")
    if ema_pred_a.shape[2] > s.shape[2] or ema_pred_a.shape[3] > s.shape[3]:
        raise ValueError("the teacher predictions must not be larger than the student's")
    if fused is None:
        fused = s.shape[1] <= 4
    if not fused:
        from . import cowmix
        dummy = mask.detach()                      # first pair of the fused mix: one channel, discarded
        _, mixed = cowmix.mix2_with_mask(dummy, dummy, ema_pred_a.detach(), ema_pred_b.detach(), mask.detach())
        return confidence_masked_consistency(s, mixed, confidence_threshold)
    loss, stats = _ConsistencyMixed.apply(s.contiguous(), ema_pred_a.detach().contiguous(),
                                          ema_pred_b.detach().contiguous(), mask.detach().contiguous(),
                                          confidence_threshold)
    return loss, stats[2]
