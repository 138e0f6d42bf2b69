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
