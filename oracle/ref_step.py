"""oracle.ref_step -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

One loss-path step driven through the UNMODIFIED reference modules staged in oracle/_ref/ by
oracle/stage_ref.py (the reference's own cowmix / losses / lovasz / mean_teacher functions, called in the
order of train.py:65-130).  Used by bench.py as the `--impl reference` arm and the `cpu_baseline` leg with
`kind: "reference"`, and by tests/test_ref_step.py to pin oracle/torch_port.py on the same call sequence.
The only restated piece is the confusion matrix, for which the reference has no function (SURVEY a15:
`bincount(l*C+p)`), and the two-module holder that update_ema_variables' signature needs.
"""
import importlib
import os
import sys
import warnings

import torch

from . import stage_ref

_mods = None


def modules():
    """Import the staged reference modules (cowmix, losses, lovasz, mean_teacher, metrics)."""
    global _mods
    if _mods is None:
        if not stage_ref.available():
            raise ImportError("oracle/_ref is not staged (run `python -m oracle.stage_ref` where /root/reference is mounted)")
        sys.dont_write_bytecode = True
        if stage_ref.REF_DIR not in sys.path:
            sys.path.insert(0, stage_ref.REF_DIR)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")          # the reference has `is 'present'` and invalid escapes
            _mods = {n: importlib.import_module(n) for n in ("cowmix", "lovasz", "losses", "mean_teacher", "metrics")}
        for n, m in _mods.items():
            assert os.path.dirname(os.path.abspath(m.__file__)) == stage_ref.REF_DIR, (n, m.__file__)
    return _mods


class _Holder(torch.nn.Module):
    """update_ema_variables(model, ema_model, alpha) walks .parameters() / .buffers() (mean_teacher.py:10-18)."""

    def __init__(self, tensors):
        super().__init__()
        self.ps = torch.nn.ParameterList([torch.nn.Parameter(t, requires_grad=False) for t in tensors])


_holders = {}


def _holder(tensors):
    """one holder per tensor list (built once, like the two models of the training loop)"""
    key = id(tensors)
    h = _holders.get(key)
    if h is None or h[0] is not tensors:
        if len(_holders) > 8:
            _holders.clear()
        h = _holders[key] = (tensors, _Holder(tensors))
    return h[1]


def confusion_matrix(labels, preds, num_classes, ignore_index=None):
    """SURVEY a15 (no reference function): bincount(l*C+p) over valid pixels."""
    l, p = labels.reshape(-1), preds.reshape(-1)
    if ignore_index is not None:
        keep = l != ignore_index
        l, p = l[keep], p[keep]
    return torch.bincount(l * num_classes + p, minlength=num_classes * num_classes).view(num_classes, num_classes)


def loss_path_step(image_a, image_b, teacher_a, teacher_b, scores, target, params, ema_params, mode="binary",
                   mask_proportion_range=(0.45, 0.55), sigma_range=(8, 32), alpha=0.99, classes="present",
                   per_image=False, ignore=255, num_classes=2):
    """train.py:77-86 (mask, two mixes), losses.py:239-250 or lovasz.py:155-170 forward + backward,
    mean_teacher.py:5-18, then the confusion matrix of (labels, argmax scores)."""
    m = modules()
    with torch.no_grad():
        mask = m["cowmix"].generate_cowmix_masks_like(image_a, mask_proportion_range=mask_proportion_range,
                                                      sigma_range=sigma_range)
        mixed_teacher = m["cowmix"].mix_with_mask(teacher_a, teacher_b, mask)
        mixed_images = m["cowmix"].mix_with_mask(image_a, image_b, mask)
    s = scores.detach().clone().requires_grad_(True)
    if mode == "binary":
        loss = m["losses"].binary_lovasz_loss_with_logits(s, target)
        labels = torch.argmax(target, dim=1)
    else:
        loss = m["lovasz"].lovasz_softmax(s, target, classes=classes, per_image=per_image, ignore=ignore)
        labels = target
    loss.backward()
    m["mean_teacher"].update_ema_variables(_holder(params), _holder(ema_params), alpha)
    with torch.no_grad():
        cm = confusion_matrix(labels, torch.argmax(scores, dim=1), num_classes, ignore_index=ignore)
    return {"mask": mask, "mixed_images": mixed_images, "mixed_teacher": mixed_teacher,
            "loss": loss.detach(), "grad": s.grad, "cm": cm}
