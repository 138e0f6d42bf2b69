"""The reference's training step (train.py:47-130: supervised Lovasz loss, teacher forward + up-sampling,
CowMix mask, two mixes, student forward, confidence-masked consistency loss, backward, gradient clip,
SGD step, zero_grad, EMA) restated twice around the SAME tiny CNNs on the same GPU:

    reference flavour : the ATen operator sequence the reference issues (oracle.torch_port, torch.optim.SGD,
                        torch.nn.utils.clip_grad_norm_, F.interpolate)
    b200ssl flavour   : cowmix.masks_from_noise / mix2_with_mask with low-resolution teacher logits (N2),
                        losses.binary_lovasz_loss_with_logits, consistency.confidence_masked_consistency (N1),
                        optim.FusedSGD.step(max_grad_norm, ema_params, zero_grad) (N4)

Same noise, p, sigma and initial weights; after three steps the student's and the teacher's parameter
UPDATES agree to 1e-3 in relative L2 norm (the backbone's cuDNN kernels and the mask's documented margin
are shared or negligible; every replaced op is individually held to 1e-5 / bit-exact elsewhere)."""
import copy

import pytest
import torch
import torch.nn.functional as F

from oracle import torch_port

pytestmark = pytest.mark.gpu


class TinyNet(torch.nn.Module):
    """stride-4 logits like HRNet's head; returns (features, [pred_maps]) like train.py:47."""

    def __init__(self, classes=2):
        super().__init__()
        self.c1 = torch.nn.Conv2d(3, 8, 3, stride=2, padding=1)
        self.c2 = torch.nn.Conv2d(8, classes, 3, stride=2, padding=1)

    def forward(self, x):
        f = F.relu(self.c1(x))
        return f, [self.c2(f) * 4.0]


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def test_three_training_steps_match_the_reference_sequence():
    import b200ssl
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(21)
    n, c, h, w = 4, 2, 64, 64
    thr, weight, alpha, clip = 0.6, 10.0, 0.99, 5.0
    hp = dict(lr=0.05, momentum=0.9, weight_decay=0.0005)
    student0 = TinyNet(c).to(dev)
    teacher0 = copy.deepcopy(student0)
    with torch.no_grad():
        for p in teacher0.parameters():
            p.add_(0.01 * torch.randn(p.shape, generator=g).to(dev))
    steps = []
    for _ in range(3):
        blob = F.avg_pool2d(torch.randn(n, c, h, w, generator=g), 9, 1, 4)
        steps.append(dict(
            image=torch.rand(n, 3, h, w, generator=g).to(dev),
            target=F.one_hot(blob.argmax(1), c).permute(0, 3, 1, 2).float().contiguous().to(dev),
            ua=torch.rand(n, 3, h, w, generator=g).to(dev), ub=torch.rand(n, 3, h, w, generator=g).to(dev),
            noise=torch.randn(n, 1, h, w, generator=g).to(dev),
            p=torch.rand(n, generator=g) * 0.1 + 0.45, sigmas=torch.rand(n, generator=g) * 2 + 2))

    def up(x):
        return F.interpolate(x, (h, w), mode="bilinear", align_corners=False)

    # ---------------- reference flavour ----------------
    student, teacher = copy.deepcopy(student0), copy.deepcopy(teacher0)
    for p in teacher.parameters():
        p.detach_()                                                              # mean_teacher.py:20-22
    opt = torch.optim.SGD(student.parameters(), **hp)
    ref_losses = []
    for s in steps:
        _, preds = student(s["image"])
        sup = torch_port.binary_lovasz_loss_with_logits(up(preds[-1]), s["target"])        # losses.py:15-22,239-250
        sup.backward()
        with torch.no_grad():
            ema_a, ema_b = up(teacher(s["ua"])[-1][-1]), up(teacher(s["ub"])[-1][-1])      # train.py:71-75
            mask = torch_port.masks_from_noise(s["noise"], s["p"], s["sigmas"])           # cowmix.py:56-68
            mixed_ema = torch_port.mix_with_mask(ema_a, ema_b, mask)                        # train.py:82
            mixed_img = torch_port.mix_with_mask(s["ua"], s["ub"], mask)                    # train.py:84-86
        mixed_student = up(student(mixed_img)[-1][-1])                                       # train.py:92-94
        cons, conf = torch_port.confidence_masked_consistency(mixed_student, mixed_ema, thr)  # train.py:98-107
        (cons * weight).backward()
        torch.nn.utils.clip_grad_norm_(student.parameters(), clip)                          # train.py:122
        opt.step()
        opt.zero_grad()
        torch_port.update_ema_variables(student, teacher, alpha)                            # train.py:130
        ref_losses.append((float(sup.detach()), float(cons.detach()), float(conf)))
        ref_mask = mask
    ref_student = [p.detach().clone() for p in student.parameters()]
    ref_teacher = [p.detach().clone() for p in teacher.parameters()]

    # ---------------- b200ssl flavour ----------------
    student, teacher = copy.deepcopy(student0), copy.deepcopy(teacher0)
    b200ssl.mean_teacher.detach_model_parameters(teacher)
    ema_list = list(teacher.parameters())
    opt = b200ssl.optim.FusedSGD(student.parameters(), **hp)
    our_losses = []
    for s in steps:
        _, preds = student(s["image"])
        sup = b200ssl.losses.binary_lovasz_loss_with_logits(up(preds[-1]), s["target"])
        sup.backward()
        with torch.no_grad():
            lo_a, lo_b = teacher(s["ua"])[-1][-1], teacher(s["ub"])[-1][-1]                 # stay at stride 4
            mask = b200ssl.cowmix.masks_from_noise(s["noise"], s["p"], s["sigmas"])
            mixed_img, mixed_ema = b200ssl.cowmix.mix2_with_mask(s["ua"], s["ub"], lo_a.contiguous(), lo_b.contiguous(), mask)
        mixed_student = up(student(mixed_img)[-1][-1])
        cons, conf = b200ssl.consistency.confidence_masked_consistency(mixed_student, mixed_ema, thr)
        (cons * weight).backward()
        opt.step(max_grad_norm=clip, ema_params=ema_list, ema_alpha=alpha, zero_grad=True)
        our_losses.append((float(sup.detach()), float(cons.detach()), float(conf)))
        assert all(not bool(p.grad.any()) for p in student.parameters())
    assert int((mask != ref_mask).sum()) <= 1
    for (a_sup, a_cons, a_conf), (b_sup, b_cons, b_conf) in zip(our_losses, ref_losses):
        assert abs(a_sup - b_sup) <= 1e-4 * abs(b_sup)
        assert abs(a_cons - b_cons) <= 1e-3 * abs(b_cons)
        assert abs(a_conf - b_conf) <= 2e-3
    for p_ours, p_ref, p0 in zip(student.parameters(), ref_student, student0.parameters()):
        assert rel_l2(p_ours.detach() - p0.detach(), p_ref - p0.detach()) <= 1e-3
    for p_ours, p_ref, p0 in zip(teacher.parameters(), ref_teacher, teacher0.parameters()):
        assert rel_l2(p_ours.detach() - p0.detach(), p_ref - p0.detach()) <= 1e-3
