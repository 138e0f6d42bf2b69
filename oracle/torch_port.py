"""oracle.torch_port -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The reference's loss-and-mixing path restated with the same ATen (torch CPU) operator sequence
its Python issues, so that (a) it can be diffed against the real reference in the build container
(tests/golden/make_golden.py) and (b) it can be timed on the GPU box's host cores as the
`cpu_baseline` / `--impl reference` arm of bench.py (kind = "port": /root/reference is Python and
does not travel to the GPU box).  Only tests/ and bench.py import this module.

Reference lines followed are cited per function.
"""
import math

import torch
import torch.nn.functional as F


# ---------------------------------------------------------------- cowmix.py ----------------------
def taps_1d(size, sigma):
    """cowmix.py:6-11.  x = -size//2 ... size//2-1 (off-centre for odd size)."""
    x = torch.arange(-size // 2, size // 2).float()
    if size % 2 == 0:
        x = x + 0.5
    g = torch.exp((-x.pow(2.0) / float(2 * sigma ** 2)))
    return g / g.sum()


def smooth_separable(planes, sigmas):
    """cowmix.py:27-37.  planes [1,N,H,W]; one K for the batch; depthwise cross-correlation along
    H then along W with zero padding floor(K/2)."""
    n = sigmas.shape[0]
    assert planes.shape[1] == n
    size = int(round(sigmas.max().item() * 3) * 2) + 1
    col = torch.stack([taps_1d(size, sigmas[i]) for i in range(n)], dim=0).view(n, 1, size, 1).to(planes)
    row = col.transpose(2, 3)
    pad = math.floor(float(size) / 2)
    out = F.conv2d(planes, weight=col, padding=(pad, 0), groups=n)
    return F.conv2d(out, weight=row, padding=(0, pad), groups=n)


def draw_parameters(n, mask_proportion_range, sigma_range):
    """cowmix.py:44-51: p first, then sigma, both from the CPU generator."""
    p = torch.distributions.Uniform(torch.tensor(mask_proportion_range[0]),
                                    torch.tensor(mask_proportion_range[1])).rsample(sample_shape=[n])
    lo, hi = math.log(float(sigma_range[0])), math.log(float(sigma_range[1]))
    sigmas = torch.exp(torch.distributions.Uniform(torch.tensor(lo), torch.tensor(hi)).rsample([n]))
    return p, sigmas


def masks_from_noise(noise, p, sigmas, return_field=False):
    """cowmix.py:56-68 for a given noise field [N,1,H,W]."""
    with torch.no_grad():
        field = smooth_separable(noise.transpose(0, 1), sigmas).transpose(1, 0)
        mean = field.mean(dim=(1, 2, 3), keepdim=True)
        std = field.std(dim=(1, 2, 3), keepdim=True)
        fac = (torch.erfinv(2 * p - 1) * math.sqrt(2.0)).to(noise).reshape_as(mean)
        tau = fac * std + mean
        mask = (field > tau).to(field)
    return (mask, field, tau) if return_field else mask


def generate_cowmix_masks_like(example, mask_proportion_range, sigma_range):
    """cowmix.py:40-69."""
    with torch.no_grad():
        p, sigmas = draw_parameters(example.size(0), mask_proportion_range, sigma_range)
        shape = list(example.size())
        shape[1] = 1
        noise = torch.normal(mean=0, std=1, size=shape, dtype=example.dtype, device=example.device)
        return masks_from_noise(noise, p, sigmas)


def mix_with_mask(a, b, mask):
    """cowmix.py:72-73."""
    return a * mask + b * (1. - mask)


# ---------------------------------------------------------------- lovasz.py ----------------------
def jaccard_deltas(gt_sorted):
    """lovasz.py:19-31."""
    n = len(gt_sorted)
    total = gt_sorted.sum()
    inter = total - gt_sorted.float().cumsum(0)
    union = total + (1 - gt_sorted).float().cumsum(0)
    jac = 1. - inter / union
    if n > 1:
        jac[1:n] = jac[1:n] - jac[0:-1]
    return jac


def _running_mean(values, empty=0):
    """lovasz.py:235-253 (without the nan filter, which the path never enables)."""
    it = iter(values)
    try:
        acc = next(it)
    except StopIteration:
        return empty
    n = 1
    for n, v in enumerate(it, 2):
        acc += v
    return acc if n == 1 else acc / n


def _flatten(probas, labels, ignore):
    """lovasz.py:204-220."""
    if probas.dim() == 3:
        probas = probas.unsqueeze(1)
    c = probas.size(1)
    flat = probas.permute(0, 2, 3, 1).contiguous().view(-1, c)
    lab = labels.view(-1)
    if ignore is None:
        return flat, lab
    keep = lab != ignore
    return flat[keep.nonzero().squeeze()], lab[keep]


def _flat_loss(flat, lab, classes):
    """lovasz.py:173-201."""
    if flat.numel() == 0:
        return flat * 0.
    c = flat.size(1)
    todo = list(range(c)) if classes in ('all', 'present') else classes
    out = []
    for k in todo:
        fg = (lab == k).float()
        if classes == 'present' and fg.sum() == 0:
            continue
        if c == 1:
            if len(classes) > 1:
                raise ValueError('Sigmoid output possible only with 1 class')
            pred = flat[:, 0]
        else:
            pred = flat[:, k]
        err = (fg - pred).abs()
        err_sorted, perm = torch.sort(err, 0, descending=True)
        out.append(torch.dot(err_sorted, jaccard_deltas(fg[perm.data])))
    return _running_mean(out)


def lovasz_softmax(probas, labels, classes='present', per_image=False, ignore=None):
    """lovasz.py:155-170."""
    if per_image:
        return _running_mean(_flat_loss(*_flatten(pr.unsqueeze(0), lb.unsqueeze(0), ignore), classes=classes)
                             for pr, lb in zip(probas, labels))
    return _flat_loss(*_flatten(probas, labels, ignore), classes=classes)


def binary_lovasz_loss_with_logits(inp, target):
    """losses.py:239-250."""
    labels = torch.argmax(target, dim=1, keepdim=False)
    total = 0
    n_valid = 0
    for x, t in zip(torch.split(inp, 1, dim=0), torch.split(labels, 1, dim=0)):
        w = (t.sum() > 0).to(inp)
        n_valid += w
        total += lovasz_softmax(x, t, classes=[1], ignore=255, per_image=True) * w
    return total / (n_valid + 0.001)


def iou(preds, labels, C, EMPTY=1., ignore=None, per_image=False):
    """lovasz.py:54-73."""
    import numpy as np
    if not per_image:
        preds, labels = (preds,), (labels,)
    rows = []
    for pr, lb in zip(preds, labels):
        row = []
        for i in range(C):
            if i != ignore:
                inter = ((lb == i) & (pr == i)).sum()
                union = ((lb == i) | ((pr == i) & (lb != ignore))).sum()
                row.append(EMPTY if not union else float(inter) / float(union))
        rows.append(row)
    return 100 * np.array([_running_mean(col) for col in zip(*rows)])


# ---------------------------------------------------------------- mean_teacher.py / metrics.py ---
def update_ema_tensors(ema_tensors, tensors, alpha):
    """mean_teacher.py:10-11 on plain tensor lists."""
    with torch.no_grad():
        for e, p in zip(ema_tensors, tensors):
            e.mul_(alpha).add_(other=p, alpha=1. - alpha)


def update_ema_variables(model, ema_model, alpha):
    """mean_teacher.py:5-18 (buffers are re-pointed at the student's storage)."""
    with torch.no_grad():
        for e, p in zip(ema_model.parameters(), model.parameters()):
            e.data.mul_(alpha).add_(other=p.data, alpha=1. - alpha)
        for eb, b in zip(ema_model.buffers(), model.buffers()):
            eb.data = b.data


def dice_metric(x, y):
    """metrics.py:1-7."""
    inter = (x * y).sum(dim=(1, 2, 3))
    card = (x + y).sum(dim=(1, 2, 3))
    return (2. * inter + 1.) / (card + 1.)


def confusion_matrix(labels, preds, num_classes, ignore_index=None):
    """Restated confusion-matrix oracle: bincount(label*C + pred) over valid pixels."""
    lab, pr = labels.reshape(-1), preds.reshape(-1)
    if ignore_index is not None:
        keep = lab != ignore_index
        lab, pr = lab[keep], pr[keep]
    return torch.bincount(lab * num_classes + pr, minlength=num_classes * num_classes).view(num_classes, num_classes)


# ---------------------------------------------------------------- train.py:98-107 ----------------
def confidence_masked_consistency(student_logits, teacher_logits, confidence_threshold):
    """The inline consistency-loss code of train.py:98-107 as a function: returns
    (consistency_loss, confidence_modulator.mean())."""
    t = torch.sigmoid(teacher_logits)
    s = torch.sigmoid(student_logits)
    conf = (t.max(dim=1).values > confidence_threshold).to(t)
    sq = torch.pow(s - t, exponent=2.0)
    loss = (sq.sum(dim=1) * conf).sum() / conf.sum()
    return loss.mean(), conf.mean()


# ---------------------------------------------------------------- whole step (bench baseline) ----
def loss_path_step(image_a, image_b, teacher_a, teacher_b, scores, target, params, ema_params,
                   mode="binary", mask_proportion_range=(0.45, 0.55), sigma_range=(8, 32), alpha=0.99,
                   classes="present", per_image=False, ignore=255, num_classes=2):
    """The sequence bench.py times, in reference order (train.py:65-130): mask, two mixes, Lovasz
    forward+backward, EMA, confusion matrix of (labels, argmax scores)."""
    mask = generate_cowmix_masks_like(image_a, mask_proportion_range, sigma_range)
    with torch.no_grad():
        mixed_teacher = mix_with_mask(teacher_a, teacher_b, mask)
        mixed_images = mix_with_mask(image_a, image_b, mask)
    s = scores.detach().clone().requires_grad_(True)
    if mode == "binary":
        loss = binary_lovasz_loss_with_logits(s, target)
        labels = torch.argmax(target, dim=1)
    else:
        loss = lovasz_softmax(s, target, classes=classes, per_image=per_image, ignore=ignore)
        labels = target
    loss.backward()
    update_ema_tensors(ema_params, params, alpha)
    with torch.no_grad():
        cm = confusion_matrix(labels, torch.argmax(scores, dim=1), num_classes, ignore_index=ignore)
    return {"mask": mask, "mixed_images": mixed_images, "mixed_teacher": mixed_teacher,
            "loss": loss.detach(), "grad": s.grad, "cm": cm}
