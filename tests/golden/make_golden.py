"""Generates tests/golden/*.npz by executing the REFERENCE's own functions (imported read-only from
/root/reference) on seeded CPU inputs.  Run in the build container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

The reference ships no tests or golden vectors; these files are the pin for oracle/ and for the
CUDA kernels.  torch 2.11.0+cu128 (CPU ops) produced the committed vectors.
"""
import os
import sys
import warnings

import numpy as np
import torch

warnings.filterwarnings("ignore")
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
import cowmix as ref_cowmix          # noqa: E402
import lovasz as ref_lovasz          # noqa: E402
import mean_teacher as ref_mt        # noqa: E402
import metrics as ref_metrics        # noqa: E402
import losses as ref_losses          # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(1)


def save(name, **arrays):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in arrays.items()})
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def coherent_labels(gen, n, c, h, w, k=9):
    blob = torch.nn.functional.avg_pool2d(torch.randn(n, c, h, w, generator=gen), k, 1, k // 2)
    return blob.argmax(1)


def cowmix_case(name, n, h, w, p_range, sigma_range, seed):
    """Replays generate_cowmix_masks_like (cowmix.py:40-69) step by step so that p, sigma, noise and
    the smoothed field can be stored next to the mask the reference returns for the same seed."""
    torch.manual_seed(seed)
    example = torch.zeros(n, 3, h, w)
    mask = ref_cowmix.generate_cowmix_masks_like(example, p_range, sigma_range)
    # replay the RNG stream: p, sigma, noise (cowmix.py:44-55)
    torch.manual_seed(seed)
    import math
    p = torch.distributions.Uniform(torch.tensor(p_range[0]), torch.tensor(p_range[1])).rsample(sample_shape=[n])
    sig = torch.exp(torch.distributions.Uniform(torch.tensor(math.log(float(sigma_range[0]))),
                                                torch.tensor(math.log(float(sigma_range[1])))).rsample([n]))
    noise = torch.normal(mean=0, std=1, size=[n, 1, h, w], dtype=torch.float32)
    field = ref_cowmix.dual_pass_gaussian_fileter2d(noise.transpose(0, 1), sig).transpose(1, 0).contiguous()
    size = int(round(sig.max().item() * 3) * 2) + 1
    taps = ref_cowmix.gaussian_kernel_2d_vertical(size, sig).reshape(n, size)
    mean = field.mean(dim=(1, 2, 3), keepdim=True)
    std = field.std(dim=(1, 2, 3), keepdim=True)
    fac = (torch.erfinv(2 * p - 1) * math.sqrt(2.0)).reshape_as(mean)
    tau = fac * std + mean
    replay = (field > tau).float()
    assert torch.equal(replay, mask), "RNG replay does not reproduce the reference mask"
    save(name, p=p.numpy(), sigmas=sig.numpy(), noise=noise.numpy(), taps=taps.numpy(), size=size,
         field=field.numpy(), tau=tau.reshape(-1).numpy(), mean=mean.reshape(-1).numpy(),
         std=std.reshape(-1).numpy(), factors=fac.reshape(-1).numpy(),
         mask_bits=np.packbits(mask.numpy().astype(np.uint8).reshape(-1)), shape=np.array([n, 1, h, w]),
         p_range=np.array(p_range), sigma_range=np.array(sigma_range), seed=seed)


def lovasz_cases():
    gen = torch.Generator().manual_seed(11)
    out = {}
    # known-answer vectors for lovasz_grad (lovasz.py:19-31)
    kats = [[1, 0, 1], [0, 0, 0, 0], [1], [0], [1, 1, 0, 0, 1, 0, 0, 1]]
    rnd = (torch.rand(4096, generator=gen) < 0.3).float()
    for i, v in enumerate(kats):
        out[f"kat{i}_in"] = np.array(v, np.float32)
        out[f"kat{i}_out"] = ref_lovasz.lovasz_grad(torch.tensor(v, dtype=torch.float32)).numpy()
    out["kat_rand_in"] = rnd.numpy()
    out["kat_rand_out"] = ref_lovasz.lovasz_grad(rnd).numpy()
    out["n_kats"] = len(kats)

    # multi-class problems: probabilities, 3 classes + ignore, tie-free by construction (randn)
    n, c, h, w = 2, 3, 24, 20
    logits = torch.randn(n, c, h, w, generator=gen) * 2
    probas = torch.softmax(logits, 1)
    labels = coherent_labels(gen, n, c, h, w, 7)
    labels_ign = labels.clone()
    labels_ign[torch.rand(n, h, w, generator=gen) < 0.1] = 255
    labels_absent = labels.clone()
    labels_absent[labels_absent == 2] = 0     # class 2 absent everywhere
    out["probas"] = probas.numpy()
    out["labels"] = labels.numpy()
    out["labels_ign"] = labels_ign.numpy()
    out["labels_absent"] = labels_absent.numpy()
    cases = [
        ("present_batch", dict(classes="present", per_image=False, ignore=None), "labels"),
        ("present_image", dict(classes="present", per_image=True, ignore=None), "labels"),
        ("all_batch_ign", dict(classes="all", per_image=False, ignore=255), "labels_ign"),
        ("present_image_ign", dict(classes="present", per_image=True, ignore=255), "labels_ign"),
        ("present_absent", dict(classes="present", per_image=False, ignore=None), "labels_absent"),
        ("all_absent", dict(classes="all", per_image=True, ignore=None), "labels_absent"),
        ("list_02", dict(classes=[0, 2], per_image=False, ignore=255), "labels_ign"),
        ("list_1_image", dict(classes=[1], per_image=True, ignore=255), "labels_ign"),
    ]
    names = []
    for name, kw, lab_key in cases:
        pr = probas.clone().requires_grad_(True)
        lab = {"labels": labels, "labels_ign": labels_ign, "labels_absent": labels_absent}[lab_key]
        loss = ref_lovasz.lovasz_softmax(pr, lab, **kw)
        loss.backward()
        out[f"{name}_loss"] = loss.detach().numpy()
        out[f"{name}_grad"] = pr.grad.numpy()
        names.append(f"{name}|{kw['classes']}|{int(kw['per_image'])}|{kw['ignore']}|{lab_key}")
    out["cases"] = np.array(names)

    # sigmoid mode: [B,H,W] probabilities, classes=[1]
    sig = torch.sigmoid(torch.randn(n, h, w, generator=gen))
    lab01 = (labels > 0).long()
    pr = sig.clone().requires_grad_(True)
    loss = ref_lovasz.lovasz_softmax(pr, lab01, classes=[1], per_image=False, ignore=None)
    loss.backward()
    out["sigmoid_probas"], out["sigmoid_labels"] = sig.numpy(), lab01.numpy()
    out["sigmoid_loss"], out["sigmoid_grad"] = loss.detach().numpy(), pr.grad.numpy()
    save("lovasz", **out)


def binary_lovasz_case():
    """losses.binary_lovasz_loss_with_logits (losses.py:239-250): raw logits, soft one-hot target."""
    gen = torch.Generator().manual_seed(23)
    n, c, h, w = 3, 2, 40, 36
    logits = torch.randn(n, c, h, w, generator=gen) * 3
    lab = coherent_labels(gen, n, c, h, w, 9)
    lab[2] = 0                                            # third image has no foreground -> weight 0
    target = torch.nn.functional.one_hot(lab, c).permute(0, 3, 1, 2).float().contiguous()
    x = logits.clone().requires_grad_(True)
    loss = ref_losses.binary_lovasz_loss_with_logits(x, target)
    loss.backward()
    save("binary_lovasz", logits=logits.numpy(), target=target.numpy(), loss=loss.detach().numpy(),
         grad=x.grad.numpy())


def ema_case():
    gen = torch.Generator().manual_seed(5)
    shapes = [(1,), (3,), (17, 5), (4, 3, 3, 3), (1025,), (4096,), (4097,), (2, 8191)]
    params = [torch.randn(s, generator=gen) for s in shapes]
    ema0 = [torch.randn(s, generator=gen) for s in shapes]

    class M(torch.nn.Module):
        def __init__(self, ts):
            super().__init__()
            self.ps = torch.nn.ParameterList([torch.nn.Parameter(t.clone()) for t in ts])
            self.register_buffer("stat", torch.arange(4.0))

    out = {"n": len(shapes)}
    for alpha_name, alpha in [("a099", 0.99), ("a0999", 0.999), ("a05", 0.5)]:
        student, teacher = M(params), M(ema0)
        ref_mt.detach_model_parameters(teacher)
        for step in range(3):
            ref_mt.update_ema_variables(student, teacher, alpha)
        for i, t in enumerate(teacher.parameters()):
            out[f"{alpha_name}_ema{i}"] = t.detach().numpy()
        assert teacher.stat.data_ptr() == student.stat.data_ptr()   # buffers are aliased
    for i, (p, e) in enumerate(zip(params, ema0)):
        out[f"param{i}"], out[f"ema0_{i}"] = p.numpy(), e.numpy()
    save("ema", **out)


def metrics_case():
    gen = torch.Generator().manual_seed(31)
    n, h, w, c = 3, 32, 28, 5
    labels = coherent_labels(gen, n, c, h, w, 7)
    preds = coherent_labels(gen, n, c, h, w, 7)
    labels_ign = labels.clone()
    labels_ign[torch.rand(n, h, w, generator=gen) < 0.07] = 255
    preds_void = preds.clone()
    preds_void[torch.rand(n, h, w, generator=gen) < 0.05] = 255   # void class among predictions
    out = dict(labels=labels.numpy(), preds=preds.numpy(), labels_ign=labels_ign.numpy(),
               preds_void=preds_void.numpy(), C=c)
    out["iou_plain"] = ref_lovasz.iou(preds, labels, c)
    out["iou_ign"] = ref_lovasz.iou(preds, labels_ign, c, ignore=255)
    out["iou_ign_per_image"] = ref_lovasz.iou(preds, labels_ign, c, ignore=255, per_image=True)
    out["iou_void_pred"] = ref_lovasz.iou(preds_void, labels_ign, c, ignore=255)
    out["iou_binary"] = ref_lovasz.iou_binary((preds > 2).long(), (labels > 2).long())
    out["iou_binary_batch"] = ref_lovasz.iou_binary((preds > 2).long(), (labels > 2).long(), per_image=False)
    # restated confusion-matrix oracle (no reference function exists): bincount
    out["cm_plain"] = torch.bincount((labels * c + preds).reshape(-1), minlength=c * c).view(c, c).numpy()
    keep = labels_ign != 255
    out["cm_ign"] = torch.bincount((labels_ign[keep] * c + preds[keep]), minlength=c * c).view(c, c).numpy()
    # Dice (metrics.py:1-7) on {0,1} maps shaped like train.py:171-175
    x = (preds > 2).float().unsqueeze(1)
    y = (labels > 2).float().unsqueeze(1)
    out["dice_x"], out["dice_y"] = x.numpy(), y.numpy()
    out["dice"] = ref_metrics.dice_metric(x, y).numpy()
    xs = torch.rand(n, 2, h, w, generator=gen)
    ys = torch.rand(n, 2, h, w, generator=gen)
    out["dice_soft_x"], out["dice_soft_y"] = xs.numpy(), ys.numpy()
    out["dice_soft"] = ref_metrics.dice_metric(xs, ys).numpy()
    save("metrics", **out)


def mix_case():
    gen = torch.Generator().manual_seed(41)
    n, c, h, w = 2, 3, 12, 10
    a = torch.randn(n, c, h, w, generator=gen)
    b = torch.randn(n, c, h, w, generator=gen)
    mask = (torch.rand(n, 1, h, w, generator=gen) > 0.5).float()
    soft = torch.rand(n, 1, h, w, generator=gen)
    a2, b2 = a.clone(), b.clone()
    a2[0, 0, 0, 0], b2[0, 0, 0, 1] = float("inf"), float("-inf")
    a2[0, 1, 0, 2], b2[0, 1, 0, 3] = float("nan"), -0.0
    save("mix", a=a.numpy(), b=b.numpy(), mask=mask.numpy(), soft=soft.numpy(), a2=a2.numpy(), b2=b2.numpy(),
         out=ref_cowmix.mix_with_mask(a, b, mask).numpy(),
         out_soft=ref_cowmix.mix_with_mask(a, b, soft).numpy(),
         out_special=ref_cowmix.mix_with_mask(a2, b2, mask).numpy())


def reference_statements(path, first, last):
    """Source lines first..last (1-based, inclusive) of a reference file, dedented: the reference's own statements,
    compiled from the read-only tree at generation time (nothing is copied into this repository)."""
    import textwrap
    with open(path) as f:
        lines = f.readlines()[first - 1:last]
    return compile(textwrap.dedent("".join(lines)), f"{path}:{first}-{last}", "exec")


def consistency_case():
    """train.py:97-108 is inline code inside train() and train.py does not import here (kornia is not installed).
    Round 2: the golden no longer comes from the restated port -- the reference's OWN statements (those source
    lines, compiled from /root/reference/train.py) are executed on seeded tensors, with `config` reduced to the one
    key they read.  oracle/torch_port.confidence_masked_consistency is then checked against these vectors like
    every other restatement (tests/test_oracle_golden.py)."""
    code = reference_statements("/root/reference/train.py", 97, 108)
    gen = torch.Generator().manual_seed(53)
    n, c, h, w = 2, 3, 20, 28
    student = torch.randn(n, c, h, w, generator=gen) * 3
    teacher = torch.randn(n, c, h, w, generator=gen) * 3
    out = dict(student=student.numpy(), teacher=teacher.numpy())
    for tag, thr in [("t097", 0.97), ("t06", 0.6)]:
        x = student.clone().requires_grad_(True)
        ns = {"torch": torch, "mixed_ema_pred": teacher, "mixed_student_pred": x,
              "config": {"train": {"confidence_threshold": thr}}}
        exec(code, ns)
        loss, conf = ns["consistency_loss"], ns["confidence_modulator"]
        loss.backward()
        out[f"{tag}_loss"], out[f"{tag}_conf"], out[f"{tag}_grad"] = loss.detach().numpy(), conf.numpy(), x.grad.numpy()
    save("consistency", **out)


def softmax_lovasz_case():
    """Row N3: lovasz_softmax(F.softmax(logits, 1), labels) with autograd, by the reference's own function."""
    gen = torch.Generator().manual_seed(123)
    out = {}
    cases = {"c5_present": (2, 5, 24, 32, "present", False, None), "c21_all": (2, 21, 20, 20, "all", False, 255),
             "c3_perimg": (3, 3, 16, 28, "present", True, 255), "c4_list": (2, 4, 18, 22, [0, 2], False, None)}
    for tag, (n, c, h, w, classes, per_image, ignore) in cases.items():
        logits = torch.randn(n, c, h, w, generator=gen) * 2.5
        labels = coherent_labels(gen, n, c, h, w)
        if ignore is not None:
            labels[torch.rand(n, h, w, generator=gen) < 0.1] = ignore
        x = logits.clone().requires_grad_(True)
        loss = ref_lovasz.lovasz_softmax(torch.softmax(x, 1), labels, classes=classes, per_image=per_image, ignore=ignore)
        loss.backward()
        out[f"{tag}_logits"], out[f"{tag}_labels"] = logits.numpy(), labels.numpy()
        out[f"{tag}_loss"], out[f"{tag}_grad"] = loss.detach().numpy(), x.grad.numpy()
    save("softmax_lovasz", **out)


def upsample_case():
    """Row N2: the interpolate calls of train.py:72-75 (teacher predictions to image size) followed by the
    reference's mix (train.py:82)."""
    # ATen's CPU bilinear kernel is not a single function of its input: with ONE intra-op thread and a
    # batched input it takes a differently associated loop (results differ by up to 4 ulp) [probed].
    # The goldens are made with the multi-threaded path, whose arithmetic is the documented formula
    # (and the CUDA kernel's); "_1t" stores the single-thread variant for the tolerance check.
    gen = torch.Generator().manual_seed(91)
    out = {}
    torch.set_num_threads(8)
    for tag, (n, c, h, w, H, W) in {"x4": (2, 3, 16, 24, 64, 96), "x1": (1, 2, 20, 28, 20, 28),
                                     "ragged": (2, 2, 13, 17, 50, 61), "x2": (1, 19, 8, 16, 16, 32)}.items():
        a = torch.randn(n, c, h, w, generator=gen) * 3
        b = torch.randn(n, c, h, w, generator=gen) * 3
        mask = (torch.rand(n, 1, H, W, generator=gen) > 0.5).float()
        ua = torch.nn.functional.interpolate(a, (H, W), mode='bilinear', align_corners=False)
        ub = torch.nn.functional.interpolate(b, (H, W), mode='bilinear', align_corners=False)
        out[f"{tag}_a"], out[f"{tag}_b"], out[f"{tag}_mask_bits"] = a.numpy(), b.numpy(), np.packbits(mask.numpy().astype(np.uint8))
        out[f"{tag}_size"] = np.array([H, W])
        out[f"{tag}_up_a"] = ua.numpy()
        out[f"{tag}_mixed"] = ref_cowmix.mix_with_mask(ua, ub, mask).numpy()
        if tag == "x4":
            torch.set_num_threads(1)
            out["x4_up_a_1t"] = torch.nn.functional.interpolate(a, (H, W), mode='bilinear', align_corners=False).numpy()
            torch.set_num_threads(8)
    torch.set_num_threads(1)
    save("upsample", **out)


def sgd_case():
    """Row N4: train.py:121-130 -- clip_grad_norm_ -> SGD.step() -> zero_grad() -> update_ema_variables,
    executed with torch's own optimizer / clip function and the reference's EMA function."""
    gen = torch.Generator().manual_seed(77)
    shapes = [(1,), (3,), (17, 5), (4, 3, 3, 3), (1025,), (4097,)]
    params0 = [torch.randn(s, generator=gen) for s in shapes]
    ema0 = [torch.randn(s, generator=gen) for s in shapes]
    grads = [[torch.randn(s, generator=gen) * (0.02 if step == 1 else 1.5) for s in shapes] for step in range(3)]

    class M(torch.nn.Module):
        def __init__(self, ts):
            super().__init__()
            self.ps = torch.nn.ParameterList([torch.nn.Parameter(t.clone()) for t in ts])

    out = {"n": len(shapes), "steps": 3}
    for i, (p, e) in enumerate(zip(params0, ema0)):
        out[f"param0_{i}"], out[f"ema0_{i}"] = p.numpy(), e.numpy()
    for step in range(3):
        for i, g in enumerate(grads[step]):
            out[f"grad{step}_{i}"] = g.numpy()
    base_lr = 0.0001 * 1 / 4 * 9                               # configs/default_config.py:139
    variants = {"default": dict(lr=base_lr, momentum=0.9, weight_decay=0.0005, clip=5.0, alpha=0.99),
                "nesterov": dict(lr=0.05, momentum=0.8, weight_decay=0.0, nesterov=True, clip=None, alpha=0.999),
                "plain": dict(lr=0.1, momentum=0.0, weight_decay=0.01, clip=0.5, alpha=0.5),
                "damp": dict(lr=0.01, momentum=0.9, dampening=0.25, weight_decay=0.001, clip=1e9, alpha=0.99)}
    for name, v in variants.items():
        v = dict(v)
        clip, alpha = v.pop("clip"), v.pop("alpha")
        student, teacher = M(params0), M(ema0)
        ref_mt.detach_model_parameters(teacher)
        opt = torch.optim.SGD(student.parameters(), **v)
        for step in range(3):
            for p, g in zip(student.parameters(), grads[step]):
                p.grad = g.clone()
            if clip is not None:
                tn = torch.nn.utils.clip_grad_norm_(student.parameters(), clip)    # train.py:122
                out[f"{name}_norm{step}"] = tn.numpy()
            opt.step()                                                                # train.py:123
            opt.zero_grad()                                                           # train.py:124
            ref_mt.update_ema_variables(student, teacher, alpha)                      # train.py:130
            if step != 2:
                continue                                   # only the state after the last step is stored
            for i, (p, e) in enumerate(zip(student.parameters(), teacher.parameters())):
                out[f"{name}_s{step}_param{i}"] = p.detach().numpy().copy()
                out[f"{name}_s{step}_ema{i}"] = e.detach().numpy().copy()
                if v.get("momentum", 0.0) != 0.0:
                    out[f"{name}_s{step}_mom{i}"] = opt.state[p]["momentum_buffer"].numpy().copy()
    save("sgd", **out)


def hinge_case():
    """lovasz.lovasz_hinge (lovasz.py:79-111) + autograd on seeded logits: per image / per batch, with and without a
    void label, an image with only void pixels, an image without foreground, confident logits (most errors <= 0)."""
    gen = torch.Generator().manual_seed(77)
    out = {}
    n, h, w = 3, 28, 36
    lab = (coherent_labels(gen, n, 2, h, w, 7) > 0).long()
    lab[2] = 0                                              # image 2: no foreground
    lab_ign = lab.clone()
    lab_ign[torch.rand(n, h, w, generator=gen) < 0.15] = 255
    lab_ign[1] = 255                                        # image 1: only void pixels
    noisy = torch.randn(n, h, w, generator=gen) * 2.0
    confident = (2.0 * lab.float() - 1.0) * 3.0 + torch.randn(n, h, w, generator=gen) * 1.5   # mostly error <= 0
    out["labels"], out["labels_ign"] = lab.numpy(), lab_ign.numpy()
    out["logits_noisy"], out["logits_confident"] = noisy.numpy(), confident.numpy()
    names = []
    for lg_key, lg in (("logits_noisy", noisy), ("logits_confident", confident)):
        for lab_key, lb, ignore in (("labels", lab, None), ("labels_ign", lab_ign, 255)):
            for per_image in (True, False):
                x = lg.clone().requires_grad_(True)
                loss = ref_lovasz.lovasz_hinge(x, lb, per_image=per_image, ignore=ignore)
                loss.backward()
                name = f"{lg_key}|{lab_key}|{int(per_image)}|{ignore}"
                out[name + "|loss"] = loss.detach().numpy()
                out[name + "|grad"] = x.grad.numpy()
                names.append(name)
    out["cases"] = np.array(names)
    save("lovasz_hinge", **out)


if __name__ == "__main__":
    if len(sys.argv) > 1:                       # regenerate selected files only: make_golden.py sgd_case ...
        for fn in sys.argv[1:]:
            globals()[fn]()
        sys.exit(0)
    cowmix_case("cowmix_small", 3, 40, 56, (0.4, 0.6), (1.0, 3.0), seed=3)
    cowmix_case("cowmix_c1", 2, 256, 256, (0.45, 0.55), (8, 32), seed=0)      # BASELINE configs[0]
    lovasz_cases()
    hinge_case()
    binary_lovasz_case()
    ema_case()
    metrics_case()
    mix_case()
    consistency_case()
