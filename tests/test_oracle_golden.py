"""The CPU oracle (oracle/oracle.c through ctypes, and oracle/torch_port.py) against golden vectors
produced by the reference's own functions (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

import oracle
from oracle import torch_port
from conftest import load_golden, unpack_mask


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def same_floats(a, b):
    """bit-equal, except that any NaN matches any NaN (payloads are platform specific)."""
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    nan = np.isnan(a) & np.isnan(b)
    return bool(np.all((bits(a) == bits(b)) | nan))


MASK_MARGIN = 2e-7      # a differing pixel must sit within this distance of tau


def check_mask(mask, ref_mask, field, tau):
    diff = mask != ref_mask
    n_diff = int(diff.sum())
    assert n_diff <= max(1, int(1e-5 * mask.size)), f"{n_diff} mask pixels differ"
    if n_diff:
        t = np.broadcast_to(np.asarray(tau).reshape(-1, 1, 1, 1), field.shape)
        assert np.all(np.abs(field[diff] - t[diff]) <= MASK_MARGIN)


# ------------------------------------------------------------------------------ CowMix
@pytest.mark.parametrize("name", ["cowmix_small", "cowmix_c1"])
def test_cowmix_host_quantities(name):
    g = load_golden(name)
    sig = torch.from_numpy(g["sigmas"])
    assert oracle.kernel_size(sig.max().item()) == int(g["size"])
    taps = oracle.gaussian_taps(int(g["size"]), sig)
    assert np.array_equal(bits(taps), bits(g["taps"]))
    assert np.array_equal(bits(oracle.threshold_factors(g["p"])), bits(g["factors"]))


def test_kernel_size_known_answers():
    # SURVEY 8c known answers; python round is half-to-even
    for sigma, k in [(4, 25), (7.9, 49), (8, 49), (16, 97), (31.99, 193), (32, 193), (2.5, 17)]:
        assert oracle.kernel_size(sigma) == k
    assert oracle.kernel_size(8.5 / 3) == 17    # round(8.5) == 8


def test_gaussian_taps_off_centre():
    taps = oracle.gaussian_taps(7, torch.tensor([1.5]))[0]
    assert int(np.argmax(taps)) == 4
    assert abs(float(taps[4]) - 0.2787) < 1e-4 and abs(float(taps[0]) - 0.00796) < 1e-5


@pytest.mark.parametrize("name", ["cowmix_small", "cowmix_c1"])
def test_cowmix_field_and_mask(name):
    g = load_golden(name)
    ref_mask = unpack_mask(g)
    o = oracle.cowmix_masks_from_noise(g["noise"], g["p"], g["sigmas"])
    # the C layer accumulates with fma in tap order, mkldnn in its own order: fields agree to ~1e-7
    assert np.max(np.abs(o["field"] - g["field"])) <= MASK_MARGIN
    assert np.allclose(o["tau"], g["tau"], rtol=0, atol=2e-8)
    assert np.allclose(o["std"], g["std"], rtol=2e-6, atol=0)
    check_mask(o["mask"], ref_mask, g["field"], g["tau"])
    frac = o["mask"].reshape(o["mask"].shape[0], -1).mean(1)
    assert np.all(np.abs(frac - (1 - g["p"])) < 0.12)      # fraction of ones ~ 1 - p


@pytest.mark.parametrize("name", ["cowmix_small", "cowmix_c1"])
def test_torch_port_cowmix(name):
    g = load_golden(name)
    mask, field, tau = torch_port.masks_from_noise(torch.from_numpy(g["noise"]), torch.from_numpy(g["p"]),
                                                   torch.from_numpy(g["sigmas"]), return_field=True)
    check_mask(mask.numpy(), unpack_mask(g), g["field"], g["tau"])
    # whole function incl. the RNG stream (p, sigma, noise order)
    torch.manual_seed(int(g["seed"]))
    n, _, h, w = (int(x) for x in g["shape"])
    m2 = torch_port.generate_cowmix_masks_like(torch.zeros(n, 3, h, w), tuple(g["p_range"].tolist()),
                                               tuple(g["sigma_range"].tolist()))
    check_mask(m2.numpy(), unpack_mask(g), g["field"], g["tau"])


def test_mix():
    g = load_golden("mix")
    assert same_floats(oracle.mix(g["a"], g["b"], g["mask"]), g["out"])
    assert same_floats(oracle.mix(g["a"], g["b"], g["soft"]), g["out_soft"])
    assert same_floats(oracle.mix(g["a2"], g["b2"], g["mask"]), g["out_special"])
    # the unselected operand poisons the result (inf * 0 = NaN): the arithmetic form, not a select
    assert np.isnan(oracle.mix(g["a2"], g["b2"], g["mask"])).sum() == np.isnan(g["out_special"]).sum() > 0
    a, b, m = (torch.from_numpy(g[k]) for k in ("a", "b", "mask"))
    assert same_floats(torch_port.mix_with_mask(a, b, m).numpy(), g["out"])


# ------------------------------------------------------------------------------ Lovasz
def test_lovasz_grad_known_answers():
    g = load_golden("lovasz")
    for i in range(int(g["n_kats"])):
        assert np.array_equal(bits(oracle.lovasz_grad(g[f"kat{i}_in"])), bits(g[f"kat{i}_out"])), i
    assert np.array_equal(bits(oracle.lovasz_grad(g["kat_rand_in"])), bits(g["kat_rand_out"]))
    out = oracle.lovasz_grad([1, 0, 1])
    assert out.tolist() == [0.5, 0.16666662693023682, 0.3333333730697632]


def _cases(g):
    for spec in g["cases"]:
        name, classes, per_image, ignore, lab_key = str(spec).split("|")
        classes = classes if classes in ("all", "present") else eval(classes)
        yield name, dict(classes=classes, per_image=bool(int(per_image)),
                         ignore=None if ignore == "None" else int(ignore)), lab_key


def test_lovasz_softmax_cases():
    g = load_golden("lovasz")
    n_checked = 0
    for name, kw, lab_key in _cases(g):
        loss, grad, _ = oracle.lovasz_softmax(g["probas"], g[lab_key], **kw)
        ref_loss, ref_grad = float(g[f"{name}_loss"]), g[f"{name}_grad"]
        assert abs(float(loss) - ref_loss) <= 1e-5 * abs(ref_loss), name
        # no ties in these inputs -> the IEEE sequence reproduces autograd's gradient bit for bit
        assert np.array_equal(bits(grad), bits(ref_grad)), name
        n_checked += 1
    assert n_checked == 8
    loss, grad, _ = oracle.lovasz_softmax(g["sigmoid_probas"], g["sigmoid_labels"], classes=[1])
    assert abs(float(loss) - float(g["sigmoid_loss"])) <= 1e-5 * abs(float(g["sigmoid_loss"]))
    assert np.array_equal(bits(grad.reshape(g["sigmoid_grad"].shape)), bits(g["sigmoid_grad"]))


def test_torch_port_lovasz():
    g = load_golden("lovasz")
    for name, kw, lab_key in _cases(g):
        pr = torch.from_numpy(g["probas"]).requires_grad_(True)
        loss = torch_port.lovasz_softmax(pr, torch.from_numpy(g[lab_key]), **kw)
        loss.backward()
        assert float(loss) == float(g[f"{name}_loss"]), name
        assert np.array_equal(bits(pr.grad.numpy()), bits(g[f"{name}_grad"])), name


def test_binary_lovasz_shim():
    g = load_golden("binary_lovasz")
    loss, grad = oracle.binary_lovasz_loss_with_logits(g["logits"], g["target"])
    assert abs(float(loss) - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    assert np.array_equal(grad, g["grad"])       # equal values; the sign of a zero gradient is free
    nz = g["grad"] != 0
    assert np.array_equal(bits(grad)[nz], bits(g["grad"])[nz])
    assert np.all(grad[:, 0] == 0) and np.all(grad[2] == 0)     # only channel 1; empty image has weight 0
    x = torch.from_numpy(g["logits"]).requires_grad_(True)
    tl = torch_port.binary_lovasz_loss_with_logits(x, torch.from_numpy(g["target"]))
    tl.backward()
    assert float(tl) == float(g["loss"]) and np.array_equal(bits(x.grad.numpy()), bits(g["grad"]))


def test_lovasz_tie_order_is_stable_by_index():
    # equal errors: ascending pixel index decides (the reference leaves tie order unspecified)
    pred = np.array([0.5, 0.5, 0.5, 0.5], np.float32)
    fg = np.array([1, 0, 1, 0])
    _, grad = oracle.lovasz_segment(pred, fg)
    d = oracle.lovasz_grad([1, 0, 1, 0])
    assert np.array_equal(np.abs(grad), d)


# ------------------------------------------------------------------------------ EMA / metrics
def test_ema_bit_exact():
    g = load_golden("ema")
    n = int(g["n"])
    for tag, alpha in [("a099", 0.99), ("a0999", 0.999), ("a05", 0.5)]:
        ema = [g[f"ema0_{i}"].copy() for i in range(n)]
        par = [g[f"param{i}"] for i in range(n)]
        tp = [torch.from_numpy(x.copy()) for x in ema]
        for _ in range(3):
            oracle.ema_update(ema, par, alpha)
            torch_port.update_ema_tensors(tp, [torch.from_numpy(x) for x in par], alpha)
        for i in range(n):
            assert np.array_equal(bits(ema[i]), bits(g[f"{tag}_ema{i}"])), (tag, i)
            assert np.array_equal(bits(tp[i].numpy()), bits(g[f"{tag}_ema{i}"])), (tag, i)
    assert np.float32(0.99) == np.float32(0.9900000095367432)
    assert np.float32(1.0 - 0.99) == np.float32(0.009999999776482582)


def test_confusion_matrix_and_iou():
    g = load_golden("metrics")
    c = int(g["C"])
    cm, dropped = oracle.confusion_matrix(g["labels"], g["preds"], c)
    assert dropped == 0 and np.array_equal(cm, g["cm_plain"]) and cm.sum() == g["labels"].size
    cm_i, _ = oracle.confusion_matrix(g["labels_ign"], g["preds"], c, ignore_index=255)
    assert np.array_equal(cm_i, g["cm_ign"]) and cm_i.sum() == int((g["labels_ign"] != 255).sum())
    assert np.array_equal(torch_port.confusion_matrix(torch.from_numpy(g["labels_ign"]), torch.from_numpy(g["preds"]), c, 255).numpy(), g["cm_ign"])
    # lovasz.iou from the matrix (exact: same integer ratios)
    assert np.array_equal(oracle.iou_from_cm(cm, c), g["iou_plain"])
    cm_o, _ = oracle.confusion_matrix(g["labels_ign"], g["preds"], c, ignore_index=255, other_bucket=True)
    assert np.array_equal(oracle.iou_from_cm(cm_o, c, ignore=255), g["iou_ign"])
    cm_v, _ = oracle.confusion_matrix(g["labels_ign"], g["preds_void"], c, ignore_index=255, other_bucket=True)
    assert np.array_equal(oracle.iou_from_cm(cm_v, c, ignore=255), g["iou_void_pred"])
    per, _ = oracle.confusion_matrix(g["labels_ign"], g["preds"], c, ignore_index=255, other_bucket=True, per_image=True)
    per_iou = np.stack([oracle.iou_from_cm(m, c, ignore=255) for m in per])
    assert np.allclose(per_iou.mean(0), g["iou_ign_per_image"], rtol=1e-12)
    no_bucket, dropped = oracle.confusion_matrix(g["labels_ign"], g["preds_void"], c, ignore_index=255)
    assert dropped == int(((g["preds_void"] == 255) & (g["labels_ign"] != 255)).sum())
    assert no_bucket.sum() + dropped == int((g["labels_ign"] != 255).sum())


def test_dice():
    g = load_golden("metrics")
    assert np.array_equal(bits(oracle.dice_metric(g["dice_x"], g["dice_y"])), bits(g["dice"]))
    assert np.allclose(oracle.dice_metric(g["dice_soft_x"], g["dice_soft_y"]), g["dice_soft"], rtol=1e-6)
    # Dice from a per-image 2x2 confusion matrix equals metrics.dice_metric on {0,1} maps
    lab, pr = g["dice_y"][:, 0].astype(np.int64), g["dice_x"][:, 0].astype(np.int64)
    cm, _ = oracle.confusion_matrix(lab, pr, 2, per_image=True)
    tp, fp, fn = cm[:, 1, 1], cm[:, 0, 1], cm[:, 1, 0]
    d = (np.float32(2) * tp.astype(np.float32) + np.float32(1)) / ((2 * tp + fp + fn).astype(np.float32) + np.float32(1))
    assert np.array_equal(bits(d), bits(g["dice"]))
    assert np.array_equal(bits(torch_port.dice_metric(torch.from_numpy(g["dice_x"]), torch.from_numpy(g["dice_y"])).numpy()), bits(g["dice"]))


# ------------------------------------------------------------------------------ consistency (N1)
def test_consistency_oracle_vs_golden():
    g = load_golden("consistency")
    for tag, thr in [("t097", 0.97), ("t06", 0.6)]:
        loss, conf, grad = oracle.consistency_loss(g["student"], g["teacher"], thr)
        assert abs(float(loss) - float(g[f"{tag}_loss"])) <= 1e-5 * abs(float(g[f"{tag}_loss"]))
        assert float(conf) == float(g[f"{tag}_conf"])
        ref = g[f"{tag}_grad"]
        assert np.linalg.norm(grad - ref) <= 1e-5 * np.linalg.norm(ref)
        assert np.array_equal(grad == 0, ref == 0)                 # same confident pixels


# ------------------------------------------------------------------------------------------------
# Row N4: clip_grad_norm_ -> SGD.step -> EMA (train.py:122-130), golden = torch.optim.SGD + the
# reference's update_ema_variables executed in the build container
# ------------------------------------------------------------------------------------------------
SGD_VARIANTS = {"default": dict(lr=0.0001 / 4 * 9, momentum=0.9, weight_decay=0.0005, clip=5.0, alpha=0.99),
                "nesterov": dict(lr=0.05, momentum=0.8, weight_decay=0.0, nesterov=True, clip=None, alpha=0.999),
                "plain": dict(lr=0.1, momentum=0.0, weight_decay=0.01, clip=0.5, alpha=0.5),
                "damp": dict(lr=0.01, momentum=0.9, dampening=0.25, weight_decay=0.001, clip=1e9, alpha=0.99)}


@pytest.mark.parametrize("name", list(SGD_VARIANTS))
def test_sgd_ema_oracle_matches_torch_bit_for_bit(name):
    g = load_golden("sgd")
    v = dict(SGD_VARIANTS[name])
    clip, alpha = v.pop("clip"), v.pop("alpha")
    n = int(g["n"])
    params = [g[f"param0_{i}"].copy() for i in range(n)]
    ema = [g[f"ema0_{i}"].copy() for i in range(n)]
    moms = [np.zeros_like(p) for p in params] if v.get("momentum", 0.0) else None
    for step in range(3):
        grads = [g[f"grad{step}_{i}"] for i in range(n)]
        coef = None
        if clip is not None:
            ref_norm = np.float32(g[f"{name}_norm{step}"])
            ours = oracle.grad_total_norm(grads)
            assert abs(float(ours) - float(ref_norm)) <= 1e-6 * float(ref_norm)
            # the coefficient formula is pinned on torch's own norm; the fp64-accumulated norm may differ
            # from torch's fp32 norm-of-norms in the last bit
            coef = oracle.clip_coef(ref_norm, clip)
        oracle.sgd_ema_step(params, grads, moms, ema, first_step=(step == 0), coef=coef, ema_alpha=alpha, **v)
    for i in range(n):
        assert np.array_equal(params[i].view(np.uint32), g[f"{name}_s2_param{i}"].view(np.uint32)), (name, i)
        assert np.array_equal(ema[i].view(np.uint32), g[f"{name}_s2_ema{i}"].view(np.uint32)), (name, i)
        if moms is not None:
            assert np.array_equal(moms[i].view(np.uint32), g[f"{name}_s2_mom{i}"].view(np.uint32)), (name, i)


# ------------------------------------------------------------------------------------------------
# Row N2: bilinear up-sampling (train.py:72-75) + mix (train.py:82), golden = F.interpolate + the
# reference's mix_with_mask
# ------------------------------------------------------------------------------------------------
UPSAMPLE_TAGS = ["x4", "x1", "ragged", "x2"]
UPSAMPLE_BIT_EXACT = {"x4", "x1"}      # goldens produced by ATen's multi-threaded loop (see make_golden.py)


def unpack_bits(bits, shape):
    n = int(np.prod(shape))
    return np.unpackbits(bits)[:n].reshape(shape).astype(np.float32)


@pytest.mark.parametrize("tag", UPSAMPLE_TAGS)
def test_upsample_oracle_matches_aten_bit_for_bit(tag):
    g = load_golden("upsample")
    a, b = g[f"{tag}_a"], g[f"{tag}_b"]
    H, W = (int(v) for v in g[f"{tag}_size"])
    ua = oracle.upsample_bilinear(a, (H, W))
    mask = unpack_bits(g[f"{tag}_mask_bits"], (a.shape[0], 1, H, W))
    mixed = oracle.mix(ua, oracle.upsample_bilinear(b, (H, W)), mask)
    if tag in UPSAMPLE_BIT_EXACT:
        assert np.array_equal(ua.view(np.uint32), g[f"{tag}_up_a"].view(np.uint32))
        assert np.array_equal(mixed.view(np.uint32), g[f"{tag}_mixed"].view(np.uint32))
    else:
        # small planes stay below ATen's parallel grain and run its single-thread loop, which associates
        # the same weights differently (<= 4 ulp); the bar there is north_star's 1e-5
        tol = 1e-5 * float(np.abs(a).max())
        assert np.allclose(ua, g[f"{tag}_up_a"], rtol=1e-5, atol=tol)
        assert np.allclose(mixed, g[f"{tag}_mixed"], rtol=1e-5, atol=tol)
    if tag == "x4":
        # ATen's single-thread CPU loop associates differently (<= 4 ulp): inside north_star's 1e-5 bar
        assert np.allclose(ua, g["x4_up_a_1t"], rtol=1e-5, atol=1e-5 * float(np.abs(a).max()))
        assert not np.array_equal(ua, g["x4_up_a_1t"])


# ------------------------------------------------------------------------------------------------
# Row N3: lovasz_softmax(F.softmax(logits, 1), labels), golden = the reference's function + autograd
# ------------------------------------------------------------------------------------------------
SOFTMAX_LOVASZ_CASES = {"c5_present": ("present", False, None), "c21_all": ("all", False, 255),
                        "c3_perimg": ("present", True, 255), "c4_list": ([0, 2], False, None)}


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


@pytest.mark.parametrize("tag", list(SOFTMAX_LOVASZ_CASES))
def test_lovasz_from_logits_oracle_matches_reference(tag):
    g = load_golden("softmax_lovasz")
    classes, per_image, ignore = SOFTMAX_LOVASZ_CASES[tag]
    loss, grad = oracle.lovasz_softmax_with_logits(g[f"{tag}_logits"], g[f"{tag}_labels"], classes=classes,
                                                   per_image=per_image, ignore=ignore)
    ref = float(g[f"{tag}_loss"])
    assert abs(float(loss) - ref) <= 1e-5 * abs(ref)
    assert rel_l2(grad, g[f"{tag}_grad"]) <= 1e-5


# ------------------------------------------------------- row N2 (student side) and the validation path
def test_lowres_lovasz_oracle_vs_reference_golden():
    """oracle.binary_lovasz_loss_lowres (up-sampling + loss + transposed interpolation) against vectors made by
    the reference's losses.CalculateLoss + autograd (tests/golden/make_golden_lowres.py)."""
    g = load_golden("lowres_lovasz")
    for tag in ("s4", "s2", "ragged", "s8"):
        loss, g_low, _ = oracle.binary_lovasz_loss_lowres(g[f"{tag}_low"], g[f"{tag}_target"].astype(np.float32),
                                                           grad_out=0.7)
        ref = g[f"{tag}_grad"]
        assert abs(float(loss) * 0.7 - float(g[f"{tag}_loss"])) <= 1e-5 * abs(float(g[f"{tag}_loss"]))
        assert np.linalg.norm(g_low - ref) <= 1e-5 * np.linalg.norm(ref)
        assert not g_low[:, [c for c in range(g_low.shape[1]) if c != 1]].any()


def test_upsample_backward_oracle_vs_autograd():
    gen = torch.Generator().manual_seed(4)
    for h, w, H, W in [(8, 8, 32, 32), (7, 9, 28, 36), (5, 5, 13, 17), (16, 16, 16, 16), (3, 4, 27, 20)]:
        x = torch.randn(2, 3, h, w, generator=gen, requires_grad=True)
        up = torch.randn(2, 3, H, W, generator=gen)
        torch.nn.functional.interpolate(x, size=(H, W), mode="bilinear", align_corners=False).backward(up)
        got = oracle.upsample_bilinear_backward(up.numpy(), (h, w))
        assert np.linalg.norm(got - x.grad.numpy()) <= 1e-6 * np.linalg.norm(x.grad.numpy())


def test_validation_dice_oracle_vs_reference_golden():
    g = load_golden("validation")
    for tag in ("v4", "vr", "v1"):
        dice, cm = oracle.validation_dice(g[f"{tag}_pred"], g[f"{tag}_mask"])
        assert np.array_equal(dice.view(np.uint32), g[f"{tag}_dice"].view(np.uint32))
        assert int(cm.sum()) == g[f"{tag}_mask"].shape[0] * g[f"{tag}_mask"].shape[2] * g[f"{tag}_mask"].shape[3]


def test_lovasz_hinge_oracle_vs_reference_golden():
    """oracle.lovasz_hinge against lovasz.lovasz_hinge + autograd executed by the reference itself
    (tests/golden/make_golden.py hinge_case; lovasz.py:79-111)."""
    g = load_golden("lovasz_hinge")
    for name in g["cases"]:
        lg_key, lab_key, per_image, ignore = str(name).split("|")
        ignore = None if ignore == "None" else int(ignore)
        loss, grad = oracle.lovasz_hinge(g[lg_key], g[lab_key], per_image=bool(int(per_image)), ignore=ignore)
        want = float(g[str(name) + "|loss"])
        assert abs(float(loss) - want) <= 1e-5 * max(1.0, abs(want)), name
        assert np.array_equal(grad, g[str(name) + "|grad"]), name   # deltas are integer-count arithmetic: exact
