"""Parity of the CUDA path (through the C ABI, via the ctypes mirror modules) with the CPU oracle
and with the reference's golden vectors.  Bars: integer/byte/mask work bit-exact (masks: flips only
inside the documented margin around tau), Lovasz loss <= 1e-5 relative, Lovasz gradients bit-exact
against the stable-order oracle, EMA bit-exact."""
import numpy as np
import pytest
import torch

import oracle
from conftest import load_golden, unpack_mask

pytestmark = pytest.mark.gpu

REL = 1e-5              # north_star: fp32 results within 1e-5 relative
MASK_MARGIN = 2e-7


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def same_floats(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return bool(np.all((bits(a) == bits(b)) | (np.isnan(a) & np.isnan(b))))


def same_nonzero_bits(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    if not np.array_equal(a, b):
        return False
    nz = b != 0
    return bool(np.array_equal(bits(a)[nz], bits(b)[nz]))


@pytest.fixture(scope="module")
def ssl():
    import b200ssl
    return b200ssl


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def coherent_labels(gen, n, c, h, w, k=9):
    blob = torch.nn.functional.avg_pool2d(torch.randn(n, c, h, w, generator=gen), k, 1, k // 2)
    return blob.argmax(1)


def check_mask(mask, ref_mask, field, tau):
    diff = mask != ref_mask
    n_diff = int(diff.sum())
    assert n_diff <= max(1, int(1e-5 * mask.size)), f"{n_diff} mask pixels differ"
    if n_diff:
        t = np.broadcast_to(np.asarray(tau).reshape(-1, 1, 1, 1), field.shape)
        assert np.all(np.abs(field[diff] - t[diff]) <= MASK_MARGIN)
    return n_diff


# ================================================================================ CowMix mask
@pytest.mark.parametrize("name", ["cowmix_small", "cowmix_c1"])
def test_cowmix_golden(ssl, dev, name):
    g = load_golden(name)
    p, sig = torch.from_numpy(g["p"]), torch.from_numpy(g["sigmas"])
    mask, field = ssl.cowmix.masks_from_noise(torch.from_numpy(g["noise"]).to(dev), p, sig, return_field=True)
    mask, field = mask.cpu().numpy(), field.cpu().numpy()
    o = oracle.cowmix_masks_from_noise(g["noise"], g["p"], g["sigmas"])
    assert np.array_equal(bits(field), bits(o["field"])), "smoothed field is not bit-identical to the oracle"
    check_mask(mask, o["mask"], o["field"], o["tau"])
    check_mask(mask, unpack_mask(g), g["field"], g["tau"])          # the reference's own mask
    assert set(np.unique(mask)) <= {0.0, 1.0}


@pytest.mark.parametrize("n,h,w,sig", [
    (1, 7, 5, [1.0]),                  # kernel larger than the image
    (3, 33, 47, [0.7, 2.0, 3.3]),      # odd extents: scalar tails everywhere
    (2, 64, 130, [4.0, 9.5]),          # w/2 not a multiple of the block
    (1, 200, 16, [12.0]),              # tall and narrow, K=73
    (5, 16, 512, [1.5, 2.5, 3.5, 0.5, 6.0]),
])
def test_cowmix_ragged(ssl, dev, n, h, w, sig):
    gen = torch.Generator().manual_seed(h * 1000 + w)
    noise = torch.randn(n, 1, h, w, generator=gen)
    p = torch.rand(n, generator=gen) * 0.2 + 0.4
    sig = torch.tensor(sig)
    mask, field = ssl.cowmix.masks_from_noise(noise.to(dev), p, sig, return_field=True)
    o = oracle.cowmix_masks_from_noise(noise.numpy(), p, sig)
    assert np.array_equal(bits(field.cpu().numpy()), bits(o["field"]))
    check_mask(mask.cpu().numpy(), o["mask"], o["field"], o["tau"])
    # without the diagnostic field output the mask must be the same
    mask2 = ssl.cowmix.masks_from_noise(noise.to(dev), p, sig)
    assert torch.equal(mask, mask2)


def test_cowmix_config2_full_size(ssl, dev):
    """BASELINE configs[1]: 16x512x512, sigma in (8,32) -> K up to 193."""
    torch.manual_seed(0)
    n, h, w = 16, 512, 512
    p, sig = ssl.cowmix.draw_mask_parameters(n, (0.45, 0.55), (8, 32))
    noise = torch.randn(n, 1, h, w)
    mask, field = ssl.cowmix.masks_from_noise(noise.to(dev), p, sig, return_field=True)
    o = oracle.cowmix_masks_from_noise(noise.numpy(), p, sig)
    assert np.array_equal(bits(field.cpu().numpy()), bits(o["field"]))
    check_mask(mask.cpu().numpy(), o["mask"], o["field"], o["tau"])
    frac = mask.mean(dim=(1, 2, 3)).cpu()
    assert torch.all((frac - (1 - p)).abs() < 0.15)


def test_cowmix_generate_like_reference_rng_order(ssl, dev):
    """CPU generator: p then sigma (2N uniforms), then the device generator for the noise."""
    torch.manual_seed(5)
    ex = torch.zeros(4, 3, 64, 48, device=dev)
    m1 = ssl.cowmix.generate_cowmix_masks_like(ex, (0.4, 0.6), (2, 6))
    torch.manual_seed(5)
    p, sig = ssl.cowmix.draw_mask_parameters(4, (0.4, 0.6), (2, 6))
    noise = torch.normal(mean=0, std=1, size=[4, 1, 64, 48], dtype=torch.float32, device=dev)
    m2 = ssl.cowmix.masks_from_noise(noise, p, sig)
    assert torch.equal(m1, m2) and m1.shape == (4, 1, 64, 48) and m1.dtype == torch.float32
    o = oracle.cowmix_masks_from_noise(noise.cpu().numpy(), p, sig)
    check_mask(m1.cpu().numpy(), o["mask"], o["field"], o["tau"])


# ================================================================================ mix
def test_mix_golden(ssl, dev):
    g = load_golden("mix")
    t = {k: torch.from_numpy(v).to(dev) for k, v in g.items() if v.dtype == np.float32}
    assert same_floats(ssl.cowmix.mix_with_mask(t["a"], t["b"], t["mask"]).cpu().numpy(), g["out"])
    assert same_floats(ssl.cowmix.mix_with_mask(t["a"], t["b"], t["soft"]).cpu().numpy(), g["out_soft"])
    assert same_floats(ssl.cowmix.mix_with_mask(t["a2"], t["b2"], t["mask"]).cpu().numpy(), g["out_special"])


@pytest.mark.parametrize("n,c0,c1,h,w", [(2, 3, 2, 64, 64), (1, 3, 21, 17, 13), (3, 1, 0, 5, 7), (2, 3, 19, 32, 36), (4, 5, 7, 1, 1)])
def test_mix_fused_pairs(ssl, dev, n, c0, c1, h, w):
    gen = torch.Generator().manual_seed(c0 * 100 + c1)
    a0, b0 = torch.randn(n, c0, h, w, generator=gen), torch.randn(n, c0, h, w, generator=gen)
    mask = (torch.rand(n, 1, h, w, generator=gen) > 0.5).float()
    if c1:
        a1, b1 = torch.randn(n, c1, h, w, generator=gen), torch.randn(n, c1, h, w, generator=gen)
        o0, o1 = ssl.cowmix.mix2_with_mask(a0.to(dev), b0.to(dev), a1.to(dev), b1.to(dev), mask.to(dev))
        assert same_floats(o1.cpu().numpy(), oracle.mix(a1.numpy(), b1.numpy(), mask.numpy()))
    else:
        o0, _ = ssl.cowmix.mix2_with_mask(a0.to(dev), b0.to(dev), None, None, mask.to(dev))
    assert same_floats(o0.cpu().numpy(), oracle.mix(a0.numpy(), b0.numpy(), mask.numpy()))


def test_mix_unaligned_views_and_channel_mask(ssl, dev):
    gen = torch.Generator().manual_seed(9)
    big = torch.randn(3 * 2 * 3 * 10 * 12 + 8, generator=gen).to(dev)
    a = big[1:1 + 2 * 3 * 120].view(2, 3, 10, 12)      # 4-byte aligned only -> scalar path
    b = big[1 + 720 + 1:1 + 720 + 1 + 720].view(2, 3, 10, 12)
    mask = (torch.rand(2, 1, 10, 12, generator=gen) > 0.3).float().to(dev)
    out = ssl.cowmix.mix_with_mask(a, b, mask)
    assert same_floats(out.cpu().numpy(), oracle.mix(a.cpu().numpy(), b.cpu().numpy(), mask.cpu().numpy()))
    cm = torch.rand(2, 3, 10, 12, generator=gen).to(dev)
    out = ssl.cowmix.mix_with_mask(a, b, cm)
    assert same_floats(out.cpu().numpy(), oracle.mix(a.cpu().numpy(), b.cpu().numpy(), cm.cpu().numpy()))


def test_mix_autograd_matches_arithmetic(ssl, dev):
    gen = torch.Generator().manual_seed(2)
    a = torch.randn(2, 3, 8, 8, generator=gen).to(dev).requires_grad_(True)
    b = torch.randn(2, 3, 8, 8, generator=gen).to(dev).requires_grad_(True)
    mask = (torch.rand(2, 1, 8, 8, generator=gen) > 0.5).float().to(dev)
    go = torch.randn(2, 3, 8, 8, generator=gen).to(dev)
    ssl.cowmix.mix_with_mask(a, b, mask).backward(go)
    assert torch.equal(a.grad, go * mask) and torch.equal(b.grad, go * (1 - mask))


# ================================================================================ Lovasz
def _cases(g):
    for spec in g["cases"]:
        name, classes, per_image, ignore, lab_key = str(spec).split("|")
        classes = classes if classes in ("all", "present") else eval(classes)
        yield name, dict(classes=classes, per_image=bool(int(per_image)),
                         ignore=None if ignore == "None" else int(ignore)), lab_key


def run_lovasz(ssl, dev, probas, labels, **kw):
    pr = torch.from_numpy(np.asarray(probas)).to(dev).requires_grad_(True)
    lab = labels if isinstance(labels, torch.Tensor) else torch.from_numpy(np.asarray(labels))
    loss = ssl.lovasz.lovasz_softmax(pr, lab.to(dev), **kw)
    loss.backward()
    return float(loss), pr.grad.cpu().numpy()


def test_lovasz_golden(ssl, dev):
    g = load_golden("lovasz")
    for name, kw, lab_key in _cases(g):
        loss, grad = run_lovasz(ssl, dev, g["probas"], g[lab_key], **kw)
        ref = float(g[f"{name}_loss"])
        assert abs(loss - ref) <= REL * abs(ref), (name, loss, ref)
        assert same_nonzero_bits(grad, g[f"{name}_grad"]), name
    loss, grad = run_lovasz(ssl, dev, g["sigmoid_probas"], g["sigmoid_labels"], classes=[1])
    assert abs(loss - float(g["sigmoid_loss"])) <= REL * abs(float(g["sigmoid_loss"]))
    assert same_nonzero_bits(grad, g["sigmoid_grad"])


def test_binary_lovasz_golden(ssl, dev):
    g = load_golden("binary_lovasz")
    x = torch.from_numpy(g["logits"]).to(dev).requires_grad_(True)
    loss = ssl.losses.binary_lovasz_loss_with_logits(x, torch.from_numpy(g["target"]).to(dev))
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) <= REL * abs(float(g["loss"]))
    assert same_nonzero_bits(x.grad.cpu().numpy(), g["grad"])


@pytest.mark.parametrize("n,c,h,w,kw,ldt", [
    (2, 2, 256, 256, dict(classes="present", per_image=False, ignore=None), torch.int64),   # multi-tile segments
    (2, 3, 67, 61, dict(classes="all", per_image=True, ignore=255), torch.int64),           # odd hw: scalar path
    (3, 5, 64, 96, dict(classes="present", per_image=True, ignore=255), torch.uint8),
    (2, 21, 64, 64, dict(classes="present", per_image=False, ignore=255), torch.int32),     # absent classes
    (1, 4, 128, 130, dict(classes=[3, 1], per_image=False, ignore=None), torch.int64),
    (4, 2, 96, 100, dict(classes=[1], per_image=True, ignore=255), torch.uint8),
    (1, 19, 96, 192, dict(classes="all", per_image=False, ignore=255), torch.int64),
])
def test_lovasz_vs_oracle(ssl, dev, n, c, h, w, kw, ldt):
    gen = torch.Generator().manual_seed(n * 7 + c)
    probas = torch.softmax(torch.randn(n, c, h, w, generator=gen) * 2, 1)
    labels = coherent_labels(gen, n, min(c, 6), h, w, 11)           # classes >= 6 absent when c > 6
    if kw["ignore"] is not None:
        labels[torch.rand(n, h, w, generator=gen) < 0.05] = 255
    loss, grad = run_lovasz(ssl, dev, probas.numpy(), labels.to(ldt), **kw)
    o_loss, o_grad, _ = oracle.lovasz_softmax(probas.numpy(), labels.numpy(), **kw)
    assert abs(loss - float(o_loss)) <= REL * abs(float(o_loss)), (loss, float(o_loss))
    assert same_nonzero_bits(grad, o_grad)
    rel = np.linalg.norm(grad - o_grad) / max(np.linalg.norm(o_grad), 1e-30)
    assert rel <= REL


def test_lovasz_heavy_ties_stable_order(ssl, dev):
    """Probabilities quantised to 1/16: huge tie groups.  The CUDA sort is stable by pixel index, as
    the oracle; the reference leaves tie order unspecified, so the loss (tie-invariant) is the
    comparable quantity there and the gradient is compared with the stable oracle."""
    gen = torch.Generator().manual_seed(77)
    n, c, h, w = 2, 3, 80, 72
    probas = torch.round(torch.softmax(torch.randn(n, c, h, w, generator=gen), 1) * 16) / 16
    labels = coherent_labels(gen, n, c, h, w)
    loss, grad = run_lovasz(ssl, dev, probas.numpy(), labels, classes="all", per_image=False)
    o_loss, o_grad, _ = oracle.lovasz_softmax(probas.numpy(), labels.numpy(), classes="all", per_image=False)
    assert abs(loss - float(o_loss)) <= REL * abs(float(o_loss))
    assert same_nonzero_bits(grad, o_grad)


def test_lovasz_raw_logits_unbounded_keys(ssl, dev):
    """losses.py:241 feeds raw logits: errors are arbitrary non-negative floats, incl. > 1, 0, denormals."""
    gen = torch.Generator().manual_seed(78)
    n, c, h, w = 2, 2, 72, 64
    x = torch.randn(n, c, h, w, generator=gen) * 30
    x[0, 1, 0, :8] = torch.tensor([0.0, 1.0, 1e-40, -1e-40, 3e38, -3e38, 1.0, 0.0])
    labels = coherent_labels(gen, n, c, h, w)
    labels[0, 0, :8] = torch.tensor([0, 1, 1, 0, 1, 0, 1, 0])
    loss, grad = run_lovasz(ssl, dev, x.numpy(), labels, classes=[1], per_image=True, ignore=255)
    o_loss, o_grad, _ = oracle.lovasz_softmax(x.numpy(), labels.numpy(), classes=[1], per_image=True, ignore=255)
    assert abs(loss - float(o_loss)) <= REL * abs(float(o_loss))
    assert same_nonzero_bits(grad, o_grad)


def test_lovasz_config2_binary_shim_full_size(ssl, dev):
    """BASELINE configs[1]: 16x2x512x512 logits through binary_lovasz_loss_with_logits."""
    gen = torch.Generator().manual_seed(1234)
    n, c, h, w = 16, 2, 512, 512
    logits = torch.randn(n, c, h, w, generator=gen) * 3
    lab = coherent_labels(gen, n, c, h, w, 33)
    lab[5] = 0
    target = torch.nn.functional.one_hot(lab, c).permute(0, 3, 1, 2).float().contiguous()
    x = logits.to(dev).requires_grad_(True)
    loss = ssl.losses.binary_lovasz_loss_with_logits(x, target.to(dev))
    loss.backward()
    o_loss, o_grad = oracle.binary_lovasz_loss_with_logits(logits.numpy(), target.numpy())
    assert abs(float(loss) - float(o_loss)) <= REL * abs(float(o_loss))
    assert same_nonzero_bits(x.grad.cpu().numpy(), o_grad)
    assert float(x.grad[5].abs().sum()) == 0.0 and float(x.grad[:, 0].abs().sum()) == 0.0


def test_lovasz_degenerate_inputs(ssl, dev):
    probas = torch.softmax(torch.randn(2, 3, 8, 8), 1).to(dev)
    labels = torch.full((2, 8, 8), 255, dtype=torch.int64, device=dev)
    # all pixels void: zero loss and zero gradient
    pr = probas.clone().requires_grad_(True)
    loss = ssl.lovasz.lovasz_softmax(pr, labels, classes="present", ignore=255)
    loss.backward()
    assert float(loss) == 0.0 and float(pr.grad.abs().sum()) == 0.0
    # 'present' with no class present -> mean([]) = 0 ; 'all' -> loss = max error per class
    labels0 = torch.zeros((2, 8, 8), dtype=torch.int64, device=dev)
    l_all = ssl.lovasz.lovasz_softmax(probas, labels0, classes="all")
    o_all, _, _ = oracle.lovasz_softmax(probas.cpu().numpy(), labels0.cpu().numpy(), classes="all")
    assert abs(float(l_all) - float(o_all)) <= REL * abs(float(o_all))
    with pytest.raises(ValueError):
        ssl.lovasz.lovasz_softmax(probas[:, :1], labels0, classes="present")
    # single pixel (lovasz.py:29 `if p > 1`)
    one = torch.tensor([[[[0.3]], [[0.7]]]], device=dev)
    l1 = ssl.lovasz.lovasz_softmax(one, torch.ones((1, 1, 1), dtype=torch.int64, device=dev), classes="all")
    o1, _, _ = oracle.lovasz_softmax(one.cpu().numpy(), np.ones((1, 1, 1), np.int64), classes="all")
    assert abs(float(l1) - float(o1)) <= 1e-7


def test_lovasz_is_invariant_to_pixel_permutation(ssl, dev):
    gen = torch.Generator().manual_seed(3)
    probas = torch.softmax(torch.randn(1, 3, 64, 64, generator=gen), 1)
    labels = coherent_labels(gen, 1, 3, 64, 64)
    perm = torch.randperm(64 * 64, generator=gen)
    l1, _ = run_lovasz(ssl, dev, probas.numpy(), labels, classes="all")
    p2 = probas.reshape(1, 3, -1)[:, :, perm].reshape(1, 3, 64, 64).contiguous()
    lab2 = labels.reshape(1, -1)[:, perm].reshape(1, 64, 64).contiguous()
    l2, _ = run_lovasz(ssl, dev, p2.numpy(), lab2, classes="all")
    assert abs(l1 - l2) <= REL * abs(l1)


# ================================================================================ EMA
def test_ema_golden(ssl, dev):
    g = load_golden("ema")
    n = int(g["n"])
    for tag, alpha in [("a099", 0.99), ("a0999", 0.999), ("a05", 0.5)]:
        ema = [torch.from_numpy(g[f"ema0_{i}"].copy()).to(dev) for i in range(n)]
        par = [torch.from_numpy(g[f"param{i}"]).to(dev) for i in range(n)]
        upd = ssl.mean_teacher.EmaUpdater()
        for _ in range(3):
            upd(ema, par, alpha)
        for i in range(n):
            assert np.array_equal(bits(ema[i].cpu().numpy()), bits(g[f"{tag}_ema{i}"])), (tag, i)


def test_ema_modules_alias_buffers_and_unaligned(ssl, dev):
    torch.manual_seed(4)

    def make():
        return torch.nn.Sequential(torch.nn.Conv2d(3, 7, 3), torch.nn.BatchNorm2d(7), torch.nn.Conv2d(7, 5, 1),
                                   torch.nn.BatchNorm2d(5), torch.nn.Linear(11, 3)).to(dev)
    student, teacher = make(), make()
    ssl.mean_teacher.detach_model_parameters(teacher)
    assert all(not p.requires_grad for p in teacher.parameters())
    e0 = [p.detach().cpu().numpy().copy() for p in teacher.parameters()]
    for step in range(4):
        with torch.no_grad():
            for p in student.parameters():
                p.add_(torch.randn_like(p) * 0.1)
        ssl.mean_teacher.update_ema_variables(student, teacher, 0.99)
        oracle.ema_update(e0, [p.detach().cpu().numpy() for p in student.parameters()], 0.99)
    for a, b in zip(e0, teacher.parameters()):
        assert np.array_equal(bits(a), bits(b.detach().cpu().numpy()))
    for eb, b in zip(teacher.buffers(), student.buffers()):
        assert eb.data_ptr() == b.data_ptr()                      # mean_teacher.py:13-18
    # views at odd element offsets: 4-byte aligned only
    big_e, big_p = torch.randn(10000, device=dev), torch.randn(10000, device=dev)
    es = [big_e[1:4100], big_e[4101:4104], big_e[5001:9999]]
    ps = [big_p[3:4102], big_p[4103:4106], big_p[5000:9998]]
    ref = [t.cpu().numpy().copy() for t in es]
    ssl.mean_teacher.EmaUpdater()(es, ps, 0.9)
    oracle.ema_update(ref, [t.cpu().numpy() for t in ps], 0.9)
    for a, b in zip(ref, es):
        assert np.array_equal(bits(a), bits(b.cpu().numpy()))


def test_ema_alpha_limits(ssl, dev):
    e, p = torch.randn(5000, device=dev), torch.randn(5000, device=dev)
    e1 = e.clone()
    ssl.mean_teacher.EmaUpdater()([e1], [p], 1.0)
    assert torch.equal(e1, e)                 # alpha = 1: identity
    ssl.mean_teacher.EmaUpdater()([e1], [p], 0.0)
    assert torch.equal(e1, p)                 # alpha = 0: copy


# ================================================================================ confusion matrix / metrics
def test_confusion_golden(ssl, dev):
    g = load_golden("metrics")
    c = int(g["C"])
    t = {k: torch.from_numpy(g[k]).to(dev) for k in ("labels", "preds", "labels_ign", "preds_void")}
    assert np.array_equal(ssl.metrics.confusion_matrix(t["labels"], t["preds"], c).cpu().numpy(), g["cm_plain"])
    assert np.array_equal(ssl.metrics.confusion_matrix(t["labels_ign"], t["preds"], c, ignore_index=255).cpu().numpy(), g["cm_ign"])
    # lovasz.iou / iou_binary (lovasz.py:34-73) derived from the matrix
    assert np.array_equal(ssl.lovasz.iou(t["preds"], t["labels"], c), g["iou_plain"])
    assert np.array_equal(ssl.lovasz.iou(t["preds"], t["labels_ign"], c, ignore=255), g["iou_ign"])
    assert np.allclose(ssl.lovasz.iou(t["preds"], t["labels_ign"], c, ignore=255, per_image=True), g["iou_ign_per_image"], rtol=1e-12)
    assert np.array_equal(ssl.lovasz.iou(t["preds_void"], t["labels_ign"], c, ignore=255), g["iou_void_pred"])
    pb, lb = (t["preds"] > 2).long(), (t["labels"] > 2).long()
    assert np.isclose(ssl.lovasz.iou_binary(pb, lb), float(g["iou_binary"]), rtol=1e-12)
    assert np.isclose(ssl.lovasz.iou_binary(pb, lb, per_image=False), float(g["iou_binary_batch"]), rtol=1e-12)
    # Dice
    d = ssl.metrics.dice_metric(torch.from_numpy(g["dice_x"]).to(dev), torch.from_numpy(g["dice_y"]).to(dev))
    assert np.array_equal(bits(d.cpu().numpy()), bits(g["dice"]))
    ds = ssl.metrics.dice_metric(torch.from_numpy(g["dice_soft_x"]).to(dev), torch.from_numpy(g["dice_soft_y"]).to(dev))
    assert np.allclose(ds.cpu().numpy(), g["dice_soft"], rtol=1e-6)
    cm2 = ssl.metrics.confusion_matrix(lb, pb, 2, per_image=True)
    assert np.array_equal(bits(ssl.metrics.dice_from_cm(cm2).cpu().numpy()), bits(g["dice"]))


@pytest.mark.parametrize("dtype", [torch.int64, torch.int32, torch.uint8])
@pytest.mark.parametrize("c,n,h,w,ignore", [(19, 2, 1024, 2048, 255), (2, 3, 255, 257, None), (21, 1, 300, 301, 255), (150, 1, 128, 128, None)])
def test_confusion_vs_oracle(ssl, dev, dtype, c, n, h, w, ignore):
    gen = torch.Generator().manual_seed(c)
    labels = coherent_labels(gen, n, min(c, 24), h // 4 + 1, w // 4 + 1, 5)
    labels = torch.nn.functional.interpolate(labels[:, None].float(), size=(h, w), mode="nearest")[:, 0].long()
    preds = labels.clone()
    flip = torch.rand(n, h, w, generator=gen) < 0.2
    preds[flip] = torch.randint(0, c, (int(flip.sum()),), generator=gen)
    if ignore is not None:
        labels[torch.rand(n, h, w, generator=gen) < 0.05] = ignore
    lab_d, pr_d = labels.to(dtype).to(dev), preds.to(dtype).to(dev)
    cm = ssl.metrics.confusion_matrix(lab_d, pr_d, c, ignore_index=ignore)
    o, _ = oracle.confusion_matrix(labels.numpy(), preds.numpy(), c, ignore_index=ignore)
    assert np.array_equal(cm.cpu().numpy(), o)
    assert int(cm.sum()) == (int((labels != ignore).sum()) if ignore is not None else labels.numel())
    per = ssl.metrics.confusion_matrix(lab_d, pr_d, c, ignore_index=ignore, per_image=True)
    assert np.array_equal(per.sum(0).cpu().numpy(), o)
    o_per, _ = oracle.confusion_matrix(labels.numpy(), preds.numpy(), c, ignore_index=ignore, per_image=True)
    assert np.array_equal(per.cpu().numpy(), o_per)
    # streaming accumulation into an existing matrix
    acc = torch.zeros(c, c, dtype=torch.int64, device=dev)
    for i in range(n):
        ssl.metrics.confusion_matrix(lab_d[i], pr_d[i], c, ignore_index=ignore, out=acc)
    assert np.array_equal(acc.cpu().numpy(), o)


def test_confusion_out_of_range_and_other_bucket(ssl, dev):
    gen = torch.Generator().manual_seed(8)
    labels = torch.randint(0, 7, (2, 50, 51), generator=gen)
    preds = torch.randint(-1, 9, (2, 50, 51), generator=gen)
    cm, dropped = ssl.metrics.confusion_matrix(labels.to(dev), preds.to(dev), 5, ignore_index=6, return_dropped=True)
    o, od = oracle.confusion_matrix(labels.numpy(), preds.numpy(), 5, ignore_index=6)
    assert np.array_equal(cm.cpu().numpy(), o) and int(dropped) == od
    cmb = ssl.metrics.confusion_matrix(labels.to(dev), preds.to(dev), 5, ignore_index=6, other_bucket=True)
    ob, _ = oracle.confusion_matrix(labels.numpy(), preds.numpy(), 5, ignore_index=6, other_bucket=True)
    assert np.array_equal(cmb.cpu().numpy(), ob)


@pytest.mark.parametrize("run", [1, 3, 16, 17, 100, 5000])
@pytest.mark.parametrize("n_pixels", [16 * 32 * 7, 100003, 15, 1 << 20])
def test_confusion_uint8_run_stitching(ssl, dev, run, n_pixels):
    """uint8 labels take the kernel that merges runs of equal (label, pred) pairs across the lanes of a warp:
    run lengths around the 16 pixels of a lane and the 512 of a warp step, ragged ends, void labels, labels and
    predictions outside [0, C) (dropped and counted, or the other bucket)."""
    rng = np.random.default_rng(run * 7 + n_pixels % 1000)
    c = 19

    def runs(values):
        out = np.empty(0, np.uint8)
        while out.size < n_pixels:
            k = max(1, int(n_pixels / max(run, 1) / 4) + 1)
            out = np.concatenate([out, np.repeat(rng.choice(values, k).astype(np.uint8), rng.integers(1, 2 * run + 1, k))])
        return out[:n_pixels]

    labels = runs(np.array(list(range(c)) + [255, 255, 30]))
    preds = np.where(rng.random(n_pixels) < 0.5, labels, runs(np.array(list(range(c)) + [200]))).astype(np.uint8)
    lab_d, pr_d = torch.from_numpy(labels).to(dev), torch.from_numpy(preds).to(dev)
    for kw in (dict(ignore_index=255), dict(), dict(ignore_index=255, other_bucket=True)):
        cm, dropped = ssl.metrics.confusion_matrix(lab_d, pr_d, c, return_dropped=True, **kw)
        o, od = oracle.confusion_matrix(labels.astype(np.int64), preds.astype(np.int64), c, **kw)
        assert np.array_equal(cm.cpu().numpy(), o) and int(dropped) == od
        # without the dropped counter (the kernel's branch-light instantiation when there is no other bucket)
        assert np.array_equal(ssl.metrics.confusion_matrix(lab_d, pr_d, c, **kw).cpu().numpy(), o)
    # the same pixels as int64 through the generic kernel
    cm64 = ssl.metrics.confusion_matrix(lab_d.long(), pr_d.long(), c, ignore_index=255)
    assert torch.equal(cm64, ssl.metrics.confusion_matrix(lab_d, pr_d, c, ignore_index=255))


@pytest.mark.parametrize("c,h,w", [(2, 256, 256), (19, 128, 260), (21, 65, 63)])
def test_confusion_from_logits(ssl, dev, c, h, w):
    gen = torch.Generator().manual_seed(c + h)
    n = 3
    logits = torch.randn(n, c, h, w, generator=gen)
    logits[0, :, 0, 0] = 1.0                     # exact tie: first maximum wins
    labels = coherent_labels(gen, n, c, h, w)
    labels[torch.rand(n, h, w, generator=gen) < 0.05] = 255
    for dt in (torch.int64, torch.uint8):
        cm = ssl.metrics.confusion_matrix_from_logits(logits.to(dev), labels.to(dt).to(dev), ignore_index=255)
        o, _ = oracle.confusion_matrix(labels.numpy(), oracle.argmax_channels(logits.numpy()), c, ignore_index=255)
        assert np.array_equal(cm.cpu().numpy(), o)
        assert torch.equal(torch.from_numpy(oracle.argmax_channels(logits.numpy())), logits.argmax(1))


def test_miou_sweep_streaming_property(ssl, dev):
    """configs[4] shape (19 classes, 1024x2048 masks) at a sampled count: streaming accumulation of
    per-chunk matrices equals the matrix of the concatenation, and sums to the number of valid pixels."""
    gen = torch.Generator(device=dev).manual_seed(0)
    total = torch.zeros(19, 19, dtype=torch.int64, device=dev)
    chunks = []
    for i in range(3):
        lab = torch.randint(0, 20, (4, 1024, 2048), generator=gen, device=dev)
        lab[lab == 19] = 255
        prd = torch.randint(0, 19, (4, 1024, 2048), generator=gen, device=dev)
        ssl.metrics.confusion_matrix(lab, prd, 19, ignore_index=255, out=total)
        chunks.append((lab, prd))
    lab = torch.cat([c[0] for c in chunks])
    prd = torch.cat([c[1] for c in chunks])
    whole = ssl.metrics.confusion_matrix(lab, prd, 19, ignore_index=255)
    assert torch.equal(whole, total) and int(total.sum()) == int((lab != 255).sum())
    ref = torch.bincount((lab[lab != 255] * 19 + prd[lab != 255]), minlength=361).view(19, 19)
    assert torch.equal(ref, total)


# ================================================================================ whole step
def test_loss_path_step_matches_oracle(ssl, dev):
    gen = torch.Generator().manual_seed(21)
    n, c, h, w = 2, 2, 128, 96
    ia, ib = torch.rand(n, 3, h, w, generator=gen), torch.rand(n, 3, h, w, generator=gen)
    ta, tb = torch.randn(n, c, h, w, generator=gen), torch.randn(n, c, h, w, generator=gen)
    logits = torch.randn(n, c, h, w, generator=gen) * 3
    lab = coherent_labels(gen, n, c, h, w)
    target = torch.nn.functional.one_hot(lab, c).permute(0, 3, 1, 2).float().contiguous()
    params = [torch.randn(s, generator=gen) for s in [(64, 3, 3, 3), (64,), (5000,), (1,)]]
    ema = [torch.randn(p.shape, generator=gen) for p in params]
    d = lambda t: t.to(dev)
    step = ssl.LossPathStep(num_classes=c, sigma_range=(2, 6))
    d_ema = [d(t) for t in ema]
    torch.manual_seed(99)
    out = step(d(ia), d(ib), d(ta), d(tb), d(logits), d(target), [d(t) for t in params], d_ema)
    mask = out["mask"].cpu().numpy()
    assert same_floats(out["mixed_images"].cpu().numpy(), oracle.mix(ia.numpy(), ib.numpy(), mask))
    assert same_floats(out["mixed_teacher"].cpu().numpy(), oracle.mix(ta.numpy(), tb.numpy(), mask))
    o_loss, o_grad = oracle.binary_lovasz_loss_with_logits(logits.numpy(), target.numpy())
    assert abs(float(out["loss"]) - float(o_loss)) <= REL * abs(float(o_loss))
    assert same_nonzero_bits(out["grad"].cpu().numpy(), o_grad)
    e_np = [t.numpy().copy() for t in ema]
    oracle.ema_update(e_np, [t.numpy() for t in params], 0.99)
    for a, b in zip(e_np, d_ema):
        assert np.array_equal(bits(a), bits(b.cpu().numpy()))
    o_cm, _ = oracle.confusion_matrix(lab.numpy(), logits.argmax(1).numpy(), c, ignore_index=255)
    assert np.array_equal(out["cm"].cpu().numpy(), o_cm)


def test_loss_path_step_softmax_mode(ssl, dev):
    """configs[2]-shaped (21 classes, probabilities, batch-level 'present') through the fused C entry."""
    gen = torch.Generator().manual_seed(33)
    n, c, h, w = 2, 21, 64, 80
    ia, ib = torch.rand(n, 3, h, w, generator=gen), torch.rand(n, 3, h, w, generator=gen)
    ta, tb = torch.randn(n, c, h, w, generator=gen), torch.randn(n, c, h, w, generator=gen)
    probas = torch.softmax(torch.randn(n, c, h, w, generator=gen) * 2, 1)
    lab = coherent_labels(gen, n, 7, h, w)
    lab[torch.rand(n, h, w, generator=gen) < 0.05] = 255
    params = [torch.randn(s, generator=gen) for s in [(33, 7), (9000,)]]
    ema = [torch.randn(p.shape, generator=gen) for p in params]
    d = lambda t: t.to(dev)
    step = ssl.LossPathStep(num_classes=c, sigma_range=(2, 5), mode="softmax", classes="present",
                            per_image=False, ignore=255)
    d_ema = [d(t) for t in ema]
    out = step(d(ia), d(ib), d(ta), d(tb), d(probas), d(lab), [d(t) for t in params], d_ema)
    mask = out["mask"].cpu().numpy()
    assert same_floats(out["mixed_images"].cpu().numpy(), oracle.mix(ia.numpy(), ib.numpy(), mask))
    assert same_floats(out["mixed_teacher"].cpu().numpy(), oracle.mix(ta.numpy(), tb.numpy(), mask))
    o_loss, o_grad, _ = oracle.lovasz_softmax(probas.numpy(), lab.numpy(), classes="present", ignore=255)
    assert abs(float(out["loss"]) - float(o_loss)) <= REL * abs(float(o_loss))
    assert same_nonzero_bits(out["grad"].cpu().numpy(), o_grad)
    o_cm, _ = oracle.confusion_matrix(lab.numpy(), probas.argmax(1).numpy(), c, ignore_index=255)
    assert np.array_equal(out["cm"].cpu().numpy(), o_cm)
    e_np = [t.numpy().copy() for t in ema]
    oracle.ema_update(e_np, [t.numpy() for t in params], 0.99)
    for a, b in zip(e_np, d_ema):
        assert np.array_equal(bits(a), bits(b.cpu().numpy()))
    # Lovasz-only entry and a second step reuse the cached scratch
    loss2, grad2, _ = step.lovasz_loss_and_grad(d(probas), d(lab))
    assert float(loss2) == float(out["loss"]) and torch.equal(grad2, out["grad"])


@pytest.mark.parametrize("n,c,h,w", [(3, 2, 33, 35), (2, 3, 64, 48), (1, 2, 8, 4), (2, 17, 32, 32)])
def test_loss_path_step_binary_fused_and_fallback_shapes(ssl, dev, n, c, h, w):
    """hw % 4 != 0 or C > 16 takes the unfused front end, the others the fused one: same results."""
    gen = torch.Generator().manual_seed(n * 100 + c * 10 + h)
    logits = torch.randn(n, c, h, w, generator=gen) * 3
    lab = coherent_labels(gen, n, c, h, w, 5)
    lab[0] = 0                                            # an image without foreground: weight 0
    target = torch.nn.functional.one_hot(lab, c).permute(0, 3, 1, 2).float().contiguous()
    target = target * 0.9 + 0.05 / c                      # soft one-hot (label smoothing) keeps the argmax
    step = ssl.LossPathStep(num_classes=c, mode="binary")
    out = step(None, None, None, None, logits.to(dev), target.to(dev), None, None)
    o_loss, o_grad = oracle.binary_lovasz_loss_with_logits(logits.numpy(), target.numpy())
    assert abs(float(out["loss"]) - float(o_loss)) <= REL * abs(float(o_loss)) + 1e-12
    assert same_nonzero_bits(out["grad"].cpu().numpy(), o_grad)
    assert np.array_equal(out["labels"].cpu().numpy(), lab.numpy().astype(np.uint8))
    o_cm, _ = oracle.confusion_matrix(lab.numpy(), logits.argmax(1).numpy(), c, ignore_index=255)
    assert np.array_equal(out["cm"].cpu().numpy(), o_cm)
    # separate labels for the matrix (not the Lovasz labels)
    other = coherent_labels(gen, n, c, h, w, 3)
    out2 = step(None, None, None, None, logits.to(dev), target.to(dev), None, None, cm_labels=other.to(dev))
    o_cm2, _ = oracle.confusion_matrix(other.numpy(), logits.argmax(1).numpy(), c, ignore_index=255)
    assert np.array_equal(out2["cm"].cpu().numpy(), o_cm2)
    # and the autograd shim gives the same numbers
    x = logits.to(dev).requires_grad_(True)
    l2 = ssl.losses.binary_lovasz_loss_with_logits(x, target.to(dev))
    l2.backward()
    assert float(l2) == float(out["loss"]) and same_nonzero_bits(x.grad.cpu().numpy(), out["grad"].cpu().numpy())


# ================================================================================ consistency loss (N1)
def _clear_of_threshold(teacher, thr, margin=1e-4):
    """move teacher logits whose sigmoid is within `margin` of the threshold (confidence flips there
    depend on the last bit of exp)"""
    t = torch.sigmoid(teacher)
    near = (t - thr).abs() < margin
    return torch.where(near, teacher + 0.05, teacher)


def test_consistency_golden(ssl, dev):
    g = load_golden("consistency")
    for tag, thr in [("t097", 0.97), ("t06", 0.6)]:
        x = torch.from_numpy(g["student"]).to(dev).requires_grad_(True)
        loss, conf = ssl.consistency.confidence_masked_consistency(x, torch.from_numpy(g["teacher"]).to(dev), thr)
        loss.backward()
        assert abs(float(loss) - float(g[f"{tag}_loss"])) <= REL * abs(float(g[f"{tag}_loss"]))
        assert float(conf) == float(g[f"{tag}_conf"])
        ref = g[f"{tag}_grad"]
        got = x.grad.cpu().numpy()
        assert np.linalg.norm(got - ref) <= REL * np.linalg.norm(ref)
        assert np.array_equal(got == 0, ref == 0)


@pytest.mark.parametrize("n,c,h,w,thr", [(2, 2, 64, 64, 0.97), (3, 19, 33, 47, 0.6), (1, 1, 5, 7, 0.5), (16, 2, 512, 512, 0.97)])
def test_consistency_vs_oracle(ssl, dev, n, c, h, w, thr):
    gen = torch.Generator().manual_seed(n + c + h)
    student = torch.randn(n, c, h, w, generator=gen) * 3
    teacher = _clear_of_threshold(torch.randn(n, c, h, w, generator=gen) * 3, thr)
    x = student.to(dev).requires_grad_(True)
    loss, conf = ssl.consistency.confidence_masked_consistency(x, teacher.to(dev), thr)
    (loss * 10.0).backward()                                   # consistency_loss_weight (train.py:112)
    o_loss, o_conf, o_grad = oracle.consistency_loss(student.numpy(), teacher.numpy(), thr, grad_out=10.0)
    assert abs(float(loss) - float(o_loss)) <= REL * abs(float(o_loss))
    assert abs(float(conf) - float(o_conf)) <= 1e-7
    got = x.grad.cpu().numpy()
    assert np.linalg.norm(got - o_grad) <= REL * np.linalg.norm(o_grad)
    # elementwise: s - t cancels where student and teacher agree, so the bound is absolute (1e-6 of the largest entry)
    assert np.allclose(got, o_grad, rtol=1e-4, atol=1e-6 * float(np.abs(o_grad).max()))


def test_consistency_low_threshold_negative_teacher(ssl, dev):
    """Every teacher logit far below -1 and a threshold below sigmoid(-1): forward and backward must take
    the same confidence decision (the gradient kernel maximises raw logits, the forward sigmoid values)."""
    gen = torch.Generator().manual_seed(77)
    n, c, h, w, thr = 2, 3, 40, 52, 0.1
    student = torch.randn(n, c, h, w, generator=gen) * 2
    teacher = -1.5 - torch.rand(n, c, h, w, generator=gen) * 4          # sigmoid in (0.004, 0.18)
    teacher = _clear_of_threshold(teacher, thr)
    x = student.to(dev).requires_grad_(True)
    loss, conf = ssl.consistency.confidence_masked_consistency(x, teacher.to(dev), thr)
    loss.backward()
    o_loss, o_conf, o_grad = oracle.consistency_loss(student.numpy(), teacher.numpy(), thr)
    assert 0.0 < float(o_conf) < 1.0                                    # both decisions occur
    assert abs(float(loss) - float(o_loss)) <= REL * abs(float(o_loss))
    assert abs(float(conf) - float(o_conf)) <= 1e-7
    got = x.grad.cpu().numpy()
    assert np.array_equal(got == 0, o_grad == 0)                        # masked pixels carry no gradient
    assert np.linalg.norm(got - o_grad) <= REL * np.linalg.norm(o_grad)


def test_consistency_no_confident_pixel_is_nan_like_the_reference(ssl, dev):
    student = torch.zeros(1, 2, 8, 8, device=dev)
    teacher = torch.zeros(1, 2, 8, 8, device=dev)              # sigmoid = 0.5 < 0.97 everywhere
    loss, conf = ssl.consistency.confidence_masked_consistency(student, teacher, 0.97)
    assert torch.isnan(loss) and float(conf) == 0.0


@pytest.mark.parametrize("n,c,H,W,th,tw,thr", [
    (2, 2, 64, 96, 16, 24, 0.97),        # stride 4, vector path
    (1, 19, 48, 64, 24, 32, 0.6),        # stride 2, 19 classes
    (2, 3, 40, 56, 40, 56, 0.6),         # teacher at full resolution (read as it is)
    (2, 2, 33, 47, 9, 12, 0.5),          # ragged ratio, W % 4 != 0 (scalar path)
    (1, 2, 8, 8, 1, 1, 0.5),             # one teacher pixel
    (4, 2, 256, 256, 64, 64, 0.97),
])
def test_consistency_mixed_equals_mix_then_loss(ssl, dev, n, c, H, W, th, tw, thr):
    """train.py:69-82 + 98-107 with the teacher formed on the fly (b200ssl_consistency_mixed_*): the same values as
    mix2_with_mask (low-resolution second pair) followed by confidence_masked_consistency -- confidence count
    exact, gradients bit for bit, loss to the last bits of the fp64 partial sums -- and the oracle's
    upsample -> mix -> loss chain within the float tolerance."""
    gen = torch.Generator().manual_seed(n * 100 + c + H + th)
    student = torch.randn(n, c, H, W, generator=gen) * 2
    ta, tb = torch.randn(n, c, th, tw, generator=gen) * 3, torch.randn(n, c, th, tw, generator=gen) * 3
    mask = (torch.nn.functional.avg_pool2d(torch.randn(n, 1, H, W, generator=gen), 5, 1, 2) > 0).float()
    img = torch.zeros(n, 1, H, W)
    d = lambda t: t.to(dev)
    # three calls: (interpolate +) mix, then the loss
    _, mixed = ssl.cowmix.mix2_with_mask(d(img), d(img), d(ta), d(tb), d(mask))
    x1 = d(student).requires_grad_(True)
    loss1, conf1 = ssl.consistency.confidence_masked_consistency(x1, mixed, thr)
    (loss1 * 10.0).backward()
    # fused
    x2 = d(student).requires_grad_(True)
    loss2, conf2 = ssl.consistency.confidence_masked_consistency_mixed(x2, d(ta), d(tb), d(mask), thr, fused=True)
    (loss2 * 10.0).backward()
    # the wrapper's own choice of route (by channel count) gives the same numbers
    x4 = d(student).requires_grad_(True)
    loss4, conf4 = ssl.consistency.confidence_masked_consistency_mixed(x4, d(ta), d(tb), d(mask), thr)
    (loss4 * 10.0).backward()
    assert float(conf4) == float(conf2) and (float(conf2) == 0 or torch.equal(x4.grad, x2.grad))
    assert float(conf1) == float(conf2)
    if float(conf1) > 0:
        assert abs(float(loss1) - float(loss2)) <= 1e-6 * abs(float(loss1))
        assert torch.equal(x1.grad, x2.grad)
    else:
        assert torch.isnan(loss1) and torch.isnan(loss2)
    # straight through the C ABI without the confidence bytes: the backward recomputes the decision
    from b200ssl import _lib
    import ctypes as C
    xs, a_d, b_d, m_d = d(student), d(ta), d(tb), d(mask)
    stats = torch.empty(3, device=dev)
    ws = torch.empty(max(int(_lib.lib.b200ssl_consistency_mixed_workspace_bytes(n, H, W)), 16), dtype=torch.uint8, device=dev)
    go = torch.tensor([10.0], device=dev)
    g3 = torch.empty_like(xs)
    _lib.check(_lib.lib.b200ssl_consistency_mixed_forward(xs.data_ptr(), a_d.data_ptr(), b_d.data_ptr(), m_d.data_ptr(),
                                                          n, c, H, W, th, tw, thr, stats.data_ptr(), None, ws.data_ptr(),
                                                          ws.numel(), _lib.stream_ptr(dev)), "consistency_mixed_forward")
    _lib.check(_lib.lib.b200ssl_consistency_mixed_backward(xs.data_ptr(), a_d.data_ptr(), b_d.data_ptr(), m_d.data_ptr(),
                                                           n, c, H, W, th, tw, thr, stats.data_ptr(), None, go.data_ptr(),
                                                           g3.data_ptr(), _lib.stream_ptr(dev)), "consistency_mixed_backward")
    if float(conf1) > 0:
        assert torch.equal(g3, x2.grad) and float(stats[2]) == float(conf2)
    # the oracle chain
    up_a = oracle.upsample_bilinear(ta.numpy(), (H, W)) if (th, tw) != (H, W) else ta.numpy()
    up_b = oracle.upsample_bilinear(tb.numpy(), (H, W)) if (th, tw) != (H, W) else tb.numpy()
    o_mixed = oracle.mix(up_a, up_b, mask.numpy())
    assert np.array_equal(o_mixed, mixed.cpu().numpy())
    near = int((np.abs(1.0 / (1.0 + np.exp(-o_mixed.astype(np.float64))).max(1) - thr) < 1e-5).sum())
    o_loss, o_conf, o_grad = oracle.consistency_loss(student.numpy(), o_mixed, thr, grad_out=10.0)
    assert abs(float(conf2) - float(o_conf)) * n * H * W <= near + 1e-3
    if near == 0 and float(o_conf) > 0:
        assert abs(float(loss2) - float(o_loss)) <= REL * abs(float(o_loss))
        got = x2.grad.cpu().numpy()
        assert np.linalg.norm(got - o_grad) <= REL * np.linalg.norm(o_grad)


def test_consistency_mixed_rejects_bad_shapes(ssl, dev):
    s = torch.zeros(2, 2, 16, 16, device=dev)
    t = torch.zeros(2, 2, 4, 4, device=dev)
    m = torch.zeros(2, 1, 16, 16, device=dev)
    f = ssl.consistency.confidence_masked_consistency_mixed
    with pytest.raises(ValueError):
        f(s, t, torch.zeros(2, 2, 4, 5, device=dev), m, 0.9)          # the two teachers differ
    with pytest.raises(ValueError):
        f(s, t, t, torch.zeros(2, 2, 16, 16, device=dev), 0.9)        # per-channel mask
    with pytest.raises(ValueError):
        f(s, torch.zeros(2, 2, 32, 32, device=dev), torch.zeros(2, 2, 32, 32, device=dev), m, 0.9)   # teacher larger
    with pytest.raises(ValueError):
        f(s, torch.zeros(2, 3, 4, 4, device=dev), torch.zeros(2, 3, 4, 4, device=dev), m, 0.9)       # channel mismatch


def test_loss_path_step_forked_equals_serial(ssl, dev):
    """The internal fork/join of the three chains must not change a single bit, and work queued on the
    caller's stream right after the call must see all results."""
    gen = torch.Generator().manual_seed(5)
    n, c, h, w = 4, 2, 128, 128
    d = lambda t: t.to(dev)
    ia, ib = d(torch.rand(n, 3, h, w, generator=gen)), d(torch.rand(n, 3, h, w, generator=gen))
    ta, tb = d(torch.randn(n, c, h, w, generator=gen)), d(torch.randn(n, c, h, w, generator=gen))
    logits = d(torch.randn(n, c, h, w, generator=gen) * 3)
    lab = coherent_labels(gen, n, c, h, w)
    target = d(torch.nn.functional.one_hot(lab, c).permute(0, 3, 1, 2).float().contiguous())
    params = [d(torch.randn(s, generator=gen)) for s in [(300, 7), (9000,), (3,)]]
    outs = []
    for serial in (True, False, False):
        ema = [p.clone() * 0.5 for p in params]
        torch.manual_seed(123)
        step = ssl.LossPathStep(num_classes=c, sigma_range=(2, 6), serial=serial)
        o = step(ia, ib, ta, tb, logits, target, params, ema)
        # consumers on the caller's stream, queued immediately
        summary = torch.stack([o["mask"].sum(), o["mixed_images"].sum(), o["mixed_teacher"].sum(),
                               o["grad"].abs().sum(), o["loss"], o["cm"].sum().float(), sum(e.sum() for e in ema)])
        outs.append((o, ema, summary))
    for o, ema, summary in outs[1:]:
        assert torch.equal(summary, outs[0][2])
        for k in ("mask", "mixed_images", "mixed_teacher", "grad", "cm", "labels"):
            assert torch.equal(o[k], outs[0][0][k]), k
        assert all(torch.equal(a, b) for a, b in zip(ema, outs[0][1]))


@pytest.mark.parametrize("c", [100, 126, 127, 128])
def test_confusion_uint8_many_classes(ssl, dev, c):
    """uint8 labels with class counts whose matrix needs one shared-memory copy per block (c = 100), the opt-in
    above 48 KB (c = 126, 127 with the other bucket = 128 x 128 bins = exactly the 64 KB budget) and the generic
    kernel beyond it (c = 128 with the other bucket)."""
    rng = np.random.default_rng(c)
    n_pixels = 300_000
    labels = np.repeat(rng.integers(0, c + 3, n_pixels // 10), rng.integers(1, 40, n_pixels // 10))[:n_pixels]
    assert labels.size == n_pixels
    labels = np.minimum(labels, 255).astype(np.uint8)
    preds = np.where(rng.random(n_pixels) < 0.6, labels, rng.integers(0, c + 2, n_pixels)).astype(np.uint8)
    lab_d, pr_d = torch.from_numpy(labels).to(dev), torch.from_numpy(preds).to(dev)
    for kw in (dict(), dict(ignore_index=c + 1), dict(other_bucket=True), dict(ignore_index=3, other_bucket=True)):
        cm = ssl.metrics.confusion_matrix(lab_d, pr_d, c, **kw)
        o, _ = oracle.confusion_matrix(labels.astype(np.int64), preds.astype(np.int64), c, **kw)
        assert np.array_equal(cm.cpu().numpy(), o), kw
