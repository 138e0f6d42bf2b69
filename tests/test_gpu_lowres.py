"""Row N2, student side and the validation path, on the GPU.

  * losses.CalculateLoss + binary_lovasz_loss_with_logits from LOW-RESOLUTION logits: the bilinear up-sampling
    (losses.py:18-19 / train.py:93-94) happens inside the Lovasz front end and the gradient comes back through the
    transposed interpolation.  Bars: bit-identical to the CPU oracle (same interpolation arithmetic, same stable
    sort, same gather order); loss <= 1e-5 relative and gradient <= 1e-5 in relative L2 against vectors produced
    by the REFERENCE's CalculateLoss + autograd (tests/golden/lowres_lovasz.npz), and against the unfused path
    (F.interpolate + the full-resolution loss) on the same GPU.
  * metrics.validation_dice = train.py:171-175 in one pass: Dice bit-exact against the reference
    (tests/golden/validation.npz), matrices exact against the numpy oracle.
"""
import numpy as np
import pytest
import torch

import oracle
from conftest import load_golden

pytestmark = pytest.mark.gpu
REL = 1e-5


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def ssl():
    import b200ssl
    return b200ssl


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


@pytest.mark.parametrize("tag", ["s4", "s2", "ragged", "s8"])
def test_lowres_lovasz_matches_reference_and_oracle(ssl, dev, tag):
    g = load_golden("lowres_lovasz")
    low = torch.from_numpy(g[f"{tag}_low"]).to(dev).requires_grad_(True)
    target = torch.from_numpy(g[f"{tag}_target"].astype(np.float32)).to(dev)
    crit = ssl.losses.CalculateLoss([{"loss_fn": ssl.losses.binary_lovasz_loss_with_logits, "weight": [0.7]}])
    loss = crit([low], target)
    loss.backward()
    got = low.grad.cpu().numpy()
    ref = g[f"{tag}_grad"]
    assert abs(float(loss) - float(g[f"{tag}_loss"])) <= REL * abs(float(g[f"{tag}_loss"]))
    assert np.linalg.norm(got - ref) <= REL * np.linalg.norm(ref)
    assert not got[:, [c for c in range(got.shape[1]) if c != 1]].any()     # the loss reads channel 1 only
    # bit for bit against the oracle (unit upstream gradient, then the weight as autograd applies it)
    x = torch.from_numpy(g[f"{tag}_low"]).to(dev).requires_grad_(True)
    l1 = ssl.losses.binary_lovasz_loss_with_logits(x, target)
    l1.backward()
    o_loss, o_low, _ = oracle.binary_lovasz_loss_lowres(g[f"{tag}_low"], g[f"{tag}_target"].astype(np.float32))
    assert abs(float(l1) - float(o_loss)) <= REL * abs(float(o_loss))
    assert np.array_equal(bits(x.grad.cpu().numpy()), bits(o_low)), "low-resolution gradient differs from the oracle"


@pytest.mark.parametrize("n,c,lh,lw,H,W", [(16, 2, 128, 128, 512, 512), (2, 3, 17, 29, 100, 116), (1, 2, 64, 64, 64, 128)])
def test_lowres_lovasz_equals_unfused_path(ssl, dev, n, c, lh, lw, H, W):
    gen = torch.Generator().manual_seed(lh + W)
    low = (torch.randn(n, c, lh, lw, generator=gen) * 3).to(dev)
    lab = torch.nn.functional.avg_pool2d(torch.randn(n, c, H, W, generator=gen), 9, 1, 4).argmax(1)
    target = torch.nn.functional.one_hot(lab, c).permute(0, 3, 1, 2).float().contiguous().to(dev)
    a = low.clone().requires_grad_(True)
    la = ssl.losses.binary_lovasz_loss_with_logits(a, target)
    (la * 2.5).backward()
    b = low.clone().requires_grad_(True)
    up = torch.nn.functional.interpolate(b, size=(H, W), mode="bilinear", align_corners=False)
    lb = ssl.losses.binary_lovasz_loss_with_logits(up, target)
    (lb * 2.5).backward()
    assert abs(float(la) - float(lb)) <= REL * abs(float(lb))
    ga, gb = a.grad.double(), b.grad.double()
    assert float((ga - gb).norm()) <= REL * float(gb.norm())
    # the transposed interpolation on its own: oracle bit for bit, autograd to rounding
    gfull = torch.randn(n, H, W, generator=gen)
    out = torch.empty(n, lh, lw, device=dev)
    from b200ssl import _lib
    _lib.check(_lib.lib.b200ssl_upsample_bilinear_backward(gfull.to(dev).data_ptr(), n, H, W, out.data_ptr(), lh, lw,
                                                           _lib.stream_ptr(dev)))
    assert np.array_equal(bits(out.cpu().numpy()), bits(oracle.upsample_bilinear_backward(gfull.numpy(), (lh, lw))))
    x = torch.zeros(n, 1, lh, lw, requires_grad=True)
    torch.nn.functional.interpolate(x, size=(H, W), mode="bilinear", align_corners=False).backward(gfull[:, None])
    assert float((out.cpu() - x.grad[:, 0]).norm()) <= REL * float(x.grad.norm())


def test_lowres_falls_back_when_the_shape_is_not_supported(ssl, dev):
    gen = torch.Generator().manual_seed(1)
    low = (torch.randn(1, 2, 9, 9, generator=gen)).to(dev).requires_grad_(True)
    lab = (torch.rand(1, 30, 30, generator=gen) > 0.5).long()                     # W % 4 != 0
    target = torch.nn.functional.one_hot(lab, 2).permute(0, 3, 1, 2).float().contiguous().to(dev)
    loss = ssl.losses.binary_lovasz_loss_with_logits(low, target)
    loss.backward()
    o_loss, o_low, _ = oracle.binary_lovasz_loss_lowres(low.detach().cpu().numpy(), target.cpu().numpy())
    assert abs(float(loss) - float(o_loss)) <= REL * abs(float(o_loss))
    assert np.linalg.norm(low.grad.cpu().numpy() - o_low) <= REL * np.linalg.norm(o_low)


@pytest.mark.parametrize("tag", ["v4", "vr", "v1"])
def test_validation_dice_golden(ssl, dev, tag):
    g = load_golden("validation")
    dice, cm = ssl.metrics.validation_dice(torch.from_numpy(g[f"{tag}_pred"]).to(dev), torch.from_numpy(g[f"{tag}_mask"]).to(dev))
    assert np.array_equal(bits(dice.cpu().numpy()), bits(g[f"{tag}_dice"]))
    assert float(dice.mean()) == float(g[f"{tag}_mean"])
    o_dice, o_cm = oracle.validation_dice(g[f"{tag}_pred"], g[f"{tag}_mask"])
    assert np.array_equal(cm.cpu().numpy(), o_cm)


def test_validation_dice_vs_oracle_and_torch(ssl, dev):
    gen = torch.Generator().manual_seed(8)
    for n, c, h, w, H, W in [(4, 2, 128, 256, 512, 1024), (2, 2, 37, 53, 111, 97), (1, 3, 8, 8, 64, 64)]:
        pred = torch.randn(n, c, h, w, generator=gen)
        pred[0, 0, 0, 0] = float("nan")                                               # argmax treats NaN as the maximum
        mask = torch.rand(n, max(c, 2), H, W, generator=gen)
        dice, cm = ssl.metrics.validation_dice(pred.to(dev), mask.to(dev))
        o_dice, o_cm = oracle.validation_dice(pred.numpy(), mask.numpy())
        assert np.array_equal(cm.cpu().numpy(), o_cm)
        assert np.array_equal(bits(dice.cpu().numpy()), bits(o_dice))
        assert int(cm.sum()) == n * H * W
        # the reference's op sequence on the same GPU (train.py:171-175)
        p = pred.to(dev)
        oh = torch.nn.functional.one_hot(torch.argmax(p, dim=1), num_classes=c).permute(0, 3, 1, 2).to(p)
        pb = torch.nn.functional.interpolate(oh, size=(H, W), mode="nearest")
        mb = (mask.to(dev) > 0.5).to(p)
        want = (2. * (pb[:, 1:2] * mb[:, 1:2]).sum(dim=(1, 2, 3)) + 1.) / ((pb[:, 1:2] + mb[:, 1:2]).sum(dim=(1, 2, 3)) + 1.)
        assert torch.equal(dice, want)
        # streaming accumulation
        _, cm2 = ssl.metrics.validation_dice(pred.to(dev), mask.to(dev), cm_out=cm.clone())
        assert torch.equal(cm2, 2 * cm)
