"""Row N3 on the GPU: lovasz_softmax straight from logits (soft-max statistics, probabilities formed in the
key-build, soft-max backward in place) against the reference's goldens and the oracle; plus the
multi-class key-build (classes in groups of 8) in probability mode against the per-class oracle.
Bars (north_star): loss <= 1e-5 relative, gradient <= 1e-5 in relative L2 norm."""
import numpy as np
import pytest
import torch

import oracle
from conftest import load_golden
from test_oracle_golden import SOFTMAX_LOVASZ_CASES, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ssl():
    import b200ssl
    return b200ssl


@pytest.mark.parametrize("materialize", [True, False])
@pytest.mark.parametrize("tag", list(SOFTMAX_LOVASZ_CASES))
def test_from_logits_matches_reference_golden(ssl, tag, materialize):
    dev = torch.device("cuda:0")
    g = load_golden("softmax_lovasz")
    classes, per_image, ignore = SOFTMAX_LOVASZ_CASES[tag]
    x = torch.from_numpy(g[f"{tag}_logits"]).to(dev).requires_grad_(True)
    labels = torch.from_numpy(g[f"{tag}_labels"]).to(dev)
    loss = ssl.lovasz.lovasz_softmax_with_logits(x, labels, classes=classes, per_image=per_image, ignore=ignore,
                                                 materialize=materialize)
    loss.backward()
    ref = float(g[f"{tag}_loss"])
    assert abs(float(loss) - ref) <= 1e-5 * abs(ref)
    assert rel_l2(x.grad.cpu().numpy(), g[f"{tag}_grad"]) <= 1e-5


@pytest.mark.parametrize("shape", [(2, 19, 64, 96), (1, 21, 33, 47), (3, 9, 40, 40), (2, 2, 31, 29)])
@pytest.mark.parametrize("label_dtype", [torch.int64, torch.uint8])
def test_from_logits_matches_unfused_path_and_oracle(ssl, shape, label_dtype):
    """fused == lovasz_softmax(torch.softmax(...)) on the same GPU within the bars; also vs the oracle."""
    dev = torch.device("cuda:0")
    n, c, h, w = shape
    gen = torch.Generator().manual_seed(sum(shape))
    logits = torch.randn(n, c, h, w, generator=gen) * 3
    blob = torch.nn.functional.avg_pool2d(torch.randn(n, c, h, w, generator=gen), 9, 1, 4)
    labels = blob.argmax(1)
    labels[torch.rand(n, h, w, generator=gen) < 0.05] = 255
    for per_image, materialize in ((False, True), (True, True), (False, False), (True, False)):
        x = logits.to(dev).requires_grad_(True)
        loss = ssl.lovasz.lovasz_softmax_with_logits(x, labels.to(dev).to(label_dtype), per_image=per_image, ignore=255,
                                                     materialize=materialize)
        loss.backward()
        y = logits.to(dev).requires_grad_(True)
        loss2 = ssl.lovasz.lovasz_softmax(torch.softmax(y, 1), labels.to(dev).to(label_dtype), per_image=per_image, ignore=255)
        loss2.backward()
        assert abs(float(loss) - float(loss2)) <= 1e-5 * abs(float(loss2))
        assert rel_l2(x.grad.cpu().numpy(), y.grad.cpu().numpy()) <= 1e-5
        o_loss, o_grad = oracle.lovasz_softmax_with_logits(logits.numpy(), labels.numpy(), per_image=per_image, ignore=255)
        assert abs(float(loss) - float(o_loss)) <= 1e-5 * abs(float(o_loss))
        assert rel_l2(x.grad.cpu().numpy(), o_grad) <= 1e-5


def test_softmax_stats_and_backward_kernels(ssl):
    """The two stand-alone kernels against torch's CUDA soft-max (statistics exact to 2 ulp, backward 1e-6)."""
    import ctypes as C
    from b200ssl import _lib
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(4)
    for (n, c, h, w) in [(2, 21, 32, 48), (1, 3, 17, 19), (1, 36, 24, 24), (2, 33, 9, 7)]:
        x = (torch.randn(n, c, h, w, generator=gen) * 4).to(dev)
        hw = h * w
        stats = torch.empty((2, n, hw), device=dev)
        st = _lib.stream_ptr(dev)
        _lib.check(_lib.lib.b200ssl_softmax_stats(x.data_ptr(), n, c, hw, stats[0].data_ptr(), stats[1].data_ptr(), st))
        assert torch.equal(stats[0].view(n, h, w), x.max(1).values)
        want_sum = torch.exp(x - x.max(1, keepdim=True).values).sum(1)
        assert torch.allclose(stats[1].view(n, h, w), want_sum, rtol=1e-6, atol=0)
        g = torch.randn(n, c, h, w, generator=gen).to(dev)
        p = torch.softmax(x, 1)
        want = (g - (g * p).sum(1, keepdim=True)) * p
        got = g.clone()
        _lib.check(_lib.lib.b200ssl_softmax_backward(x.data_ptr(), stats[0].data_ptr(), stats[1].data_ptr(),
                                                     got.data_ptr(), n, c, hw, st))
        assert rel_l2(got.cpu().numpy(), want.cpu().numpy()) <= 1e-6
        # round 2: the soft-max written out, and its backward from the stored probabilities
        probas = torch.empty_like(x)
        _lib.check(_lib.lib.b200ssl_softmax_forward(x.data_ptr(), n, c, hw, probas.data_ptr(), st))
        assert torch.allclose(probas, p, rtol=2e-6, atol=1e-30)
        got2 = g.clone()
        _lib.check(_lib.lib.b200ssl_softmax_backward_probas(probas.data_ptr(), got2.data_ptr(), n, c, hw, st))
        assert rel_l2(got2.cpu().numpy(), want.cpu().numpy()) <= 1e-6
