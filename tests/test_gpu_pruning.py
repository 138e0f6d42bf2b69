"""Edge cases of the exact zero-delta tail pruning (csrc/lovasz.cu: lovasz_emin_kernel, the 'nothing here' words
of the key-build, the PRUNE0 pass) against the CPU oracle, which sorts every key.  Pruning must never change a bit:
gradients are compared on every non-zero element, pruned pixels must carry an exactly zero gradient, losses to 1e-5.

  * ties AT e_min (background errors equal to the smallest foreground error must stay and keep pixel order)
  * a class whose foreground pixel is predicted perfectly (e_min = 0: nothing pruned) next to classes where almost
    everything is pruned, in one call (the per-segment choice of the pass-0 ranking loop)
  * near-uniform predictions (almost every background key pruned; tiles of pass 0 mostly empty)
  * classes absent from the labels in 'present' / 'all' / list mode, void pixels, per_image, all label dtypes,
    more than 254 classes (no one-byte label copy), ragged shapes (scalar paths)
"""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu
REL = 1e-5


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def same_nonzero_bits(a, b):
    if not np.array_equal(a, b):
        return False
    nz = b != 0
    return bool(np.array_equal(bits(a)[nz], bits(b)[nz]))


@pytest.fixture(scope="module")
def ssl():
    import b200ssl
    return b200ssl


def run_both(ssl, probas, labels, **kw):
    dev = torch.device("cuda:0")
    x = probas.to(dev).requires_grad_(True)
    loss = ssl.lovasz.lovasz_softmax(x, labels.to(dev), **kw)
    loss.backward()
    o_loss, o_grad, _ = oracle.lovasz_softmax(probas.numpy(), labels.numpy(), **kw)
    assert abs(float(loss) - float(o_loss)) <= REL * abs(float(o_loss)) + 1e-12, (float(loss), float(o_loss))
    assert same_nonzero_bits(x.grad.cpu().numpy(), o_grad)
    return x.grad.cpu().numpy(), o_grad


def test_ties_at_emin_keep_pixel_order(ssl):
    """quantised probabilities: many background errors EQUAL the smallest foreground error of their class"""
    gen = torch.Generator().manual_seed(3)
    n, c, h, w = 2, 4, 96, 100
    pr = torch.softmax(torch.randn(n, c, h, w, generator=gen), 1)
    pr = (torch.round(pr * 8) / 8).contiguous()                     # errors on a grid of 1/8
    labels = torch.randint(0, c, (n, h, w), generator=gen)
    for per_image in (False, True):
        got, ref = run_both(ssl, pr, labels, classes="present", per_image=per_image)
        assert (ref == 0).mean() > 0.05                               # a visible share of the keys is prunable


def test_mixed_segments_pruned_and_unpruned_in_one_call(ssl):
    gen = torch.Generator().manual_seed(4)
    n, c, h, w = 2, 5, 128, 128
    logits = torch.randn(n, c, h, w, generator=gen) * 0.2            # near-uniform: almost everything prunable
    labels = torch.randint(0, c, (n, h, w), generator=gen)
    pr = torch.softmax(logits, 1)
    pr[0, 2, 5, 7] = 1.0                                             # class 2: one perfectly predicted fg pixel -> e_min = 0
    labels[0, 5, 7] = 2
    pr[1, 3, 0, 0] = 1.0 - 2.0 ** -21                                # class 3: e_min just below the 2^-20 switch
    labels[1, 0, 0] = 3
    got, ref = run_both(ssl, pr.contiguous(), labels, classes="all", per_image=False)
    # (classes 2 and 3 keep every key; most of their deltas still round to exactly 0 in fp32 -- lovasz_grad's staircase --
    #  so only the bit-exact comparison above says anything about them)
    assert (ref[:, 0] == 0).mean() > 0.5


@pytest.mark.parametrize("classes", ["present", "all", [1, 3], [0, 2, 4, 6]])
@pytest.mark.parametrize("dtype", [torch.int64, torch.int32, torch.uint8])
def test_absent_classes_void_pixels_and_label_types(ssl, classes, dtype):
    gen = torch.Generator().manual_seed(11)
    n, c, h, w = 3, 7, 61, 83                                        # ragged: scalar load / store paths
    pr = torch.softmax(torch.randn(n, c, h, w, generator=gen) * 1.5, 1).contiguous()
    labels = torch.randint(0, 4, (n, h, w), generator=gen)          # classes 4..6 never occur
    labels[torch.rand(n, h, w, generator=gen) < 0.1] = 255
    labels[2] = 255                                                  # an image of void pixels only
    for per_image in (False, True):
        run_both(ssl, pr, labels.to(dtype), classes=classes, per_image=per_image, ignore=255)


def test_more_classes_than_the_one_byte_label_copy_can_hold(ssl):
    gen = torch.Generator().manual_seed(12)
    n, c, h, w = 1, 300, 32, 36
    pr = torch.softmax(torch.randn(n, c, h, w, generator=gen), 1).contiguous()
    labels = torch.randint(0, c, (n, h, w), generator=gen)
    run_both(ssl, pr, labels, classes="present", per_image=False)


@pytest.mark.parametrize("scale", [0.05, 0.5, 4.0])
def test_pruned_share_from_near_uniform_to_confident(ssl, scale):
    """the same labels with predictions from near-uniform (almost everything pruned) to confident (little pruned)"""
    gen = torch.Generator().manual_seed(13)
    n, c, h, w = 2, 19, 160, 192
    lab = torch.nn.functional.avg_pool2d(torch.randn(n, c, h, w, generator=gen), 17, 1, 8).argmax(1)
    pr = torch.softmax(torch.randn(n, c, h, w, generator=gen) * scale, 1).contiguous()
    run_both(ssl, pr, lab, classes="present", per_image=False)
    # the fused forward+backward step entry takes the same kernels with the scale folded in
    step = ssl.LossPathStep(num_classes=c, mode="softmax", classes="present", per_image=False, ignore=255)
    loss, grad, _ = step.lovasz_loss_and_grad(pr.to("cuda:0"), lab.to("cuda:0"))
    o_loss, o_grad, _ = oracle.lovasz_softmax(pr.numpy(), lab.numpy(), classes="present", ignore=255)
    assert abs(float(loss) - float(o_loss)) <= REL * abs(float(o_loss))
    assert same_nonzero_bits(grad.cpu().numpy(), o_grad)
