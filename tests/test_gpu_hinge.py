"""lovasz.lovasz_hinge (reference lovasz.py:79-111) on the CUDA path: the radix-sort core of lovasz_softmax with
the hinge error 1 - logit * (2*label - 1).  Pixels with error <= 0 are never sorted (relu = 0, they sit behind every
positive error), which must not change a bit against the oracle, which sorts every valid pixel as the reference does.

  * the reference's own outputs (tests/golden/lovasz_hinge.npz: per image / per batch, void label, an image with only
    void pixels, an image without foreground, confident logits)
  * seeded problems against the C oracle: all label dtypes, ragged shapes, ties, every error <= 0, no valid pixel
  * a 2x1024x2048 problem (2^22 keys per image, look-back over 512 tiles per segment)
"""
import ctypes as C

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


def run(ssl, dev, logits, labels, dtype=torch.int64, **kw):
    x = torch.from_numpy(np.ascontiguousarray(logits)).to(dev).requires_grad_(True)
    lab = torch.from_numpy(np.ascontiguousarray(labels)).to(dtype).to(dev)
    loss = ssl.lovasz.lovasz_hinge(x, lab, **kw)
    loss.backward()
    return float(loss.detach()), x.grad.cpu().numpy()


def check(ssl, dev, logits, labels, dtype=torch.int64, **kw):
    loss, grad = run(ssl, dev, logits, labels, dtype, **kw)
    o_loss, o_grad = oracle.lovasz_hinge(logits, labels, **kw)
    assert abs(loss - float(o_loss)) <= REL * max(1.0, abs(float(o_loss)))
    assert np.array_equal(bits(grad + 0.0), bits(o_grad + 0.0))  # +0.0 folds -0.0 (a zero delta times -1) into +0.0
    return loss, grad


def test_hinge_reference_golden(ssl, dev):
    g = load_golden("lovasz_hinge")
    for name in g["cases"]:
        lg_key, lab_key, per_image, ignore = str(name).split("|")
        ignore = None if ignore == "None" else int(ignore)
        loss, grad = run(ssl, dev, g[lg_key], g[lab_key], per_image=bool(int(per_image)), ignore=ignore)
        want = float(g[str(name) + "|loss"])
        assert abs(loss - want) <= REL * max(1.0, abs(want)), name
        assert np.array_equal(grad, g[str(name) + "|grad"]), name


@pytest.mark.parametrize("dtype", [torch.int64, torch.int32, torch.uint8])
@pytest.mark.parametrize("b,h,w", [(2, 64, 96), (3, 33, 47), (1, 300, 301), (4, 128, 128)])
@pytest.mark.parametrize("per_image", [True, False])
def test_hinge_vs_oracle(ssl, dev, dtype, b, h, w, per_image):
    rng = np.random.default_rng(b * 1000 + h + int(per_image))
    labels = (rng.random((b, h // 8 + 1, w // 8 + 1)) < 0.4).repeat(8, 1).repeat(8, 2)[:, :h, :w].astype(np.int64)
    logits = ((2.0 * labels - 1.0) * 1.5 + rng.standard_normal((b, h, w)) * 2.0).astype(np.float32)
    check(ssl, dev, logits, labels, dtype, per_image=per_image, ignore=None)
    labels[rng.random((b, h, w)) < 0.1] = 255
    check(ssl, dev, logits, labels, dtype, per_image=per_image, ignore=255)


def test_hinge_edge_cases(ssl, dev):
    rng = np.random.default_rng(5)
    b, h, w = 3, 40, 52
    labels = (rng.random((b, h, w)) < 0.5).astype(np.int64)
    sign = 2.0 * labels - 1.0
    # every error <= 0 (all pixels classified with margin >= 1): loss 0, gradient 0
    loss, grad = check(ssl, dev, (sign * 1.0).astype(np.float32), labels, per_image=True)
    assert loss == 0.0 and not grad.any()
    loss, grad = check(ssl, dev, (sign * 4.0).astype(np.float32), labels, per_image=False)
    assert loss == 0.0 and not grad.any()
    # ties: logits from a handful of values -> equal errors, stable order by pixel index decides the deltas
    tied = rng.choice(np.array([-2.0, -0.5, 0.0, 0.5, 0.999, 3.0], np.float32), (b, h, w))
    check(ssl, dev, tied, labels, per_image=True)
    check(ssl, dev, tied, labels, per_image=False)
    # image 0 only void, image 1 without foreground, image 2 only foreground
    lab = labels.copy()
    lab[0], lab[1], lab[2] = 255, 0, 1
    x = (rng.standard_normal((b, h, w)) * 2).astype(np.float32)
    check(ssl, dev, x, lab, per_image=True, ignore=255)
    check(ssl, dev, x, lab, per_image=False, ignore=255)
    # no valid pixel at all
    lab[:] = 255
    loss, grad = run(ssl, dev, x, lab, per_image=True, ignore=255)
    assert loss == 0.0 and not grad.any()
    loss, grad = run(ssl, dev, x, lab, per_image=False, ignore=255)
    assert loss == 0.0 and not grad.any()
    # one pixel
    check(ssl, dev, np.array([[[0.25]]], np.float32), np.array([[[1]]], np.int64), per_image=True)
    check(ssl, dev, np.array([[[0.25]]], np.float32), np.array([[[0]]], np.int64), per_image=False)


def test_hinge_large(ssl, dev):
    rng = np.random.default_rng(9)
    b, h, w = 2, 1024, 2048
    labels = (rng.random((b, h // 32, w // 32)) < 0.35).repeat(32, 1).repeat(32, 2).astype(np.int64)
    logits = ((2.0 * labels - 1.0) * 0.8 + rng.standard_normal((b, h, w)) * 1.5).astype(np.float32)
    labels[rng.random((b, h // 16, w // 16)).repeat(16, 1).repeat(16, 2) < 0.03] = 255
    check(ssl, dev, logits, labels, torch.uint8, per_image=True, ignore=255)
    check(ssl, dev, logits, labels, torch.uint8, per_image=False, ignore=255)


def test_hinge_descriptor_is_checked(ssl, dev):
    from b200ssl import _lib
    d = _lib.LovaszDesc()
    d.n_images, d.n_channels, d.hw, d.per_image = 1, 2, 16, 1
    d.class_mode, d.n_list = _lib.LOVASZ_LIST, 1
    d.class_list[0] = 1
    d.error_mode = _lib.LOVASZ_ERR_HINGE                       # two channels: not a logit map
    assert _lib.lib.b200ssl_lovasz_workspace_bytes(C.byref(d)) >= 0
    x = torch.zeros(1, 2, 4, 4, device=dev)
    lab = torch.zeros(1, 4, 4, dtype=torch.int64, device=dev)
    out = torch.zeros(8, device=dev)
    meta = torch.zeros(8, dtype=torch.int32, device=dev)
    ws = torch.zeros(1 << 20, dtype=torch.uint8, device=dev)
    rc = _lib.lib.b200ssl_lovasz_forward(C.byref(d), x.data_ptr(), lab.data_ptr(), out.data_ptr(), out.data_ptr() + 4,
                                         meta.data_ptr(), meta.data_ptr() + 4, x.data_ptr(), ws.data_ptr(), ws.numel(),
                                         None)
    assert rc != 0 and b"hinge" in _lib.lib.b200ssl_last_error()
    d.n_channels, d.error_mode = 1, 7
    rc = _lib.lib.b200ssl_lovasz_forward(C.byref(d), x.data_ptr(), lab.data_ptr(), out.data_ptr(), out.data_ptr() + 4,
                                         meta.data_ptr(), meta.data_ptr() + 4, x.data_ptr(), ws.data_ptr(), ws.numel(),
                                         None)
    assert rc != 0 and b"error_mode" in _lib.lib.b200ssl_last_error()


@pytest.mark.parametrize("per_image", [True, False])
def test_hinge_fused_forward_backward(ssl, dev, per_image):
    """b200ssl_lovasz_forward_backward with the hinge error: the last radix pass writes the FINAL gradient
    RN(scale * delta) * sign for a known upstream gradient (here 0.5)."""
    from b200ssl import _lib
    rng = np.random.default_rng(21)
    b, h, w = 3, 72, 100
    labels = (rng.random((b, h, w)) < 0.4).astype(np.int64)
    labels[rng.random((b, h, w)) < 0.05] = 255
    logits = ((2.0 * (labels == 1) - 1.0) * 0.7 + rng.standard_normal((b, h, w)) * 1.5).astype(np.float32)
    x = torch.from_numpy(logits).to(dev)
    lab = torch.from_numpy(labels).to(dev)
    d = _lib.LovaszDesc()
    d.n_images, d.n_channels, d.hw, d.per_image = b, 1, h * w, int(per_image)
    d.class_mode, d.n_list = _lib.LOVASZ_LIST, 1
    d.class_list[0] = 1
    d.has_ignore, d.ignore_index, d.label_dtype = 1, 255, _lib.I64
    d.error_mode = _lib.LOVASZ_ERR_HINGE
    n_seg = _lib.lib.b200ssl_lovasz_num_segments(C.byref(d))
    assert n_seg == (b if per_image else 1)
    ws = torch.empty(_lib.lib.b200ssl_lovasz_workspace_bytes(C.byref(d)), dtype=torch.uint8, device=dev)
    go = torch.tensor([0.5], device=dev)
    loss = torch.empty(1, device=dev)
    seg_loss = torch.empty(n_seg, device=dev)
    meta = torch.empty((2, n_seg), dtype=torch.int32, device=dev)
    grad = torch.full_like(x, 7.0)
    _lib.check(_lib.lib.b200ssl_lovasz_forward_backward(
        C.byref(d), x.data_ptr(), lab.data_ptr(), go.data_ptr(), None, loss.data_ptr(), None, seg_loss.data_ptr(),
        meta[0].data_ptr(), meta[1].data_ptr(), grad.data_ptr(), ws.data_ptr(), ws.numel(),
        _lib.stream_ptr(dev)), "lovasz_forward_backward")
    o_loss, o_grad = oracle.lovasz_hinge(logits, labels, per_image=per_image, ignore=255, grad_out=0.5)
    assert abs(float(loss) - float(o_loss)) <= REL * max(1.0, abs(float(o_loss)))
    assert np.array_equal(bits(grad.cpu().numpy() + 0.0), bits(o_grad + 0.0))
    # seg_fg / seg_valid count every valid pixel, sorted or not
    groups = [labels[i] for i in range(b)] if per_image else [labels]
    assert meta[0].cpu().tolist() == [int((g == 1).sum()) for g in groups]
    assert meta[1].cpu().tolist() == [int((g != 255).sum()) for g in groups]


def test_hinge_float_and_bool_masks(ssl, dev):
    """the reference takes float (and bool) masks: `signs = 2. * labels.float() - 1.` (lovasz.py:105)"""
    rng = np.random.default_rng(33)
    labels = (rng.random((2, 30, 44)) < 0.5).astype(np.int64)
    logits = rng.standard_normal((2, 30, 44)).astype(np.float32)
    want, want_grad = oracle.lovasz_hinge(logits, labels, per_image=True)
    for dtype in (torch.float32, torch.bool):
        loss, grad = run(ssl, dev, logits, labels, dtype, per_image=True)
        assert abs(loss - float(want)) <= REL * max(1.0, abs(float(want)))
        assert np.array_equal(bits(grad + 0.0), bits(want_grad + 0.0))
