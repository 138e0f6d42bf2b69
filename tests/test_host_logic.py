"""Host-side logic of the drop-in modules that needs no GPU: CowMix parameter draws, taps and
kernel sizes (bit-identical to the reference's), Lovasz descriptor validation, reduce_tensor."""
import numpy as np
import pytest
import torch

import b200ssl
from b200ssl import _lib
from conftest import load_golden


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.mark.parametrize("name", ["cowmix_small", "cowmix_c1"])
def test_parameter_draws_and_taps_match_the_reference(name):
    g = load_golden(name)
    torch.manual_seed(int(g["seed"]))
    n = int(g["shape"][0])
    p, sig = b200ssl.cowmix.draw_mask_parameters(n, tuple(g["p_range"].tolist()), tuple(g["sigma_range"].tolist()))
    assert np.array_equal(bits(p.numpy()), bits(g["p"])) and np.array_equal(bits(sig.numpy()), bits(g["sigmas"]))
    size = b200ssl.cowmix.kernel_size_for(sig.max().item())
    assert size == int(g["size"])
    taps = b200ssl.cowmix.gaussian_taps(size, sig)
    assert taps.shape == (n, size)
    assert np.array_equal(bits(taps.numpy()), bits(g["taps"])), "batched taps differ from cowmix.py:6-24"


def test_kernel_size_known_answers():
    for sigma, k in [(4, 25), (7.9, 49), (8, 49), (16, 97), (31.99, 193), (32, 193), (2.5, 17)]:
        assert b200ssl.cowmix.kernel_size_for(sigma) == k


def test_lovasz_descriptor_validation():
    probas = torch.zeros(2, 3, 4, 4)
    labels = torch.zeros(2, 4, 4, dtype=torch.int64)
    mk = b200ssl.lovasz._make_desc
    d = mk(probas, labels, "present", False, None)
    assert (d.class_mode, d.n_images, d.n_channels, d.hw, d.has_ignore) == (_lib.LOVASZ_PRESENT, 2, 3, 16, 0)
    d = mk(probas, labels.to(torch.uint8), [2, 0], True, 255)
    assert (d.class_mode, d.n_list, d.class_list[0], d.class_list[1], d.ignore_index, d.label_dtype) == \
        (_lib.LOVASZ_LIST, 2, 2, 0, 255, _lib.U8)
    with pytest.raises(ValueError, match="Sigmoid output possible only with 1 class"):
        mk(probas[:, :1], labels, "present", False, None)           # lovasz.py:191-192
    with pytest.raises(ValueError, match="Sigmoid output possible only with 1 class"):
        mk(probas[:, :1], labels, [0, 1], False, None)
    with pytest.raises(IndexError):
        mk(probas, labels, [5], False, None)
    with pytest.raises(ValueError):
        mk(probas, labels, "some", False, None)
    with pytest.raises(TypeError):
        mk(probas, labels.float(), "all", False, None)


def test_reduce_tensor_without_process_group_is_identity():
    t = torch.tensor(3.0)
    assert b200ssl.utils.reduce_tensor(t) is t          # utils/utils.py:43-54 returns the same object


def test_iou_helpers_from_matrices():
    cm = torch.tensor([[5, 1], [2, 0]])
    iou = b200ssl.metrics.iou_from_cm(cm)
    assert torch.allclose(iou, torch.tensor([5 / 8, 0.0], dtype=torch.float64))
    empty = b200ssl.metrics.iou_from_cm(torch.zeros(3, 3, dtype=torch.int64))
    assert torch.all(empty == 1.0)                       # lovasz.py EMPTY convention
    assert float(b200ssl.metrics.miou_from_cm(cm)) == pytest.approx(5 / 16)
