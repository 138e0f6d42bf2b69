"""Row N2 on the GPU: bilinear up-sampling fused into the mix (b200ssl_mix2_upsampled) and stand-alone,
against the oracle / ATen goldens.  Bar: bit-exact for out >= in (train.py:72-75 only up-samples)."""
import numpy as np
import pytest
import torch

import oracle
from conftest import load_golden
from test_oracle_golden import UPSAMPLE_TAGS, UPSAMPLE_BIT_EXACT, unpack_bits

pytestmark = pytest.mark.gpu


def u32(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def ssl():
    import b200ssl
    return b200ssl


@pytest.mark.parametrize("tag", UPSAMPLE_TAGS)
def test_upsample_and_fused_mix_match_golden(ssl, tag):
    dev = torch.device("cuda:0")
    g = load_golden("upsample")
    a, b = g[f"{tag}_a"], g[f"{tag}_b"]
    H, W = (int(v) for v in g[f"{tag}_size"])
    n = a.shape[0]
    mask = unpack_bits(g[f"{tag}_mask_bits"], (n, 1, H, W))
    up = ssl.cowmix.upsample_bilinear(torch.from_numpy(a).to(dev), (H, W))
    assert np.array_equal(u32(up.cpu().numpy()), u32(oracle.upsample_bilinear(a, (H, W))))
    if tag in UPSAMPLE_BIT_EXACT:
        assert np.array_equal(u32(up.cpu().numpy()), u32(g[f"{tag}_up_a"]))
    gen = torch.Generator().manual_seed(1)
    ia, ib = torch.rand(n, 3, H, W, generator=gen), torch.rand(n, 3, H, W, generator=gen)
    mi, mt = ssl.cowmix.mix2_with_mask(ia.to(dev), ib.to(dev), torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev),
                                       torch.from_numpy(mask).to(dev))
    want = oracle.mix(oracle.upsample_bilinear(a, (H, W)), oracle.upsample_bilinear(b, (H, W)), mask)
    assert np.array_equal(u32(mt.cpu().numpy()), u32(want))
    if tag in UPSAMPLE_BIT_EXACT:
        assert np.array_equal(u32(mt.cpu().numpy()), u32(g[f"{tag}_mixed"]))
    else:
        assert np.allclose(mt.cpu().numpy(), g[f"{tag}_mixed"], rtol=1e-5, atol=1e-5 * float(np.abs(a).max()))
    assert np.array_equal(u32(mi.cpu().numpy()), u32(oracle.mix(ia.numpy(), ib.numpy(), mask)))


@pytest.mark.parametrize("shape", [(2, 2, 128, 128, 512, 512), (1, 19, 64, 128, 256, 512), (3, 2, 31, 45, 97, 131),
                                   (1, 1, 1, 1, 7, 9), (2, 3, 40, 40, 40, 40),
                                   # exact stride 4 (the constant-tap path) down to one- and two-column inputs
                                   (1, 2, 1, 1, 4, 4), (2, 3, 2, 3, 8, 12), (1, 1, 5, 2, 20, 8), (1, 4, 33, 65, 132, 260)])
def test_upsample_matches_oracle_and_torch_cuda(ssl, shape):
    dev = torch.device("cuda:0")
    n, c, h, w, H, W = shape
    gen = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(n, c, h, w, generator=gen) * 4
    up = ssl.cowmix.upsample_bilinear(x.to(dev), (H, W)).cpu().numpy()
    assert np.array_equal(u32(up), u32(oracle.upsample_bilinear(x.numpy(), (H, W))))
    ref = torch.nn.functional.interpolate(x.to(dev), (H, W), mode="bilinear", align_corners=False).cpu().numpy()
    assert np.allclose(up, ref, rtol=1e-5, atol=1e-5 * float(x.abs().max()))


def test_loss_path_step_with_low_resolution_teacher(ssl):
    """The step entry with stride-4 teacher logits equals up-sample-then-step, bit for bit."""
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(8)
    n, c, h, w = 2, 2, 64, 96
    img = [torch.rand(n, 3, h, w, generator=g).to(dev) for _ in range(2)]
    tea = [(torch.randn(n, c, h // 4, w // 4, generator=g) * 2).to(dev) for _ in range(2)]
    scores = (torch.randn(n, c, h, w, generator=g) * 3).to(dev)
    blob = torch.nn.functional.avg_pool2d(torch.randn(n, c, h, w, generator=g), 9, 1, 4)
    target = torch.nn.functional.one_hot(blob.argmax(1), c).permute(0, 3, 1, 2).float().contiguous().to(dev)
    outs = []
    for fused in (True, False):
        torch.manual_seed(0)
        step = ssl.LossPathStep(num_classes=c, sigma_range=(2, 4))
        ta, tb = tea if fused else [ssl.cowmix.upsample_bilinear(t, (h, w)) for t in tea]
        outs.append(step(img[0], img[1], ta, tb, scores, target, None, None))
    torch.cuda.synchronize()
    for k in ("mask", "mixed_images", "mixed_teacher", "grad"):
        assert torch.equal(outs[0][k], outs[1][k]), k
    want = oracle.mix(oracle.upsample_bilinear(tea[0].cpu().numpy(), (h, w)),
                      oracle.upsample_bilinear(tea[1].cpu().numpy(), (h, w)), outs[0]["mask"].cpu().numpy())
    assert np.array_equal(u32(outs[0]["mixed_teacher"].cpu().numpy()), u32(want))
