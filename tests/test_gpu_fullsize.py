"""BASELINE.json configs[2] and configs[3] at their FULL per-GPU sizes, through properties that need no
CPU oracle at that size (the oracle would take minutes on 0.3 G keys):

  telescoping   lovasz_grad's deltas of one class sum to its final Jaccard value, 1 (lovasz.py:19-31), so
                sum_pixels |dL/dp_c| * (number of counted classes) == 1 for every present class
  consistency   the loss is the dot product of the errors with those same deltas (lovasz.py:200):
                loss == sum_c sum_pixels |fg_c - p_c| * |dL/dp_c|   -- ties loss and gradient together
  sign          dL/dp_c has the sign of -(fg_c - p_c) wherever it is non-zero (abs backward)
  void          ignored pixels get an exactly zero gradient
  determinism   a second run is bit-identical (the sort is stable, the reductions have a fixed order)
and the multi-tensor EMA / SGD on the configs[2]/[3] parameter sets against a per-tensor torch loop on the
same GPU (same arithmetic, so bit-exact)."""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ssl():
    import b200ssl
    return b200ssl


def coherent_labels_gpu(n, c, h, w, gen, dev):
    x = torch.randn(n, c, h // 32, w // 32, device=dev, generator=gen)
    return torch.nn.functional.interpolate(x, size=(h, w), mode="bilinear").argmax(1)


@pytest.mark.parametrize("n,c,h,w,from_logits", [(32, 21, 512, 512, False),      # configs[2] on one GPU
                                                 (8, 19, 1024, 2048, False),    # configs[3], per GPU
                                                 (4, 21, 512, 512, True)])      # configs[2] per GPU (8-way), row N3
def test_lovasz_full_size_properties(ssl, n, c, h, w, from_logits):
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(n * 1000 + c)
    logits = torch.randn(n, c, h, w, device=dev, generator=gen) * 2
    labels = coherent_labels_gpu(n, c, h, w, gen, dev)
    labels[torch.rand(n, h, w, device=dev, generator=gen) < 0.03] = 255
    present = [k for k in range(c) if bool((labels == k).any())]
    assert len(present) >= 2
    runs = []
    for _ in range(2):
        if from_logits:
            x = logits.clone().requires_grad_(True)
            loss = ssl.lovasz.lovasz_softmax_with_logits(x, labels, classes="present", ignore=255)
            # the properties are stated on dL/dp: take the probabilities' gradient from the unfused graph
            probas = torch.softmax(logits, 1).requires_grad_(True)
            loss_p = ssl.lovasz.lovasz_softmax(probas, labels, classes="present", ignore=255)
            loss_p.backward()
            loss.backward()
            assert abs(float(loss) - float(loss_p)) <= 1e-5 * float(loss_p)
            p64, g64 = probas.detach().double(), probas.grad.double()
            want = (g64 - (g64 * p64).sum(1, keepdim=True)) * p64                 # soft-max backward in fp64
            err = (x.grad.double() - want).norm() / want.norm()
            assert float(err) <= 1e-5
            runs.append((float(loss), x.grad.clone()))
            grad, pr = probas.grad, probas.detach()
        else:
            pr = torch.softmax(logits, 1).requires_grad_(True)
            loss = ssl.lovasz.lovasz_softmax(pr, labels, classes="present", ignore=255)
            loss.backward()
            runs.append((float(loss), pr.grad.clone()))
            grad, pr = pr.grad, pr.detach()
    assert runs[0][0] == runs[1][0] and torch.equal(runs[0][1], runs[1][1])      # determinism
    assert 0.0 < runs[0][0] <= 1.0
    valid = labels != 255
    total = 0.0
    for k in range(c):
        gk = grad[:, k]
        if k not in present:
            assert not bool(gk.any())
            continue
        s = float(gk.double().abs().sum()) * len(present)
        assert abs(s - 1.0) <= 1e-4, (k, s)                                       # telescoping
        fg = (labels == k)
        diff = fg.to(pr.dtype) - pr[:, k]
        total += float((diff.abs().double() * gk.abs().double())[valid].sum())
        nz = gk != 0
        assert bool((torch.sign(gk[nz]) == -torch.sign(diff[nz])).all())          # sign
        assert not bool(gk[~valid].any())                                         # void pixels
    assert abs(total - runs[0][0]) <= 1e-5 * runs[0][0]                           # loss <-> gradient


@pytest.mark.parametrize("key", ["deeplabv3_r101_c21", "simple_unet_c2"])
def test_ema_and_sgd_full_parameter_sets(ssl, key):
    dev = torch.device("cuda:0")
    with open(os.path.join(ROOT, "tests", "golden", "param_shapes.json")) as f:
        shapes = json.load(f)[key]["shapes"]
    gen = torch.Generator(device=dev).manual_seed(7)
    ps = [torch.nn.Parameter(torch.randn(s, device=dev, generator=gen)) for s in shapes]
    es = [torch.randn(s, device=dev, generator=gen) for s in shapes]
    gs = [torch.randn(s, device=dev, generator=gen) * 0.1 for s in shapes]
    # the per-tensor reference runs on the CPU: that ATen path is the one the oracle is pinned on bit for bit
    ref_p = [p.detach().cpu().clone() for p in ps]
    ref_e = [e.cpu().clone() for e in es]
    cpu_g = [g.cpu() for g in gs]
    # EMA alone (mean_teacher.py:10-11 per tensor)
    ssl.mean_teacher.EmaUpdater()(es, [p.detach() for p in ps], 0.99)
    for e, p in zip(ref_e, ref_p):
        e.mul_(0.99).add_(p, alpha=1 - 0.99)
    assert all(torch.equal(a.cpu(), b) for a, b in zip(es, ref_e))
    # clip + SGD + EMA (train.py:122-130) against torch's own optimiser given the same clip coefficient
    for p, g in zip(ps, gs):
        p.grad = g.clone()
    opt = ssl.optim.FusedSGD(ps, lr=2.25e-4, momentum=0.9, weight_decay=5e-4)
    ref_params = [torch.nn.Parameter(p.clone()) for p in ref_p]
    ref_opt = torch.optim.SGD(ref_params, lr=2.25e-4, momentum=0.9, weight_decay=5e-4)
    norm = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g.double()) for g in cpu_g]))
    for step in range(2):
        tn = opt.step(max_grad_norm=5.0, ema_params=es, ema_alpha=0.99).cpu()
        assert abs(float(tn) - float(norm)) <= 1e-6 * float(norm)
        coef = torch.clamp((1.0 / (tn + 1e-6)) * 5.0, max=1.0)                    # torch's formula on OUR norm
        for p, g in zip(ref_params, cpu_g):
            p.grad = g * coef
        ref_opt.step()
        for e, p in zip(ref_e, ref_params):
            e.mul_(0.99).add_(p.data, alpha=1 - 0.99)
        assert all(torch.equal(a.data.cpu(), b.data) for a, b in zip(ps, ref_params)), step
        assert all(torch.equal(a.cpu(), b) for a, b in zip(es, ref_e)), step
