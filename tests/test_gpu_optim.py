"""Row N4 on the GPU: b200ssl.optim (clip_grad_norm_ + FusedSGD with the EMA epilogue) against the
oracle and against the golden vectors made by torch.optim.SGD + the reference's EMA (train.py:122-130).
Bars: total norm <= 1e-6 relative; parameters, momentum buffers and teacher bit-exact when the oracle is
given the same clip coefficient, and within 1e-5 (relative, atol 1e-8) of torch's own result."""
import numpy as np
import pytest
import torch

import oracle
from conftest import load_golden
from test_oracle_golden import SGD_VARIANTS

pytestmark = pytest.mark.gpu


def u32(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def ssl():
    import b200ssl
    return b200ssl


@pytest.mark.parametrize("name", list(SGD_VARIANTS))
def test_fused_sgd_matches_golden_and_oracle(ssl, name):
    dev = torch.device("cuda:0")
    g = load_golden("sgd")
    v = dict(SGD_VARIANTS[name])
    clip, alpha = v.pop("clip"), v.pop("alpha")
    n = int(g["n"])
    params = [torch.nn.Parameter(torch.from_numpy(g[f"param0_{i}"].copy()).to(dev)) for i in range(n)]
    ema = [torch.from_numpy(g[f"ema0_{i}"].copy()).to(dev) for i in range(n)]
    o_params = [g[f"param0_{i}"].copy() for i in range(n)]
    o_ema = [g[f"ema0_{i}"].copy() for i in range(n)]
    o_moms = [np.zeros_like(p) for p in o_params] if v.get("momentum", 0.0) else None
    opt = ssl.optim.FusedSGD(params, **v)
    for step in range(3):
        grads = [g[f"grad{step}_{i}"] for i in range(n)]
        for p, gr in zip(params, grads):
            p.grad = torch.from_numpy(gr.copy()).to(dev)
        tn = opt.step(max_grad_norm=clip, ema_params=ema, ema_alpha=alpha, zero_grad=True)
        coef = None
        if clip is not None:
            tn = np.float32(tn.item())
            ref = float(g[f"{name}_norm{step}"])
            assert abs(float(tn) - ref) <= 1e-6 * ref
            assert abs(float(tn) - float(oracle.grad_total_norm(grads))) <= 1.2e-7 * ref   # fp64 sums: <= 1 ulp apart
            coef = oracle.clip_coef(tn, clip)
        oracle.sgd_ema_step(o_params, grads, o_moms, o_ema, first_step=(step == 0), coef=coef, ema_alpha=alpha, **v)
        for i in range(n):
            assert np.array_equal(u32(params[i].detach().cpu().numpy()), u32(o_params[i])), (name, step, i)
            assert np.array_equal(u32(ema[i].cpu().numpy()), u32(o_ema[i])), (name, step, i)
            if o_moms is not None:
                assert np.array_equal(u32(opt.state[params[i]]["momentum_buffer"].cpu().numpy()), u32(o_moms[i]))
            assert not params[i].grad.any()                                   # zero_grad=True
    for i in range(n):
        assert np.allclose(params[i].detach().cpu().numpy(), g[f"{name}_s2_param{i}"], rtol=1e-5, atol=1e-8)
        assert np.allclose(ema[i].cpu().numpy(), g[f"{name}_s2_ema{i}"], rtol=1e-5, atol=1e-8)


def test_clip_grad_norm_matches_torch(ssl):
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(3)
    shapes = [(5,), (64, 3, 3, 3), (4099,), (300, 301), (1,)]
    for max_norm in (0.5, 1e6):
        ps = [torch.nn.Parameter(torch.zeros(s, device=dev)) for s in shapes]
        ref = [torch.nn.Parameter(torch.zeros(s)) for s in shapes]
        for p, r in zip(ps, ref):
            gr = torch.randn(p.shape, generator=gen)
            p.grad, r.grad = gr.to(dev), gr.clone()
        skip = torch.nn.Parameter(torch.zeros(7, device=dev))                   # no gradient: ignored like torch
        tn = ssl.optim.clip_grad_norm_(ps + [skip], max_norm)
        tn_ref = torch.nn.utils.clip_grad_norm_(ref, max_norm)
        assert abs(float(tn) - float(tn_ref)) <= 1e-6 * float(tn_ref)
        coef = oracle.clip_coef(np.float32(tn.item()), max_norm)
        for p, r in zip(ps, ref):
            want = (r.grad.numpy() / 1.0)
            # r.grad has been scaled by torch with ITS coefficient; ours is exact given our norm
            assert np.allclose(p.grad.cpu().numpy(), want, rtol=1e-6, atol=0)
        assert float(coef) <= 1.0


def test_sgd_large_unaligned_and_views(ssl):
    """58-tensor mix of odd sizes, a parameter that is a view at a 4-byte (not 16-byte) aligned offset."""
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(11)
    base = torch.randn(10007, generator=gen).to(dev)
    sizes = [1, 2, 3, 5, 31, 127, 4095, 4096, 4097, 12289, 70001]
    params = [torch.nn.Parameter(base[1:1 + 4099])] + [torch.nn.Parameter(torch.randn(s, generator=gen).to(dev)) for s in sizes]
    ema = [torch.randn(p.shape, generator=gen).to(dev) for p in params]
    o_params = [p.detach().cpu().numpy().copy() for p in params]
    o_ema = [e.cpu().numpy().copy() for e in ema]
    o_moms = [np.zeros_like(p) for p in o_params]
    opt = ssl.optim.FusedSGD(params, lr=0.03, momentum=0.9, weight_decay=1e-4, nesterov=True)
    for step in range(2):
        grads = [torch.randn(p.shape, generator=gen) for p in params]
        for p, gr in zip(params, grads):
            p.grad = gr.to(dev)
        tn = opt.step(max_grad_norm=2.0, ema_params=ema, ema_alpha=0.99)
        coef = oracle.clip_coef(np.float32(tn.item()), 2.0)
        oracle.sgd_ema_step(o_params, [x.numpy() for x in grads], o_moms, o_ema, lr=0.03, momentum=0.9,
                            weight_decay=1e-4, nesterov=True, first_step=(step == 0), coef=coef, ema_alpha=0.99)
        for i in range(len(params)):
            assert np.array_equal(u32(params[i].detach().cpu().numpy()), u32(o_params[i])), (step, i)
            assert np.array_equal(u32(ema[i].cpu().numpy()), u32(o_ema[i])), (step, i)
            assert torch.equal(params[i].grad.cpu(), grads[i])                 # zero_grad=False: untouched


def test_optim_refuses_cpu_tensors(ssl):
    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.ones(4)
    with pytest.raises(RuntimeError):
        ssl.optim.clip_grad_norm_([p], 1.0)


def _twin(shapes, dev, gen):
    a = [torch.nn.Parameter(torch.randn(s, generator=gen).to(dev)) for s in shapes]
    b = [torch.nn.Parameter(p.detach().cpu().clone()) for p in a]      # torch.optim.SGD + reference EMA on the CPU
    return a, b


def _same(a, b):
    return all(np.array_equal(u32(x.detach().cpu().numpy()), u32(y.detach().cpu().numpy())) for x, y in zip(a, b))


def test_fused_sgd_lifecycle_like_torch(ssl):
    """torch.optim.SGD behaviours the cached tensor lists must not break: a parameter whose first
    gradient arrives later than the others, a frozen parameter whose teacher copy still moves
    (mean_teacher.py:10-11), load_state_dict between steps, add_param_group, and a returned norm that
    survives the next step.  Bit-exact against torch.optim.SGD + the reference's EMA (CPU, like the goldens)."""
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(21)
    shapes = [(33,), (17, 5), (4100,), (3,)]
    mine, ref = _twin(shapes, dev, gen)
    ema = [torch.randn(s, generator=gen).to(dev) for s in shapes]
    ema_ref = [e.cpu().clone() for e in ema]
    kw = dict(lr=0.05, momentum=0.9, weight_decay=5e-4)
    opt, opt_ref = ssl.optim.FusedSGD(mine, **kw), torch.optim.SGD(ref, **kw)
    alpha = 0.99

    def ref_ema():
        for e, p in zip(ema_ref, ref):
            e.mul_(alpha).add_(p.detach(), alpha=1 - alpha)

    def give(idx):
        for k in idx:
            gr = torch.randn(shapes[k], generator=gen)
            mine[k].grad, ref[k].grad = gr.to(dev), gr.clone()

    # step 0: parameter 3 is frozen, parameter 2 has no gradient yet
    give([0, 1])
    opt.step(ema_params=ema, ema_alpha=alpha); opt_ref.step(); ref_ema()
    assert _same(mine, ref) and _same(ema, ema_ref)
    # step 1: parameter 2 receives its first gradient one step late (its buffer starts as a copy of it)
    give([0, 1, 2])
    opt.step(ema_params=ema, ema_alpha=alpha); opt_ref.step(); ref_ema()
    assert _same(mine, ref) and _same(ema, ema_ref)
    give([0, 1, 2])
    opt.step(ema_params=ema, ema_alpha=alpha); opt_ref.step(); ref_ema()
    assert _same(mine, ref) and _same(ema, ema_ref)
    # load_state_dict: new momentum tensors must be the ones the next step uses
    sd = opt_ref.state_dict()
    for st in sd["state"].values():
        st["momentum_buffer"] = st["momentum_buffer"] * 0.5
    opt.load_state_dict(sd); opt_ref.load_state_dict(sd)
    give([0, 1, 2])
    n1 = opt.step(max_grad_norm=1e9, ema_params=ema, ema_alpha=alpha); opt_ref.step(); ref_ema()
    assert _same(mine, ref) and _same(ema, ema_ref)
    # add_param_group + the returned norm is a copy
    extra, extra_ref = _twin([(9,)], dev, gen)
    opt.add_param_group({"params": extra}); opt_ref.add_param_group({"params": extra_ref})
    ema.append(torch.zeros(9, device=dev)); ema_ref.append(torch.zeros(9))
    mine += extra; ref += extra_ref; shapes.append((9,))
    give([0, 1, 2, 4])
    kept = float(n1)
    n2 = opt.step(max_grad_norm=1e9, ema_params=ema, ema_alpha=alpha); opt_ref.step(); ref_ema()
    assert _same(mine, ref) and _same(ema, ema_ref)
    assert float(n1) == kept and float(n2) != kept


def test_clip_coefficient_nan_poisons_like_torch(ssl):
    dev = torch.device("cuda:0")
    ps = [torch.nn.Parameter(torch.zeros(8, device=dev)) for _ in range(2)]
    ps[0].grad = torch.full((8,), float("nan"), device=dev)
    ps[1].grad = torch.ones(8, device=dev)
    tn = ssl.optim.clip_grad_norm_(ps, 1.0)
    assert torch.isnan(tn) and torch.isnan(ps[1].grad).all()            # torch.clamp propagates NaN
