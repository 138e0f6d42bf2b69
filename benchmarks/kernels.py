#!/usr/bin/env python
"""Stand-alone timings of the path's kernels on the other BASELINE.json configs (parity-test shapes,
not the bench.py line): confusion-matrix sweep (configs[4]), EMA parameter sets (configs[2], [3]),
multi-class Lovasz (configs[2], [3] shapes at a reduced batch) and the mix at 19/21 classes.
Prints one JSON object; CUDA-event timing, median of `reps` after warm-up, inputs larger than L2 or
rotated across buffers.   python benchmarks/kernels.py [--quick]"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200ssl  # noqa: E402
import bench    # noqa: E402

dev = torch.device("cuda:0")
PEAK, _ = bench.measured_peak()


def timeit(fn, reps=10, warm=3, inner=10):
    """median over `reps` of (CUDA-event time of `inner` back-to-back calls) / inner: the queue stays
    full, so host launch latency does not leak into short kernels"""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(inner):
            fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b) / inner)
    ts.sort()
    return ts[len(ts) // 2]


def coherent(n, c, h, w, gen):
    """spatially coherent label maps (argmax of noise blurred at ~16 px)"""
    x = torch.randn(n, c, h // 32, w // 32, device=dev, generator=gen)
    x = torch.nn.functional.interpolate(x, size=(h, w), mode="bilinear")
    return x.argmax(1)


def main():
    quick = "--quick" in sys.argv
    gen = torch.Generator(device=dev).manual_seed(0)
    out = {"peak_GBps": PEAK}

    only_optim = "--only-optim" in sys.argv or "--only-new" in sys.argv
    # ---- configs[4]: 19-class confusion matrix over 1024x2048 masks, chunks of 32 masks (> L2)
    n, h, w, c = (1, 64, 64, 19) if only_optim else (32, 1024, 2048, 19)
    lab = coherent(n, c, h, w, gen)
    wrong = coherent(n, 5, h, w, gen) == 0                      # ~20 % of the area, in blobs
    prd = torch.where(wrong, coherent(n, c, h, w, gen), lab)    # predictions agree except in blobs
    lab[coherent(n, 30, h, w, gen) == 0] = 255                  # ~3 % void, in blobs
    for dt, name in [(torch.int64, "int64"), (torch.uint8, "uint8")]:
        l, p = lab.to(dt), prd.to(dt)
        acc = torch.zeros(c, c, dtype=torch.int64, device=dev)
        ms = timeit(lambda: b200ssl.metrics.confusion_matrix(l, p, c, ignore_index=255, out=acc))
        bytes_ = 2 * l.numel() * l.element_size()
        out[f"cm_{name}_19c_32x1024x2048"] = {"ms": round(ms, 4), "Gpix_s": round(l.numel() / ms / 1e6, 2),
                                             "GBps": round(bytes_ / ms / 1e6, 1), "frac": round(bytes_ / ms / 1e6 / PEAK, 3)}
    logits = torch.randn(min(8, n), c, h, w, device=dev, generator=gen)
    l8 = lab[:8].to(torch.uint8)
    ms = timeit(lambda: b200ssl.metrics.confusion_matrix_from_logits(logits, l8, ignore_index=255))
    bytes_ = logits.numel() * 4 + l8.numel()
    out["cm_from_logits_19c_8x1024x2048"] = {"ms": round(ms, 4), "GBps": round(bytes_ / ms / 1e6, 1),
                                             "frac": round(bytes_ / ms / 1e6 / PEAK, 3)}
    del logits, lab, prd

    # ---- EMA parameter sets
    for key in ([] if only_optim else ["unet_mnv2_c2", "simple_unet_c2", "deeplabv3_r101_c21"]):
        shapes = bench.load_param_shapes(key)
        ps = [torch.randn(s, device=dev, generator=gen) for s in shapes]
        es = [torch.randn(s, device=dev, generator=gen) for s in shapes]
        upd = b200ssl.mean_teacher.EmaUpdater()
        ms = timeit(lambda: upd(es, ps, 0.99))
        npar = sum(p.numel() for p in ps)
        out[f"ema_{key}"] = {"params": npar, "tensors": len(ps), "ms": round(ms, 4),
                             "GBps": round(12 * npar / ms / 1e6, 1), "frac": round(12 * npar / ms / 1e6 / PEAK, 3)}
        # the reference's own op sequence on the same GPU (2 launches per tensor)
        def ref():
            for e, p in zip(es, ps):
                e.mul_(0.99).add_(p, alpha=1 - 0.99)
        ms_ref = timeit(ref, reps=3, warm=1, inner=2)
        out[f"ema_{key}"]["torch_cuda_ref_ms"] = round(ms_ref, 4)
        del ps, es

    # ---- row N4: clip_grad_norm_ + SGD.step + zero_grad + EMA (train.py:122-130), fused vs torch's own
    # foreach CUDA path on the same GPU
    for key in ([] if "--only-new" in sys.argv else ["unet_mnv2_c2", "deeplabv3_r101_c21"]):
        shapes = bench.load_param_shapes(key)
        ps = [torch.nn.Parameter(torch.randn(s, device=dev, generator=gen)) for s in shapes]
        es = [torch.randn(s, device=dev, generator=gen) for s in shapes]
        gs = [torch.randn(s, device=dev, generator=gen) for s in shapes]
        for p, g in zip(ps, gs):
            p.grad = g
        opt = b200ssl.optim.FusedSGD(ps, lr=2.25e-4, momentum=0.9, weight_decay=5e-4)
        ms = timeit(lambda: opt.step(max_grad_norm=5.0, ema_params=es, ema_alpha=0.99, zero_grad=False))
        npar = sum(p.numel() for p in ps)
        alg = (4 + 28) * npar                     # norm pass reads g; update reads p,g,b,e and writes p,b,e
        ref_ps = [torch.nn.Parameter(p.detach().clone()) for p in ps]
        for p, g in zip(ref_ps, gs):
            p.grad = g.clone()
        ref_opt = torch.optim.SGD(ref_ps, lr=2.25e-4, momentum=0.9, weight_decay=5e-4)
        ref_es = [e.clone() for e in es]

        def ref():
            torch.nn.utils.clip_grad_norm_(ref_ps, 5.0)
            ref_opt.step()
            for e, p in zip(ref_es, ref_ps):
                e.mul_(0.99).add_(p.data, alpha=1 - 0.99)
        ms_ref = timeit(ref, reps=3, warm=1, inner=2)
        out[f"sgd_clip_ema_{key}"] = {"params": npar, "tensors": len(ps), "ms": round(ms, 4),
                                      "GBps": round(alg / ms / 1e6, 1), "frac": round(alg / ms / 1e6 / PEAK, 3),
                                      "torch_cuda_ref_ms": round(ms_ref, 4), "speedup": round(ms_ref / ms, 1)}
        del ps, es, gs, ref_ps, ref_es, opt, ref_opt
    if "--only-optim" in sys.argv:
        print(json.dumps(out))
        return

    # ---- multi-class Lovasz forward+backward (fused), configs[2]/[3] class counts
    for (n, c, h, w) in ([(4, 21, 512, 512)] if quick else [(4, 21, 512, 512), (1, 19, 1024, 2048)]):
        probas = torch.softmax(torch.randn(n, c, h, w, device=dev, generator=gen) * 2, 1)
        labels = coherent(n, c, h, w, gen)
        labels[torch.rand(n, h, w, device=dev, generator=gen) < 0.03] = 255
        step = b200ssl.LossPathStep(num_classes=c, mode="softmax", classes="present", per_image=False, ignore=255)
        ms = timeit(lambda: step.lovasz_loss_and_grad(probas, labels), reps=5, inner=4)
        P = n * h * w
        alg = (8 * c + 8) * P
        out[f"lovasz_softmax_present_{n}x{c}x{h}x{w}"] = {"ms": round(ms, 4), "Mpix_s": round(P / ms / 1e3, 1),
                                                       "alg_GBps": round(alg / ms / 1e6, 1),
                                                       "frac": round(alg / ms / 1e6 / PEAK, 3),
                                                       "Gkeys_s": round(P * c / ms / 1e6, 2)}
        del probas, labels

    # ---- fused mix at 21 classes (configs[2]) 8x512x512
    n, c, h, w = 8, 21, 512, 512
    ia, ib = (torch.rand(n, 3, h, w, device=dev, generator=gen) for _ in range(2))
    ta, tb = (torch.randn(n, c, h, w, device=dev, generator=gen) for _ in range(2))
    mask = (torch.rand(n, 1, h, w, device=dev, generator=gen) > 0.5).float()
    ms = timeit(lambda: b200ssl.cowmix.mix2_with_mask(ia, ib, ta, tb, mask))
    bytes_ = 4 * (3 * 3 + 3 * c + 1) * n * h * w
    out[f"mix2_{n}x{c}x{h}x{w}"] = {"ms": round(ms, 4), "GBps": round(bytes_ / ms / 1e6, 1),
                                    "frac": round(bytes_ / ms / 1e6 / PEAK, 3)}
    # ---- row N3: lovasz_softmax from logits, forward + backward, vs soft-max by torch + lovasz_softmax
    for (n, c, h, w) in [(4, 21, 512, 512)]:
        logits = torch.randn(n, c, h, w, device=dev, generator=gen) * 2
        labels = coherent(n, c, h, w, gen)
        labels[torch.rand(n, h, w, device=dev, generator=gen) < 0.03] = 255

        def fused():
            x = logits.detach().requires_grad_(True)
            b200ssl.lovasz.lovasz_softmax_with_logits(x, labels, ignore=255).backward()

        def unfused():
            x = logits.detach().requires_grad_(True)
            b200ssl.lovasz.lovasz_softmax(torch.softmax(x, 1), labels, ignore=255).backward()
        ms_f = timeit(fused, reps=5, inner=4)
        ms_u = timeit(unfused, reps=5, inner=4)
        P = n * h * w
        def nomat():
            x = logits.detach().requires_grad_(True)
            b200ssl.lovasz.lovasz_softmax_with_logits(x, labels, ignore=255, materialize=False).backward()
        ms_n = timeit(nomat, reps=5, inner=4)
        out[f"lovasz_from_logits_{n}x{c}x{h}x{w}"] = {"fused_ms": round(ms_f, 4), "never_materialised_ms": round(ms_n, 4),
                                                      "torch_softmax_plus_lovasz_ms": round(ms_u, 4),
                                                     "Mpix_s": round(P / ms_f / 1e3, 1), "gain": round(ms_u / ms_f, 3)}
        del logits, labels

    # ---- row N2: teacher logits at stride 4 up-sampled inside the mix vs F.interpolate x2 + mix
    for (n, c, h, w) in [(16, 2, 512, 512), (8, 19, 512, 1024)]:
        ia, ib = (torch.rand(n, 3, h, w, device=dev, generator=gen) for _ in range(2))
        ta, tb = (torch.randn(n, c, h // 4, w // 4, device=dev, generator=gen) for _ in range(2))
        mask = (torch.rand(n, 1, h, w, device=dev, generator=gen) > 0.5).float()
        ms_f = timeit(lambda: b200ssl.cowmix.mix2_with_mask(ia, ib, ta, tb, mask))

        def unfused():
            ua = torch.nn.functional.interpolate(ta, (h, w), mode="bilinear", align_corners=False)
            ub = torch.nn.functional.interpolate(tb, (h, w), mode="bilinear", align_corners=False)
            b200ssl.cowmix.mix2_with_mask(ia, ib, ua, ub, mask)
        ms_u = timeit(unfused)
        bytes_ = (4 * (3 * 3 + c + 1) + 8 * c / 16) * n * h * w
        out[f"mix2_upsampled_{n}x{c}x{h}x{w}"] = {"ms": round(ms_f, 4), "GBps": round(bytes_ / ms_f / 1e6, 1),
                                                  "frac": round(bytes_ / ms_f / 1e6 / PEAK, 3),
                                                  "interpolate_x2_plus_mix2_ms": round(ms_u, 4), "gain": round(ms_u / ms_f, 2)}
        del ia, ib, ta, tb, mask
    if "--only-new" in sys.argv:
        print(json.dumps(out))
        return

    # ---- the whole configs[1] step: this library vs the reference's ATen op sequence on the SAME GPU
    # (stock ATen/cub/cuDNN kernels; SURVEY 2.2: "the Blackwell kernel to beat").  oracle.torch_port is
    # test infrastructure; it is imported here only as the thing being compared against.
    from oracle import torch_port
    inp = bench.make_inputs(dev, 0)
    W = bench.WORKLOAD
    P = W["n"] * W["h"] * W["w"]
    step = b200ssl.LossPathStep(num_classes=W["c"], mode="binary")
    ms = timeit(lambda: step(inp["image_a"], inp["image_b"], inp["teacher_a"], inp["teacher_b"], inp["scores"],
                             inp["target"], inp["params"], inp["ema_params"]), reps=5, inner=20)
    ema_ref = [e.clone() for e in inp["ema_params"]]

    def aten_step():
        torch_port.loss_path_step(inp["image_a"], inp["image_b"], inp["teacher_a"], inp["teacher_b"], inp["scores"],
                                  inp["target"], inp["params"], ema_ref, mode="binary", num_classes=W["c"])
    ms_ref = timeit(aten_step, reps=3, warm=2, inner=2)
    out["step_configs1_16x512x512"] = {"b200ssl_ms": round(ms, 4), "b200ssl_Mpix_s": round(P / ms / 1e3, 1),
                                       "aten_cuda_ms": round(ms_ref, 3), "aten_cuda_Mpix_s": round(P / ms_ref / 1e3, 1),
                                       "speedup_vs_aten_cuda": round(ms_ref / ms, 1)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
