"""train.py:69-82 + 98-107, forward + backward: F.interpolate x2 + mix + loss (ATen), this repo's three calls
(mix2_with_mask with the low-resolution teacher pair, then confidence_masked_consistency) and the fused
confidence_masked_consistency_mixed, which never materialises mixed_ema_pred.
python benchmarks/consistency_mixed.py > gpurun_out/consistency_mixed.json"""
import importlib
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
b200ssl = importlib.import_module("semi-supervised_semantic_segmentation_b200")


def aten(student, ta, tb, mask, thr):
    size = student.shape[2:]
    a = F.interpolate(ta, size, mode="bilinear", align_corners=False)
    b = F.interpolate(tb, size, mode="bilinear", align_corners=False)
    mixed = a * mask + b * (1. - mask)
    tp, sp = torch.sigmoid(mixed), torch.sigmoid(student)
    conf = (tp.max(dim=1).values > thr).to(tp)
    loss = (torch.pow(sp - tp, exponent=2.0).sum(dim=1) * conf).sum() / conf.sum()
    return loss


def timed(fn, x, reps=20, warm=3):
    def once():
        x.grad = None
        fn().backward()
    for _ in range(warm):
        once()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        once()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(1)
    out = {}
    thr = 0.6
    for name, (n, c, h, w, s) in {"16x2x512x512 stride 4": (16, 2, 512, 512, 4), "8x19x512x1024 stride 4": (8, 19, 512, 1024, 4),
                                  "16x2x512x512 full-res teacher": (16, 2, 512, 512, 1)}.items():
        x = (torch.randn(n, c, h, w, device=dev, generator=gen) * 2).requires_grad_(True)
        ta = torch.randn(n, c, h // s, w // s, device=dev, generator=gen) * 3
        tb = torch.randn(n, c, h // s, w // s, device=dev, generator=gen) * 3
        mask = (F.avg_pool2d(torch.randn(n, 1, h, w, device=dev, generator=gen), 9, 1, 4) > 0).float()
        img = torch.zeros(n, 1, h, w, device=dev)

        def three():
            _, mixed = b200ssl.cowmix.mix2_with_mask(img, img, ta, tb, mask)
            return b200ssl.consistency.confidence_masked_consistency(x, mixed, thr)[0]

        def fused():
            return b200ssl.consistency.confidence_masked_consistency_mixed(x, ta, tb, mask, thr)[0]

        l3 = three(); x.grad = None; l3.backward(); g3 = x.grad.clone()
        lf = fused(); x.grad = None; lf.backward(); gf = x.grad.clone()
        rec = {"ms_aten": round(timed(lambda: aten(x, ta, tb, mask, thr), x, reps=10), 4),
               "ms_three_calls": round(timed(three, x), 4), "ms_fused": round(timed(fused, x), 4),
               "grad_bit_identical": bool(torch.equal(g3, gf)), "loss_rel_diff": abs(float(l3.detach()) - float(lf.detach())) / abs(float(l3.detach()))}
        rec["fused_vs_three_calls"] = round(rec["ms_three_calls"] / rec["ms_fused"], 2)
        rec["fused_vs_aten"] = round(rec["ms_aten"] / rec["ms_fused"], 1)
        px = n * h * w
        rec["fused_alg_GBps"] = round(px * ((4 * c + 4 + 8 * c / s / s) + (8 * c + 4 + 16 * c / s / s)) / rec["ms_fused"] / 1e6, 1)
        out[name] = rec
    print(json.dumps(out))


if __name__ == "__main__":
    main()
