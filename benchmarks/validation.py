"""One-pass validation Dice (train.py:171-175) against the reference's own ATen sequence on the same GPU:
argmax -> one_hot -> permute -> `nearest` interpolate -> mask > 0.5 -> metrics.dice_metric.
python benchmarks/validation.py > gpurun_out/validation.json"""
import importlib
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
b200ssl = importlib.import_module("semi-supervised_semantic_segmentation_b200")


def aten_sequence(pred, mask):
    """train.py:171-175 + metrics.py:1-7, restated with the same ATen calls"""
    one_hot = F.one_hot(torch.argmax(pred, dim=1), pred.shape[1]).permute(0, 3, 1, 2).float()
    binary = F.interpolate(one_hot, size=mask.shape[2:], mode="nearest")
    x, y = binary[:, 1:], (mask > 0.5)[:, 1:].float()
    inter = (x * y).sum(dim=(1, 2, 3))
    card = (x + y).sum(dim=(1, 2, 3))
    return (2.0 * inter + 1.0) / (card + 1.0)


def time_ms(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(11)
    out = {"kernel": "v1" if os.environ.get("B200SSL_VALCM_V1") else "rows"}
    for name, (n, c, h, w, H, W) in {"train.py shapes 16x2x128x128 -> 512x512": (16, 2, 128, 128, 512, 512),
                                     "64x2x128x128 -> 512x512": (64, 2, 128, 128, 512, 512),
                                     "8x2x256x512 -> 1024x2048": (8, 2, 256, 512, 1024, 2048),
                                     "ragged 5x2x100x75 -> 401x301": (5, 2, 100, 75, 401, 301)}.items():
        pool = []
        for _ in range(6):                      # rotate over > L2 worth of masks for the large shapes
            pred = torch.randn(n, c, h, w, device=dev, generator=gen)
            soft = F.interpolate(torch.randn(n, 1, H // 16 + 1, W // 16 + 1, device=dev, generator=gen), size=(H, W),
                                 mode="bilinear").sigmoid()
            pool.append((pred, torch.cat([1 - soft, soft], 1).contiguous()))
        i = [0]

        def ours():
            p, m = pool[i[0] % len(pool)]
            i[0] += 1
            return b200ssl.metrics.validation_dice(p, m)[0]

        def ref():
            p, m = pool[i[0] % len(pool)]
            i[0] += 1
            return aten_sequence(p, m)

        same = all(torch.equal(b200ssl.metrics.validation_dice(p, m)[0], aten_sequence(p, m)) for p, m in pool[:2])
        t_ours, t_ref = time_ms(ours), time_ms(ref)
        px = n * H * W
        out[name] = {"ms": round(t_ours, 4), "aten_ms": round(t_ref, 4), "speedup": round(t_ref / t_ours, 1),
                     "GB/s_mask_plane": round(px * 4 / t_ours / 1e6, 1), "bit_identical_to_aten": bool(same)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
