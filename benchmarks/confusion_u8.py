"""uint8 confusion matrix: GB/s of b200ssl_confusion_matrix on spatially coherent masks of several boundary
densities (cell = side of the low-resolution noise cell the label regions are grown from), against the int64
kernel on the same pixels.  python benchmarks/confusion_u8.py > gpurun_out/confusion_u8.json"""
import importlib
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
b200ssl = importlib.import_module("semi-supervised_semantic_segmentation_b200")


def masks(n, c, h, w, cell, gen, dev):
    x = torch.randn(n, c, max(h // cell, 1), max(w // cell, 1), device=dev, generator=gen)
    return torch.nn.functional.interpolate(x, size=(h, w), mode="bilinear", align_corners=False).argmax(1)


def main():
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(7)
    c, n, h, w = 19, 64, 1024, 2048
    out = {"pixels_per_call": n * h * w, "classes": c, "unroll": os.environ.get("B200SSL_CM8_UNROLL", "2")}
    for cell in (8, 32, 128):
        pools = []
        for _ in range(3):
            lab = torch.cat([masks(16, c, h, w, cell, gen, dev) for _ in range(n // 16)])
            prd = torch.cat([masks(16, c, h, w, cell, gen, dev) for _ in range(n // 16)])
            wrong = (torch.rand(n, h // 16, w // 16, device=dev, generator=gen) < 0.2).repeat_interleave(16, 1).repeat_interleave(16, 2)
            prd = torch.where(wrong, prd, lab)
            lab[(torch.rand(n, h // 16, w // 16, device=dev, generator=gen) < 0.02).repeat_interleave(16, 1).repeat_interleave(16, 2)] = 255
            pools.append((lab.to(torch.uint8), prd.to(torch.uint8)))
            del lab, prd, wrong
        l0, p0 = pools[0]
        key = l0.reshape(-1).to(torch.int32) | (p0.reshape(-1).to(torch.int32) << 8)
        runs_per_512 = float((key[1:] != key[:-1]).sum()) / key.numel() * 512
        del key
        rec = {"runs_per_512_pixels": round(runs_per_512, 2)}
        for tag in ("uint8", "int64"):
            data = pools if tag == "uint8" else [(a.long(), b.long()) for a, b in pools[:2]]
            cm = torch.zeros(c, c, dtype=torch.int64, device=dev)
            for a, b in data:
                b200ssl.metrics.confusion_matrix(a, b, c, ignore_index=255, out=cm)
            torch.cuda.synchronize()
            reps = 30
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            ev[0].record()
            for i in range(reps):
                a, b = data[i % len(data)]
                b200ssl.metrics.confusion_matrix(a, b, c, ignore_index=255, out=cm)
            ev[1].record()
            torch.cuda.synchronize()
            ms = ev[0].elapsed_time(ev[1]) / reps
            bpp = 2 if tag == "uint8" else 16
            rec[tag] = {"ms": round(ms, 4), "GB/s": round(n * h * w * bpp / ms / 1e6, 1), "Gpix/s": round(n * h * w / ms / 1e6, 1)}
            if tag == "uint8":
                want = torch.zeros(c * c, dtype=torch.int64, device=dev)
                a, b = data[0]
                keep = a.reshape(-1) != 255
                want += torch.bincount(a.reshape(-1)[keep].long() * c + b.reshape(-1)[keep].long(), minlength=c * c)
                got = b200ssl.metrics.confusion_matrix(a, b, c, ignore_index=255)
                rec["exact"] = bool(torch.equal(got.reshape(-1), want))
            del data
        out[f"cell_{cell}"] = rec
        del pools
    print(json.dumps(out))


if __name__ == "__main__":
    main()
