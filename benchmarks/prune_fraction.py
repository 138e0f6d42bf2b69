"""How many Lovasz sort keys would an EXACT zero-delta-tail pruning remove?  

Behind the last foreground element of a class's sorted order the intersection is 0, every Jaccard delta is
exactly 0, and those keys need no sorting.  They are the background pixels whose error p_c is strictly below
e_min = min over the class's foreground pixels of (1 - p_c).  This script measures the surviving fraction
sum_c |{fg} U {bg: p_c >= e_min}| / (C * P) on the bench's synthetic inputs and on inputs whose logits agree
with the labels (what a trained network produces), with plain torch ops on the GPU (diagnostic, not product)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def survivors(probas, labels):
    c = probas.shape[1]
    keep = total = 0
    for k in range(c):
        fg = labels == k
        if not bool(fg.any()):
            continue
        pk = probas[:, k]
        e_min = (1.0 - pk[fg]).min()
        keep += int((fg | (pk >= e_min)).sum())
        total += fg.numel()
    return keep / max(total, 1)


def main():
    dev = torch.device("cuda:0")
    out = {}
    for c in ("2", "3"):
        cfg = bench.CONFIGS[c]
        inp = bench.make_inputs(dev, 0, cfg=cfg)
        rec = {"bench_inputs (logits independent of the labels)": round(survivors(inp["scores"], inp["target"]), 4)}
        gen = torch.Generator(device=dev).manual_seed(5)
        logits = torch.randn(inp["scores"].shape, device=dev, generator=gen) * 3
        onehot = torch.nn.functional.one_hot(inp["target"], cfg["c"]).permute(0, 3, 1, 2).float()
        for boost in (3.0, 6.0, 12.0):
            p = torch.softmax(logits + boost * onehot, 1)
            acc = float((p.argmax(1) == inp["target"]).float().mean())
            rec[f"label logit +{boost:g} (pixel accuracy {acc:.2f})"] = round(survivors(p, inp["target"]), 4)
        out[cfg["key"]] = rec
        del inp, logits, onehot
        torch.cuda.empty_cache()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
