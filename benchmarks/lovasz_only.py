"""Per-kernel times of the Lovasz chain alone at the bench shapes (serial issue, CUDA events inside the library).
    python benchmarks/lovasz_only.py [1,2,3]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import b200ssl  # noqa: E402
from b200ssl import _lib  # noqa: E402


def main():
    which = (sys.argv[1] if len(sys.argv) > 1 else "1,2,3").split(",")
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else None      # soft-max mode: probabilities of logits * scale
    dev = torch.device("cuda:0")
    out = {}
    for c in which:
        cfg = bench.CONFIGS[c]
        inp = bench.make_inputs(dev, 0, cfg=cfg)
        if scale is not None and cfg["mode"] == "softmax":
            gen = torch.Generator(device=dev).manual_seed(9)
            logits = torch.randn(inp["scores"].shape, device=dev, generator=gen) * abs(scale)
            if scale < 0:      # negative: logits that agree with the labels (a trained network)
                logits += 12.0 * torch.nn.functional.one_hot(inp["target"], cfg["c"]).permute(0, 3, 1, 2)
            inp["scores"] = torch.softmax(logits, 1)
            del logits
        step = bench.make_step(b200ssl, cfg, None, serial=True, static_outputs=True, ring=1)
        for _ in range(3):
            step.lovasz_loss_and_grad(inp["scores"], inp["target"])
        torch.cuda.synchronize()
        _lib.kernel_times(True)
        reps = 10
        for _ in range(reps):
            step.lovasz_loss_and_grad(inp["scores"], inp["target"])
        torch.cuda.synchronize()
        kt = _lib.kernel_times()
        _lib.kernel_times(False)
        keys = cfg["n"] * cfg["h"] * cfg["w"] * (1 if cfg["mode"] == "binary" else cfg["c"])
        rec = {k: round(v[1] / reps * 1e3, 2) for k, v in kt.items()}
        rec["total_us"] = round(sum(rec.values()), 1)
        rec["ps_per_key"] = round(rec["total_us"] * 1e6 / keys, 2)
        out[cfg["key"]] = rec
        del inp, step
        torch.cuda.empty_cache()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
