#!/usr/bin/env python
"""A short run of the bench's configs[N] step for ncu captures (the bench itself warms up for thousands of steps):
    python profiles/capture_step.py [config=1] [steps=8]
Kernels per step at configs[1]: 12 (3 memsets are not kernels).  Typical capture, after the plain run exited 0:
    ncu --set full --clock-control none --import-source on -s 60 -c 12 -o gpurun_out/prof python profiles/capture_step.py 1 8
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import b200ssl  # noqa: E402


def main():
    cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "1"]
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    dev = torch.device("cuda:0")
    inp = bench.make_inputs(dev, 0, cfg=cfg)
    step = bench.make_step(b200ssl, cfg, None, static_outputs=True, ring=cfg["ring"])
    step.bind_parameters(inp["params"], inp["ema_params"])
    torch.manual_seed(0)
    torch.cuda.synchronize()
    for _ in range(steps):
        out = step(inp["image_a"], inp["image_b"], inp["teacher_a"], inp["teacher_b"], inp["scores"], inp["target"],
                   inp["params"], inp["ema_params"])
    torch.cuda.synchronize()
    print("loss", float(out["loss"]), "launches", b200ssl._lib.launch_count())


if __name__ == "__main__":
    main()
