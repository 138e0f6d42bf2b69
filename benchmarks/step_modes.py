"""configs[1] step in its three issue modes (eager / static outputs / CUDA graph): device time per step
(CUDA events over `steps` steps) and host time per step (perf_counter around the issuing loop)."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import b200ssl  # noqa: E402


def main(steps=300):
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    inp = bench.make_inputs(dev, 0)
    W = bench.WORKLOAD
    out = {}
    for name, kw in (("eager", {}), ("static", dict(static_outputs=True)), ("graph", dict(graph=True))):
        step = b200ssl.LossPathStep(num_classes=W["c"], mask_proportion_range=W["p_range"], sigma_range=W["sigma_range"],
                                    ema_alpha=W["alpha"], mode="binary", **kw)
        step.bind_parameters(inp["params"], inp["ema_params"])
        torch.manual_seed(0)

        def one():
            return step(inp["image_a"], inp["image_b"], inp["teacher_a"], inp["teacher_b"], inp["scores"], inp["target"],
                        inp["params"], inp["ema_params"])
        for _ in range(1500):
            one()
        torch.cuda.synchronize()
        res = []
        for _rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e0.record()
            for _ in range(steps):
                one()
            e1.record()
            t_host = time.perf_counter() - t0
            torch.cuda.synchronize()
            res.append((e0.elapsed_time(e1) / steps, t_host / steps * 1e3))
        out[name] = {"ms_per_step": [round(r[0], 4) for r in res], "host_ms_per_step": [round(r[1], 4) for r in res],
                     "graphs": step.graph_captures}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
