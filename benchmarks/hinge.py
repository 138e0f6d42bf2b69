"""lovasz.lovasz_hinge forward + backward (reference lovasz.py:79-111): this repo against the UNMODIFIED reference
function (oracle/_ref, staged by oracle/stage_ref.py) on the same GPU and on the host cores.
python benchmarks/hinge.py > gpurun_out/hinge.json"""
import importlib
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
b200ssl = importlib.import_module("semi-supervised_semantic_segmentation_b200")


def fwd_bwd(fn, x, lab, **kw):
    x.grad = None
    loss = fn(x, lab, **kw)
    loss.backward()
    return loss


def time_cuda(fn, reps=20, warm=3):
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
    from oracle import ref_step
    ref = ref_step.modules()["lovasz"]
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(3)
    out = {}
    for name, (b, h, w) in {"16x512x512": (16, 512, 512), "8x1024x2048": (8, 1024, 2048)}.items():
        lab = (torch.nn.functional.interpolate(torch.randn(b, 1, h // 32, w // 32, device=dev, generator=gen), size=(h, w),
                                               mode="bilinear")[:, 0] > 0.3).long()
        x = ((2.0 * lab.float() - 1.0) * 1.0 + torch.randn(b, h, w, device=dev, generator=gen) * 2.0).requires_grad_(True)
        lab[torch.rand(b, h // 16, w // 16, device=dev, generator=gen).repeat_interleave(16, 1).repeat_interleave(16, 2) < 0.03] = 255
        lab8 = lab.to(torch.uint8)
        rec = {"sorted_fraction": round(float(((1.0 - x.detach() * (2.0 * (lab == 1).float() - 1.0)) > 0)[lab != 255].float().mean()), 3)}
        for per_image in (True, False):
            ours = lambda: fwd_bwd(b200ssl.lovasz.lovasz_hinge, x, lab8, per_image=per_image, ignore=255)
            theirs = lambda: fwd_bwd(ref.lovasz_hinge, x, lab, per_image=per_image, ignore=255)
            l_o = float(ours().detach()); g_o = x.grad.clone()
            l_r = float(theirs().detach()); g_r = x.grad.clone()
            t_o, t_r = time_cuda(ours), time_cuda(theirs, reps=5, warm=1)
            rec["per_image" if per_image else "batch"] = {
                "ms": round(t_o, 4), "reference_aten_cuda_ms": round(t_r, 3), "speedup": round(t_r / t_o, 1),
                "Mpixels/s": round(b * h * w / t_o / 1e3, 1), "loss_rel_diff": abs(l_o - l_r) / max(1.0, abs(l_r)),
                "grad_max_abs_diff_vs_reference_cuda": float((g_o - g_r).abs().max())}
        if name == "16x512x512":      # the reference on the host cores, one call
            xc, lc = x.detach().cpu().requires_grad_(True), lab.cpu()
            t0 = time.perf_counter()
            fwd_bwd(ref.lovasz_hinge, xc, lc, per_image=True, ignore=255)
            rec["reference_cpu_ms_per_image_mode"] = round((time.perf_counter() - t0) * 1e3, 1)
            rec["cpu_threads"] = torch.get_num_threads()
        out[name] = rec
    print(json.dumps(out))


if __name__ == "__main__":
    main()
