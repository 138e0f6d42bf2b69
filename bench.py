#!/usr/bin/env python
"""bench.py -- loss-path throughput (CowMix mask + fused mix + Lovasz fwd/bwd + EMA + confusion matrix).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One "step" is one pass of the semi-supervised loss path over one synthetic batch per GPU
(BASELINE.json configs[1]: 16 x 512 x 512, 2 classes, unet+mobilenetv2 parameter set).  Rank 0 prints
ONE JSON line.  See DESIGN.md "Measurement" for the definition of every key.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "loss-path Mpixels/s (CowMix+Lovasz fwd/bwd+EMA+cm)"
UNIT = "Mpixels/s"
WORKLOAD = dict(name="configs[1]: unet 2-class, batch 16x512x512 per GPU", n=16, c=2, h=512, w=512,
                params="unet_mnv2_c2", p_range=(0.45, 0.55), sigma_range=(8, 32), alpha=0.99)
REPEATS = 3                 # timed K-step regions per run; the median one is reported
REF_SAMPLE_IMAGES = 4       # --impl reference: images per step (bounded sample of the 16-image batch)


def load_param_shapes(key):
    with open(os.path.join(ROOT, "tests", "golden", "param_shapes.json")) as f:
        return json.load(f)[key]["shapes"]


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def coherent_labels(n, c, h, w, device, gen):
    """Spatially coherent label maps: argmax over C channels of blurred noise."""
    x = torch.randn(n, c, h, w, device=device, generator=gen)
    for _ in range(3):
        x = torch.nn.functional.avg_pool2d(x, 17, 1, 8)
    return x.argmax(1)


def make_inputs(device, rank, n=None):
    W = WORKLOAD
    n = W["n"] if n is None else n
    gen = torch.Generator(device=device).manual_seed(1234 + rank)
    c, h, w = W["c"], W["h"], W["w"]
    d = {
        "image_a": torch.rand(n, 3, h, w, device=device, generator=gen),
        "image_b": torch.rand(n, 3, h, w, device=device, generator=gen),
        "teacher_a": torch.randn(n, c, h, w, device=device, generator=gen) * 3,
        "teacher_b": torch.randn(n, c, h, w, device=device, generator=gen) * 3,
        "scores": torch.randn(n, c, h, w, device=device, generator=gen) * 3,
    }
    labels = coherent_labels(n, c, h, w, device, gen)
    d["target"] = torch.nn.functional.one_hot(labels, c).permute(0, 3, 1, 2).float().contiguous()
    shapes = load_param_shapes(W["params"])
    d["params"] = [torch.randn(s, device=device, generator=gen) for s in shapes]
    d["ema_params"] = [torch.randn(s, device=device, generator=gen) for s in shapes]
    return d


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region, sampled in-process through NVML every
    20 ms (the B200_PROFILING.md clocks line; a resident `nvidia-smi -lms` process was measured to
    stall kernel launches for tens of milliseconds per poll, which is comparable to the whole timed
    region here, so the same counters are read with nvidia_ml_py instead)."""

    def __init__(self, index):
        self.index = index
        self.thread = None
        self.stop_flag = False
        self.sm, self.reasons, self.max_mhz = [], set(), None

    def start(self):
        if os.environ.get("B200SSL_BENCH_NO_CLOCKS"):
            return
        try:
            import threading
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self._physical_index(nv))
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            masks = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown,
                     "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                     "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown,
                     "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}

            def loop():
                while not self.stop_flag:
                    try:
                        self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                        r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                        for k, m in masks.items():
                            if r & m:
                                self.reasons.add(k)
                    except Exception:
                        pass
                    time.sleep(0.02)

            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
        except Exception:
            self.thread = None

    def _physical_index(self, nv):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [x.strip() for x in vis.split(",") if x.strip()]
            if self.index < len(ids) and ids[self.index].isdigit():
                return int(ids[self.index])
        return self.index

    def mark(self):
        """samples taken from here on belong to the timed region"""
        self.t_mark = len(self.sm)

    def stop(self):
        if self.thread is None:
            return None
        self.stop_flag = True
        self.thread.join(timeout=2)
        sm = self.sm[getattr(self, "t_mark", 0):] or self.sm
        if not sm:
            return None
        s_sorted = sorted(sm)
        return {"sm_mhz": s_sorted[len(s_sorted) // 2], "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(sm), "how": "NVML, 20 ms period, timed region only"}


def algorithmic_bytes(P, C, n_params, label_bytes=1):
    """Compulsory bytes per launch of each kernel for P pixels (DESIGN.md 'Kernels'); binary mode:
    one Lovasz segment per image, P keys in total."""
    elems = P * C
    return {
        "cowmix_conv_pass1": 8 * P, "cowmix_conv_pass2": 8 * P, "cowmix_threshold": 8 * P,
        "mix2": 4 * (3 * 3 + 3 * C + 1) * P,
        "mix2_threshold": 4 * (3 * 3 + 3 * C + 2) * P,
        "argmax_channels": (4 * C + label_bytes) * P,
        "lovasz_keybuild": (4 + label_bytes + 8) * P,
        "lovasz_binary_prep": (8 * C + label_bytes + 8) * P,
        "lovasz_sort_pass0": 16 * P, "lovasz_sort_pass1": 16 * P, "lovasz_sort_pass2": 16 * P,
        "lovasz_rank_grad_pass3": 12 * P,
        "lovasz_backward": 8 * elems,
        "ema_multi": 12 * n_params,
        "confusion_from_logits": (4 * C + label_bytes) * P,
    }


def step_algorithmic_bytes(P, C, n_params):
    """P*(72+20C) + 12*n_params (mask 8, mix 40+12C, Lovasz 8C+8, CM 16 per pixel)."""
    return P * (72 + 20 * C) + 12 * n_params


# ------------------------------------------------------------------------------------------------
def run_b200(args, rank, world, local_rank):
    import b200ssl
    from b200ssl import _lib
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    W = WORKLOAD
    inp = make_inputs(device, rank)
    P = W["n"] * W["h"] * W["w"]
    n_params = sum(p.numel() for p in inp["params"])
    # Multi-GPU: the batch shards by image, the only exchange is [confusion matrix || loss]
    # once per step.  It goes over NVLink peer memory (b200ssl_peer_*: posted by the step itself, collected
    # on the communicator's stream, no NCCL call on the data path); if the box cannot map peer memory
    # every rank falls back to ONE torch.distributed all-reduce per step (the choice is collective).
    peer = reducer = None
    if world > 1:
        peer = b200ssl.utils.make_peer_all_reduce(W["c"] * W["c"], 1, device)
        if peer is None:
            reducer = b200ssl.utils.StepReducer(W["c"], 1, device, backend="dist")
    step = b200ssl.LossPathStep(num_classes=W["c"], mask_proportion_range=W["p_range"],
                                sigma_range=W["sigma_range"], ema_alpha=W["alpha"], mode="binary", peer=peer)
    step.bind_parameters(inp["params"], inp["ema_params"])     # like constructing an optimizer over the lists
    torch.manual_seed(0)            # the reference seeds every rank with 0 (distributed_trainer.py:17)

    def one_step():
        out = step(inp["image_a"], inp["image_b"], inp["teacher_a"], inp["teacher_b"], inp["scores"],
                   inp["target"], inp["params"], inp["ema_params"])
        if reducer is not None:
            reducer.all_reduce(out["cm"], [out["loss"]])
        return out

    def sync_all():
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(device)

    # The clock sampler is started BEFORE the warm-up, and the warm-up runs for at least ~1.2 s of GPU
    # load so that clocks and power state have settled when the timed region starts.
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    # Every warm-up step of a multi-rank run contains a collective, so all ranks must run the SAME
    # number of steps: the "long enough" decision is taken in chunks and agreed on by all ranks (a
    # purely time-based loop per rank desynchronises the collective sequence and hangs the job).
    t_w = time.perf_counter()
    n_w = 0
    chunk = max(args.warmup, 3)
    while True:
        for _ in range(chunk):
            one_step()      # never synchronised: the caching allocator must reach its run-ahead steady state
        n_w += chunk
        done = (time.perf_counter() - t_w) >= 1.2
        if world > 1:
            flag = torch.tensor([1 if done else 0], device=device, dtype=torch.int32)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            done = bool(int(flag))
        if done:
            break
        chunk = 200
    sync_all()

    # ---- timed region: K steps, device-resident inputs ----
    # The K-step region is timed REPEATS times back to back and the median region is reported (all
    # of them are listed in "ms_per_step_runs"): one region lasts only ~30 ms, and the first region
    # after the warm-up regularly contains ONE host-side stall of 4-300 ms inside torch.empty_like
    # (the caching allocator re-establishing its pool after the synchronisation; found with
    # B200SSL_BENCH_DEBUG=1), which otherwise decides the number.
    clocks.mark()
    region_ms = []
    launches = 0
    for _rep in range(REPEATS):
        l0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        e0.record()
        t_host = []
        for _i in range(args.steps):
            t_a = time.perf_counter()
            out = one_step()
            t_host.append(time.perf_counter() - t_a)
        e1.record()
        if os.environ.get("B200SSL_BENCH_DEBUG"):
            worst = max(t_host)
            print(f"[debug] region {_rep}: worst host step {worst * 1e3:.2f} ms at {t_host.index(worst)}, "
                  f"host sum {sum(t_host) * 1e3:.1f} ms", file=sys.stderr)
        sync_all()
        launches = _lib.launch_count() - l0
        t = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)      # max over ranks, per region
        region_ms.append(float(t))
    ms = torch.tensor([sorted(region_ms)[len(region_ms) // 2]], device=device, dtype=torch.float64)
    clock_info = clocks.stop() if rank == 0 else None

    # ---- per-kernel CUDA-event timing over K more steps (explains the number above) ----
    # The timed step overlaps its three independent chains on internal streams, which stretches every
    # kernel's own duration; the per-kernel figures are therefore taken with the same kernels issued
    # one after the other (serial=True), i.e. each kernel alone on the GPU.
    prof_steps = min(args.steps, 20)
    step_serial = b200ssl.LossPathStep(num_classes=W["c"], mask_proportion_range=W["p_range"],
                                       sigma_range=W["sigma_range"], ema_alpha=W["alpha"], mode="binary",
                                       serial=True)

    def serial_step():
        return step_serial(inp["image_a"], inp["image_b"], inp["teacher_a"], inp["teacher_b"], inp["scores"],
                           inp["target"], inp["params"], inp["ema_params"])

    for _ in range(3):
        serial_step()
    torch.cuda.synchronize(device)
    _lib.kernel_times(True)
    for _ in range(prof_steps):
        serial_step()
    torch.cuda.synchronize(device)
    ktimes = _lib.kernel_times()
    _lib.kernel_times(False)

    # ---- end to end: pinned host inputs -> H2D -> step -> D2H of loss + confusion matrix ----
    # Every step copies its own 234.9 MB of inputs from pinned host memory and reads its loss and
    # confusion matrix back.  The copies run on a second stream into two alternating device buffer
    # sets, so the upload of step i+1 overlaps the kernels of step i (the host still waits for each
    # step's result before it issues the next step, as a training loop reading the loss would).
    names = ["image_a", "image_b", "teacher_a", "teacher_b", "scores", "target"]
    host = {k: inp[k].cpu().pin_memory() for k in names}
    dev_in = [{k: torch.empty_like(inp[k]) for k in names} for _ in range(2)]
    h2d_bytes = sum(host[k].numel() * host[k].element_size() for k in names)
    res_host = [torch.empty(1 + W["c"] * W["c"], dtype=torch.float64).pin_memory() for _ in range(2)]
    copy_stream = torch.cuda.Stream(device)
    main_stream = torch.cuda.current_stream(device)
    copied = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def upload(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])       # the step that last read this buffer set is done
            for k in names:
                dev_in[slot][k].copy_(host[k], non_blocking=True)
            copied[slot].record(copy_stream)

    def e2e_run(k_steps):
        for sl in range(2):
            consumed[sl].record(main_stream)
        upload(0)
        for i in range(k_steps):
            sl = i & 1
            if i + 1 < k_steps:
                upload(sl ^ 1)
            main_stream.wait_event(copied[sl])
            b = dev_in[sl]
            o = step(b["image_a"], b["image_b"], b["teacher_a"], b["teacher_b"], b["scores"], b["target"],
                     inp["params"], inp["ema_params"])
            consumed[sl].record(main_stream)
            if peer is not None:
                peer.result()                            # main stream waits for this step's collect
                packed = torch.cat([o["loss_sum"].reshape(1), o["cm_sum"].reshape(-1).to(torch.float64)])
            elif reducer is not None:
                cm, scs = reducer.all_reduce(o["cm"], [o["loss"]])
                packed = torch.cat([scs.reshape(-1)[:1], cm.reshape(-1).to(torch.float64)])
            else:
                packed = torch.cat([o["loss"].reshape(1).to(torch.float64), o["cm"].reshape(-1).to(torch.float64)])
            res_host[sl].copy_(packed, non_blocking=True)
            main_stream.synchronize()                    # the caller reads the loss every step

    e2e_run(3)
    sync_all()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    e2e_run(args.steps)
    e3.record()
    sync_all()
    ms_e2e = torch.tensor([e2.elapsed_time(e3)], device=device, dtype=torch.float64)

    if world > 1:
        dist.all_reduce(ms_e2e, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(ms), float(ms_e2e)
    if peer is not None:
        # a wait that timed out anywhere invalidates the run on every rank (checked collectively)
        try:
            peer.status()
            bad = 0
        except Exception as e:   # noqa: BLE001
            print(f"[rank {rank}] {e}", file=sys.stderr)
            bad = 1
        flag = torch.tensor([bad], device=device, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        peer.close()
        if int(flag):
            raise RuntimeError("the NVLink peer exchange timed out on at least one rank; results are invalid")
    if rank != 0:
        return None

    peak, peak_src = measured_peak()
    value = world * P * args.steps / (ms * 1e-3) / 1e6
    e2e_value = world * P * args.steps / (ms_e2e * 1e-3) / 1e6
    abytes = algorithmic_bytes(P, W["c"], n_params)
    stages = {}
    for name, (cnt, total_ms, min_ms) in ktimes.items():
        avg = total_ms / max(cnt, 1)
        b = abytes.get(name)
        stages[name] = {"launches_per_step": cnt / prof_steps, "avg_ms": round(avg, 5),
                        "alg_GB": None if b is None else round(b / 1e9, 5),
                        "GBps": None if b is None or avg <= 0 else round(b / 1e9 / (avg * 1e-3), 1)}
    per_step_kernel_ms = sum(t for (_, t, _) in ktimes.values()) / prof_steps
    top = max((k for k in ktimes if k in abytes), key=lambda k: ktimes[k][1])
    top_avg_ms = ktimes[top][1] / ktimes[top][0]
    achieved = abytes[top] / 1e9 / (top_avg_ms * 1e-3)
    # DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of this
    # workload (profiles/traffic.json: kernel -> {bytes, source}); not measured live (never under a profiler)
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get(top, {}).get("dram_bytes_per_launch")
    except (OSError, ValueError):
        pass
    roof = {"kernel": top, "bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
            "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
            "avg_launch_ms": round(top_avg_ms, 5),
            "share_of_kernel_time": round(ktimes[top][1] / prof_steps / per_step_kernel_ms, 4),
            "stages": stages,
            "step": {"alg_GB": round(step_algorithmic_bytes(P, W["c"], n_params) / 1e9, 4),
                     "GBps": round(step_algorithmic_bytes(P, W["c"], n_params) / 1e9 / (ms / args.steps * 1e-3), 1),
                     "frac": round(step_algorithmic_bytes(P, W["c"], n_params) / 1e9 / (ms / args.steps * 1e-3) / peak, 4),
                     "kernel_ms_per_step_serial": round(per_step_kernel_ms, 4),
                     "note": "stages: each kernel alone (serial issue); the timed step overlaps the mask+mix, "
                             "Lovasz and EMA chains on internal streams"}}
    if top.startswith("cowmix_conv"):
        roof["note"] = ("this kernel is fp32-FMA bound (2K FMA per pixel, K up to 193), not HBM bound; "
                        "see roofline.stages and DESIGN.md for its FMA-rate fraction")
    line = {
        "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "warmup_steps_run": n_w, "ms_per_step": round(ms / args.steps, 4),
        "ms_per_step_runs": [round(x / args.steps, 4) for x in region_ms], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": W["name"], "pixels_per_gpu_step": P, "classes": W["c"],
                   "ema_params": n_params, "ema_tensors": len(inp["params"]),
                   "lovasz": "losses.binary_lovasz_loss_with_logits (per image, class 1)",
                   "sigma_range": list(W["sigma_range"]), "parallelism": f"dp{world}",
                   "collective": ("none (single GPU)" if world == 1 else
                                  "one-shot exchange of [cm || loss] over NVLink peer memory per step "
                                  "(b200ssl_peer_*), collected lazily" if peer is not None else
                                  "one torch.distributed (NCCL) all_reduce of [cm || loss] per step"),
                   "l2": "no flush: the 234 MB of step inputs (+268 MB outputs/workspace) exceed the 126 MB L2"},
        "clocks": clock_info,
        "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": res_host[0].numel() * 8, "ms_per_step": round(ms_e2e / args.steps, 4)},
        "gpu_launches": int(launches),
        "roofline": roof,
    }
    if world == 1:
        line["cpu_baseline"] = cpu_baseline(inp, full=True, steps=2, warmup=1)
    return line


# ------------------------------------------------------------------------------------------------
def cpu_inputs(inp, n_images):
    keys = ["image_a", "image_b", "teacher_a", "teacher_b", "scores", "target"]
    d = {k: inp[k][:n_images].cpu() for k in keys}
    d["params"] = [p.cpu() for p in inp["params"]]
    d["ema_params"] = [p.cpu().clone() for p in inp["ema_params"]]
    return d


def cpu_step(d):
    from oracle import torch_port
    W = WORKLOAD
    return torch_port.loss_path_step(d["image_a"], d["image_b"], d["teacher_a"], d["teacher_b"], d["scores"],
                                     d["target"], d["params"], d["ema_params"], mode="binary",
                                     mask_proportion_range=W["p_range"], sigma_range=W["sigma_range"],
                                     alpha=W["alpha"], num_classes=W["c"])


def cpu_baseline(inp, full, steps, warmup, n_images=None):
    """The oracle's torch port (the reference's ATen op sequence) on this host's cores."""
    W = WORKLOAD
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_img = W["n"] if full else (n_images or REF_SAMPLE_IMAGES)
    d = cpu_inputs(inp, n_img)
    torch.manual_seed(0)
    for _ in range(warmup):
        cpu_step(d)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_step(d)
    dt = (time.perf_counter() - t0) / steps
    pix = n_img * W["h"] * W["w"]
    return {"value": round(pix / dt / 1e6, 3), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n_img} of {W['n']} images x {W['h']}x{W['w']} per step, full EMA parameter set, "
                      f"{steps} timed steps after {warmup} warm-up; oracle/torch_port.py (torch CPU ops, "
                      f"{torch.get_num_threads()} threads)",
            "ms_per_step": round(dt * 1e3, 2)}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (oracle port: the reference is
    Python and /root/reference does not exist on the GPU box), rank 0 only."""
    if rank != 0:
        return None
    W = WORKLOAD
    inp = make_inputs(torch.device("cpu"), 0, n=REF_SAMPLE_IMAGES)
    # bounded sample: shrink the per-step image count so that (K+W) steps end within ~3 minutes
    torch.set_num_threads(os.cpu_count() or 1)
    probe = cpu_inputs(inp, 1)
    cpu_step(probe)
    t0 = time.perf_counter()
    cpu_step(probe)
    t_img = time.perf_counter() - t0
    total_steps = max(args.steps, 1) + max(args.warmup, 1)
    n_img = int(max(1, min(REF_SAMPLE_IMAGES, 180.0 / (total_steps * t_img))))
    base = cpu_baseline(inp, full=False, steps=max(args.steps, 1), warmup=max(args.warmup, 1), n_images=n_img)
    return {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": base["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": W["name"], "classes": W["c"], "sigma_range": list(W["sigma_range"]),
                   "sample_images_per_step": n_img},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        line = run_reference(args, rank, world)
        if line is not None:
            print(json.dumps(line), flush=True)
        return 0
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        line = run_b200(args, rank, world, local_rank)
        if line is not None:
            print(json.dumps(line), flush=True)
    finally:
        if world > 1 and dist.is_initialized():
            dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
