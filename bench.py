#!/usr/bin/env python
"""bench.py -- loss-path throughput (CowMix mask + fused mix + Lovasz fwd/bwd + EMA + confusion matrix).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--configs 1,2,3,4]

One "step" is one pass of the semi-supervised loss path over one synthetic batch per GPU.  Rank 0 prints ONE
JSON line.  The headline (`value`, `e2e`, `roofline`, `cpu_baseline`, `aten_cuda_baseline`) is BASELINE.json
configs[1] (16 x 512 x 512, 2 classes, unet+mobilenetv2 parameter set); the `configs` object of the same line
carries the other configurations BASELINE.json names, each with its own value / roofline / baselines:
    configs[2]  deeplabv3 21-class, 32 x 512 x 512 across 8 GPUs  -> the per-GPU shard 4 x 21 x 512 x 512,
                lovasz_softmax(classes='present'), deeplabv3-R101 EMA set (replicated on every GPU)
    configs[3]  higher_hrnet 19-class 1024 x 2048, batch 8 per GPU, hrnet EMA set
    configs[4]  mIoU sweep: 19-class confusion matrix over 10 000 masks of 1024 x 2048, split over the ranks
                (strong scaling), reduced with the peer exchange, int64 and uint8 labels
See DESIGN.md "Measurement" for the definition of every key.
"""
import argparse
import json
import math
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "loss-path Mpixels/s (CowMix+Lovasz fwd/bwd+EMA+cm)"
UNIT = "Mpixels/s"
CONFIGS = {
    "1": dict(key="configs[1]", name="configs[1]: unet 2-class, batch 16x512x512 per GPU", n=16, c=2, h=512, w=512,
              params="unet_mnv2_c2", mode="binary", p_range=(0.45, 0.55), sigma_range=(8, 32), alpha=0.99,
              lovasz="losses.binary_lovasz_loss_with_logits (per image, class 1)", ring=2, cpu_images=16),
    "2": dict(key="configs[2]", name="configs[2]: deeplabv3 21-class, batch 32x512x512 across 8 GPUs = 4x512x512 per GPU",
              n=4, c=21, h=512, w=512, params="deeplabv3_r101_c21", mode="softmax", p_range=(0.45, 0.55),
              sigma_range=(8, 32), alpha=0.99, lovasz="lovasz.lovasz_softmax(probas, labels, classes='present')",
              ring=2, cpu_images=1),
    "3": dict(key="configs[3]", name="configs[3]: higher_hrnet 19-class 1024x2048, batch 8 per GPU", n=8, c=19, h=1024,
              w=2048, params="hrnet_small_c19", mode="softmax", p_range=(0.45, 0.55), sigma_range=(8, 32), alpha=0.99,
              lovasz="lovasz.lovasz_softmax(probas, labels, classes='present')", ring=1, cpu_images=1),
}
WORKLOAD = CONFIGS["1"]     # the headline workload (benchmarks/*.py import this name)
MIOU = dict(key="configs[4]", name="configs[4]: mIoU sweep, 19-class confusion matrix over 10000 masks of 1024x2048",
            masks=10000, c=19, h=1024, w=2048, chunk=64, pool=3)
MIN_TIMED_STEPS = 200       # the headline's K-step regions are repeated until at least this many steps are timed


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
    if h * w > 512 * 512:
        # large planes: blur by up-sampling low-resolution noise (same statistics, a fraction of the time)
        x = torch.randn(n, c, h // 32, w // 32, device=device, generator=gen)
        return torch.nn.functional.interpolate(x, size=(h, w), mode="bilinear", align_corners=False).argmax(1)
    x = torch.randn(n, c, h, w, device=device, generator=gen)
    for _ in range(3):
        x = torch.nn.functional.avg_pool2d(x, 17, 1, 8)
    return x.argmax(1)


def make_inputs(device, rank, n=None, cfg=None):
    W = cfg or WORKLOAD
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
    if W["mode"] == "binary":
        d["target"] = torch.nn.functional.one_hot(labels, c).permute(0, 3, 1, 2).float().contiguous()
    else:
        # lovasz_softmax's contract is probabilities (lovasz.py:155-160): formed once, outside the timed region
        d["scores"] = torch.softmax(d["scores"], 1)
        d["target"] = labels.contiguous()
    shapes = load_param_shapes(W["params"])
    d["params"] = [torch.randn(s, device=device, generator=gen) for s in shapes]
    d["ema_params"] = [torch.randn(s, device=device, generator=gen) for s in shapes]
    return d


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region, sampled in-process through NVML every
    10 ms (the B200_PROFILING.md clocks line; a resident `nvidia-smi -lms` process was measured to
    stall kernel launches for tens of milliseconds per poll, so the same counters are read with
    nvidia_ml_py instead).  B200SSL_BENCH_NO_CLOCKS=1 switches it off (stall diagnosis)."""

    def __init__(self, index):
        self.index = index
        self.thread = None
        self.stop_flag = False
        self.active = False
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
                    if self.active:
                        try:
                            self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                            r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                            for k, m in masks.items():
                                if r & m:
                                    self.reasons.add(k)
                        except Exception:
                            pass
                    time.sleep(0.01)

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

    def stop(self):
        if self.thread is None:
            return None
        self.stop_flag = True
        self.thread.join(timeout=2)
        if not self.sm:
            return None
        s_sorted = sorted(self.sm)
        return {"sm_mhz": s_sorted[len(s_sorted) // 2], "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "how": "NVML, 10 ms period, timed regions only"}


# ------------------------------------------------------------------------------------------------
# algorithmic (compulsory) bytes, SURVEY 8(d)
def stage_of(kernel):
    """kernel name (csrc prof_begin labels) -> stage of the path"""
    if kernel.startswith("cowmix") or kernel == "mix2_threshold_stats":
        return "mask"
    if kernel.startswith("mix2"):
        return "mix"
    if kernel.startswith("lovasz") or kernel.startswith("binary_lovasz") or kernel.startswith("softmax"):
        return "lovasz"
    if kernel.startswith("ema") or kernel.startswith("sgd"):
        return "ema"
    if kernel.startswith("confusion") or kernel.startswith("argmax"):
        return "cm"
    if kernel.startswith("peer"):
        return "exchange"
    return "other"


def stage_algorithmic_bytes(P, C, n_params):
    """SURVEY 8(d) per pixel: mask 8 (noise in, mask out), mix 40+12C (both tensors, one mask read),
    Lovasz fwd+bwd 8C+8 (scores in, gradient out, labels), confusion matrix 16 (int64 label + int64
    prediction); EMA 12 B per parameter.  The radix passes' own traffic is NOT algorithmic."""
    return {"mask": 8 * P, "mix": (40 + 12 * C) * P, "lovasz": (8 * C + 8) * P, "cm": 16 * P, "ema": 12 * n_params}


def step_algorithmic_bytes(P, C, n_params):
    """P*(72+20C) + 12*n_params."""
    return P * (72 + 20 * C) + 12 * n_params


# ------------------------------------------------------------------------------------------------
def sync_all(device, world):
    torch.cuda.synchronize(device)
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize(device)


def timed_regions(one_step, steps, n_regions, device, world, end_region=None, clocks=None, debug=None,
                  per_rank_log=None):
    """n_regions back-to-back regions of EXACTLY `steps` steps, each bracketed by barrier + synchronize on both
    sides and timed with CUDA events on the issuing stream; per region the MAX over ranks.  `end_region`
    (multi-GPU: the flush of the last step's exchange) is issued inside the region, before the stop event."""
    out = []
    align = torch.zeros(1, device=device) if world > 1 else None
    # the interpreter's cyclic garbage collector is a property of this harness, not of the path: a generation-2
    # sweep in the middle of a 6 ms region showed up as a single 1.2 ms host stall (one region of ten at 0.345 instead
    # of 0.284 ms/step).  Collect once up front, keep it off inside the timed regions.
    import gc
    gc.collect()
    gc_was_enabled = gc.isenabled()
    gc.disable()
    try:
        return _timed_regions(one_step, steps, n_regions, device, world, end_region, clocks, debug, per_rank_log, align)
    finally:
        if gc_was_enabled:
            gc.enable()


def _timed_regions(one_step, steps, n_regions, device, world, end_region, clocks, debug, per_rank_log, align):
    out = []
    for rep in range(n_regions):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all(device, world)
        if world > 1:
            # The host-side barrier releases the ranks tens of microseconds apart; the steps exchange data every
            # step, so the early rank would absorb that skew INSIDE its timed region (it waits for the late rank's
            # first post).  A device-side rendezvous queued right before the start event aligns the GPU timelines
            # to a few microseconds instead; it is outside the region on every rank.
            dist.all_reduce(align)
        if clocks is not None:
            clocks.active = True
        e0.record()
        t_host = []
        for _ in range(steps):
            t_a = time.perf_counter()
            one_step()
            t_host.append(time.perf_counter() - t_a)
        if end_region is not None:
            end_region()
        e1.record()
        sync_all(device, world)
        if clocks is not None:
            clocks.active = False
        if debug:
            worst = max(t_host)
            print(f"[debug] {debug} region {rep}: worst host step {worst * 1e3:.3f} ms at {t_host.index(worst)}, "
                  f"host sum {sum(t_host) * 1e3:.1f} ms, device {e0.elapsed_time(e1):.1f} ms", file=sys.stderr)
        t = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
        if world > 1:
            per_rank = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(per_rank, t)
            if per_rank_log is not None:
                per_rank_log.append([round(float(x), 4) for x in per_rank])
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out.append(float(t))
    return out


def make_step(b200ssl, cfg, peer, **kw):
    return b200ssl.LossPathStep(num_classes=cfg["c"], mask_proportion_range=cfg["p_range"], sigma_range=cfg["sigma_range"],
                                ema_alpha=cfg["alpha"], mode=cfg["mode"], classes="present", per_image=False, ignore=255,
                                peer=peer, **kw)


def kernel_stage_table(b200ssl, cfg, inp, device, prof_steps, P, n_params, peak):
    """Per-kernel CUDA-event timing with the same kernels issued one after the other (serial=True: each
    kernel alone on the GPU; the timed step overlaps its three chains, which stretches every kernel's own
    duration), grouped into the stages of the path with their SURVEY 8(d) algorithmic bytes."""
    from b200ssl import _lib
    step_serial = make_step(b200ssl, cfg, None, serial=True, static_outputs=True, ring=1)

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
    del step_serial
    kernels, stages = {}, {}
    for name, (cnt, total_ms, _min_ms) in ktimes.items():
        per_step = total_ms / prof_steps
        kernels[name] = {"launches_per_step": round(cnt / prof_steps, 2), "ms_per_step": round(per_step, 5)}
        st = stages.setdefault(stage_of(name), {"ms_per_step": 0.0, "kernels": []})
        st["ms_per_step"] += per_step
        st["kernels"].append(name)
    abytes = stage_algorithmic_bytes(P, cfg["c"], n_params)
    if "mix" in stages:
        stages["mix"]["note"] = "the fused threshold+mix kernel also writes the mask (4 of the mask stage's 8 B/px)"
    if "cm" not in stages and "lovasz" in stages and cfg["mode"] == "binary":
        # binary mode: the matrix comes out of the fused Lovasz front end (one read of scores and target)
        abytes["lovasz"] += abytes.pop("cm")
        stages["lovasz"]["note"] = "includes the confusion matrix (fused front end): 8C+8+16 B/px"
    total_ms = 0.0
    for st, rec in stages.items():
        b = abytes.get(st)
        ms = rec["ms_per_step"]
        total_ms += ms
        rec["ms_per_step"] = round(ms, 5)
        rec["alg_GB"] = None if b is None else round(b / 1e9, 5)
        rec["GBps"] = None if (b is None or ms <= 0) else round(b / 1e9 / (ms * 1e-3), 1)
        rec["frac_of_hbm_peak"] = None if rec["GBps"] is None else round(rec["GBps"] / peak, 4)
    return kernels, stages, total_ms


def load_traffic(stage_kernels):
    """DRAM bytes per launch of the dominant stage's kernels from the committed `ncu --set full` capture of
    this workload (profiles/traffic.json: kernel -> {dram_bytes_per_launch, source}); never measured live
    (no number is taken under a profiler)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tr = json.load(f)
    except (OSError, ValueError):
        return None, None
    total, src = 0, None
    for k in stage_kernels:
        rec = tr.get(k)
        if rec is None or rec.get("dram_bytes_per_launch") is None:
            continue
        total += rec["dram_bytes_per_launch"]
        src = rec.get("source", src)
    return (total or None), src


def measure_step_config(b200ssl, cfg, args, rank, world, device, headline, clocks=None):
    """One loss-path configuration on this rank's GPU: timed regions, stage table, baselines."""
    from b200ssl import _lib
    debug = os.environ.get("B200SSL_BENCH_DEBUG")
    inp = make_inputs(device, rank, cfg=cfg)
    P = cfg["n"] * cfg["h"] * cfg["w"]
    n_params = sum(p.numel() for p in inp["params"])
    peak, peak_src = measured_peak()
    # Multi-GPU: the batch shards by image, the only exchange is [confusion matrix || loss] once per
    # step.  It goes over NVLink peer memory (b200ssl_peer_*): posted by the block that finalises the loss, which
    # also completes the previous step's exchange -- no extra launch, no NCCL call on the data path.  If the box
    # cannot map peer memory every rank falls back to ONE torch.distributed all-reduce per step (collective choice).
    peer = reducer = None
    no_exchange = world > 1 and bool(os.environ.get("B200SSL_BENCH_NO_EXCHANGE"))   # diagnosis: per-GPU spread
    if world > 1 and not no_exchange:
        peer = b200ssl.utils.make_peer_all_reduce(cfg["c"] * cfg["c"], 1, device)
        if peer is None:
            reducer = b200ssl.utils.StepReducer(cfg["c"], 1, device, backend="dist")
    use_graph = bool(os.environ.get("B200SSL_BENCH_GRAPH"))
    step = make_step(b200ssl, cfg, peer, static_outputs=True, graph=use_graph, ring=cfg["ring"])
    step.bind_parameters(inp["params"], inp["ema_params"])     # like constructing an optimizer over the lists
    torch.manual_seed(0)            # the reference seeds every rank with 0 (distributed_trainer.py:17)

    def one_step():
        out = step(inp["image_a"], inp["image_b"], inp["teacher_a"], inp["teacher_b"], inp["scores"],
                   inp["target"], inp["params"], inp["ema_params"])
        if reducer is not None:
            reducer.all_reduce(out["cm"], [out["loss"]])
        return out

    end_region = peer.result if peer is not None else None
    # Warm-up: at least args.warmup (>= 3) steps; the headline additionally runs until ~1 s of GPU load has
    # passed so that clocks and power state have settled.  Every step of a multi-rank run contains an exchange,
    # so all ranks run the SAME number of steps: the decision is taken in chunks and agreed on by all ranks.
    n_w = 0
    t_w = time.perf_counter()
    chunk = max(args.warmup, 3)
    while True:
        for _ in range(chunk):
            one_step()
        n_w += chunk
        done = (not headline) or (time.perf_counter() - t_w) >= 1.0
        if world > 1:
            flag = torch.tensor([1 if done else 0], device=device, dtype=torch.int32)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            done = bool(int(flag))
        if done:
            break
        chunk = 200
    if end_region is not None:
        end_region()
    sync_all(device, world)

    # ---- timed regions: K steps each, device-resident inputs ----
    steps = args.steps if headline else max(3, min(args.steps, 20))
    n_regions = max(3, math.ceil(MIN_TIMED_STEPS / steps)) if headline else 3
    l0 = _lib.launch_count()
    per_rank_log = []
    region_ms = timed_regions(one_step, steps, n_regions, device, world, end_region, clocks if headline else None,
                              debug=(cfg["key"] if debug else None), per_rank_log=per_rank_log)
    launches = (_lib.launch_count() - l0) // n_regions
    ms_step_all = sum(region_ms) / (n_regions * steps)            # every timed step counts (no region is dropped)
    value = world * P / (ms_step_all * 1e-3) / 1e6
    res = {
        "workload": cfg["name"], "value": round(value, 2), "unit": UNIT, "ms_per_step": round(ms_step_all, 4),
        "steps_per_region": steps, "regions": n_regions, "warmup_steps_run": n_w,
        "ms_per_step_regions": {"min": round(min(region_ms) / steps, 4),
                                "median": round(sorted(region_ms)[len(region_ms) // 2] / steps, 4),
                                "max": round(max(region_ms) / steps, 4)},
        "pixels_per_gpu_step": P, "classes": cfg["c"], "ema_params": n_params, "ema_tensors": len(inp["params"]),
        "lovasz": cfg["lovasz"], "gpu_launches_per_region": int(launches),
    }

    if cfg["mode"] == "softmax":
        # exact zero-delta tail pruning (csrc/lovasz.cu): share of the C*P (class, pixel) keys that had to be sorted
        torch.cuda.synchronize(device)
        ns = step._n_seg
        sorted_keys = int(step._scratch["segi"][ns:2 * ns].sum())
        res["lovasz_keys_sorted_fraction"] = round(sorted_keys / float(ns * P), 4)

    # ---- multi-GPU: the exchange must have produced exactly what a library all-reduce produces ----
    if no_exchange:
        res["collective"] = "NONE (B200SSL_BENCH_NO_EXCHANGE diagnosis run: independent replicas)"
        res["region_ms_per_rank"] = per_rank_log[:6]
    elif world > 1:
        out = one_step()
        if peer is not None:
            cm_sum, loss_sum = peer.result()
            cm_sum, loss_sum = cm_sum.view(cfg["c"], cfg["c"]).clone(), loss_sum[:1].clone()
        else:
            cm_sum, scs = reducer.all_reduce(out["cm"], [out["loss"]])
            cm_sum, loss_sum = cm_sum.clone(), scs.reshape(-1)[:1].clone()
        want_cm = out["cm"].clone()
        want_loss = out["loss"].double().reshape(1).clone()
        dist.all_reduce(want_cm)
        dist.all_reduce(want_loss)
        ok = bool(torch.equal(cm_sum, want_cm)) and bool(torch.allclose(loss_sum, want_loss, rtol=1e-6, atol=0))
        flag = torch.tensor([1 if ok else 0], device=device, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        res["collective_check"] = "exact" if int(flag) else "MISMATCH"
        if peer is not None and not headline:
            bad = 0
            try:
                peer.status()
            except Exception as e:   # noqa: BLE001
                print(f"[rank {rank}] {e}", file=sys.stderr)
                bad = 1
            flag = torch.tensor([bad], device=device, dtype=torch.int32)
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
            if int(flag):
                res["collective_check"] = "TIMEOUT"
        res["collective"] = ("[cm || loss] over NVLink peer memory (b200ssl_peer_*): step s posted and step s-1 collected by "
                             "the Lovasz finalising block, last step flushed inside the timed region" if peer is not None
                             else "one torch.distributed (NCCL) all_reduce of [cm || loss] per step")
        res["region_ms_per_rank"] = per_rank_log[:4]
        res["ema_share_of_step_bytes"] = round(12 * n_params / step_algorithmic_bytes(P, cfg["c"], n_params), 4)

    # ---- per-kernel / per-stage table (explains the number above) ----
    prof_steps = 20 if headline else 5
    kernels, stages, serial_ms = kernel_stage_table(b200ssl, cfg, inp, device, prof_steps, P, n_params, peak)
    step_bytes = step_algorithmic_bytes(P, cfg["c"], n_params)
    top = max((s for s in stages if stages[s]["alg_GB"] is not None), key=lambda s: stages[s]["ms_per_step"])
    traffic, traffic_src = load_traffic(stages[top]["kernels"]) if headline else (None, None)
    roof = {
        "kernel": f"{top} stage ({' + '.join(stages[top]['kernels'])})", "bound": "hbm",
        "achieved": stages[top]["GBps"], "peak": peak, "unit": "GB/s", "frac": stages[top]["frac_of_hbm_peak"],
        "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
        "avg_launch_ms": stages[top]["ms_per_step"],
        "share_of_kernel_time": round(stages[top]["ms_per_step"] / serial_ms, 4),
        "algorithmic_bytes": "SURVEY 8(d): compulsory bytes only (sort traffic is not algorithmic)",
        "stages": stages, "kernels": kernels,
        "step": {"alg_GB": round(step_bytes / 1e9, 4), "GBps": round(step_bytes / 1e9 / (ms_step_all * 1e-3), 1),
                 "frac": round(step_bytes / 1e9 / (ms_step_all * 1e-3) / peak, 4),
                 "frac_of_8TBps_nominal": round(step_bytes / 1e9 / (ms_step_all * 1e-3) / 8000.0, 4),
                 "kernel_ms_per_step_serial": round(serial_ms, 4),
                 "note": "stages: each kernel alone (serial issue); the timed step overlaps the mask+mix, "
                         "Lovasz and EMA chains on internal streams"},
    }
    if "mask" in stages:
        roof["mask_stage_note"] = ("the smoothing is fp32-FMA bound (2K FMA per pixel and pass, K up to 193), not HBM "
                                   "bound: its HBM fraction is low by construction (DESIGN.md)")
    res["roofline"] = roof

    # ---- baselines on rank 0 at N=1 only ----
    if world == 1:
        res["aten_cuda_baseline"] = aten_cuda_baseline(cfg, inp, device, value)
        res["cpu_baseline"] = cpu_baseline(cfg, inp, steps=2, warmup=1, n_images=cfg["cpu_images"])
    return res, inp, step, peer, reducer


# ------------------------------------------------------------------------------------------------
def end_to_end(b200ssl, cfg, inp, step, peer, reducer, args, device, world):
    """pinned host inputs -> H2D -> step -> D2H of loss + confusion matrix, every step.  The copies run on a
    second stream into two alternating device buffer sets, so the upload of step i+1 overlaps the kernels of
    step i (the host still waits for each step's result before it issues the next step, as a training loop
    reading the loss would)."""
    names = ["image_a", "image_b", "teacher_a", "teacher_b", "scores", "target"]
    host = {k: inp[k].cpu().pin_memory() for k in names}
    dev_in = [{k: torch.empty_like(inp[k]) for k in names} for _ in range(2)]
    h2d_bytes = sum(host[k].numel() * host[k].element_size() for k in names)
    res_host = [torch.empty(1 + cfg["c"] * cfg["c"], dtype=torch.float64).pin_memory() for _ in range(2)]
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
                cm_sum, loss_sum = peer.result()         # this step's exchange completes before its result is read
                packed = torch.cat([loss_sum[:1], cm_sum.reshape(-1).to(torch.float64)])
            elif reducer is not None:
                cm, scs = reducer.all_reduce(o["cm"], [o["loss"]])
                packed = torch.cat([scs.reshape(-1)[:1], cm.reshape(-1).to(torch.float64)])
            else:
                packed = torch.cat([o["loss"].reshape(1).to(torch.float64), o["cm"].reshape(-1).to(torch.float64)])
            res_host[sl].copy_(packed, non_blocking=True)
            main_stream.synchronize()                    # the caller reads the loss every step

    e2e_run(3)
    sync_all(device, world)
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    e2e_run(args.steps)
    e3.record()
    sync_all(device, world)
    ms_e2e = torch.tensor([e2.elapsed_time(e3)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms_e2e, op=dist.ReduceOp.MAX)
    ms_e2e = float(ms_e2e)
    P = cfg["n"] * cfg["h"] * cfg["w"]
    return {"value": round(world * P * args.steps / (ms_e2e * 1e-3) / 1e6, 2), "unit": UNIT,
            "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": res_host[0].numel() * 8,
            "ms_per_step": round(ms_e2e / args.steps, 4)}


# ------------------------------------------------------------------------------------------------
# baselines: the reference's own functions (oracle/_ref, staged by oracle/stage_ref.py) or the port
def reference_step_fn():
    from oracle import stage_ref
    if stage_ref.available():
        from oracle import ref_step
        ref_step.modules()
        return ref_step.loss_path_step, "reference"
    from oracle import torch_port
    return torch_port.loss_path_step, "port"


def run_reference_step(fn, cfg, d):
    return fn(d["image_a"], d["image_b"], d["teacher_a"], d["teacher_b"], d["scores"], d["target"], d["params"],
              d["ema_params"], mode=cfg["mode"], mask_proportion_range=cfg["p_range"], sigma_range=cfg["sigma_range"],
              alpha=cfg["alpha"], classes="present", per_image=False, ignore=255, num_classes=cfg["c"])


def aten_cuda_baseline(cfg, inp, device, our_value):
    """The reference's own PyTorch path on the SAME B200 (stock ATen / cub / cuDNN kernels, full batch): the
    real bar.  Unmodified reference functions when oracle/_ref is staged."""
    fn, kind = reference_step_fn()
    d = dict(inp)
    d["ema_params"] = [e.clone() for e in inp["ema_params"]]
    P = cfg["n"] * cfg["h"] * cfg["w"]
    try:
        torch.manual_seed(0)
        run_reference_step(fn, cfg, d)
        torch.cuda.synchronize(device)
        reps = 3
        t0 = time.perf_counter()
        for _ in range(reps):
            run_reference_step(fn, cfg, d)
        torch.cuda.synchronize(device)
        dt = (time.perf_counter() - t0) / reps
    except Exception as e:   # noqa: BLE001  (an out-of-memory in the stock path must not cost the whole line)
        return {"unavailable": f"{type(e).__name__}: {str(e)[:160]}"}
    finally:
        del d
        torch.cuda.empty_cache()
    val = P / dt / 1e6
    return {"value": round(val, 2), "unit": UNIT, "ms_per_step": round(dt * 1e3, 3), "kind": kind,
            "how": f"{'oracle/_ref (unmodified reference functions)' if kind == 'reference' else 'oracle/torch_port.py'} on "
                   f"cuda tensors, full batch, {reps} timed steps after 1 warm-up, wall clock around synchronize",
            "speedup_of_this_repo": round(our_value / val, 1)}


def cpu_inputs(inp, n_images):
    keys = ["image_a", "image_b", "teacher_a", "teacher_b", "scores", "target"]
    d = {k: inp[k][:n_images].cpu() for k in keys}
    d["params"] = [p.cpu() for p in inp["params"]]
    d["ema_params"] = [p.cpu().clone() for p in inp["ema_params"]]
    return d


def cpu_baseline(cfg, inp, steps, warmup, n_images):
    """The reference's CPU implementation on this host's cores: the unmodified reference functions
    (kind "reference", oracle/_ref) or the restated port.  For the large configurations a sub-batch is timed
    and scaled: every stage but the EMA is per image (mask, mix) or, in batch mode, a sort whose cost grows
    slightly faster than linearly -- so the scaled figure favours the CPU; the EMA is timed on its own and
    counted once."""
    fn, kind = reference_step_fn()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_img = min(n_images, cfg["n"])
    d = cpu_inputs(inp, n_img)
    torch.manual_seed(0)
    for _ in range(warmup):
        run_reference_step(fn, cfg, d)
    t0 = time.perf_counter()
    for _ in range(steps):
        run_reference_step(fn, cfg, d)
    dt = (time.perf_counter() - t0) / steps
    factor = cfg["n"] / n_img
    dt_full = dt
    note = ""
    if factor > 1:
        t1 = time.perf_counter()
        for _ in range(2):
            for e, p in zip(d["ema_params"], d["params"]):
                e.mul_(cfg["alpha"]).add_(p, alpha=1 - cfg["alpha"])
        t_ema = (time.perf_counter() - t1) / 2
        dt_full = (dt - t_ema) * factor + t_ema
        note = f"; scaled to the full batch: (t - t_ema) x {factor:g} + t_ema, t_ema = {t_ema * 1e3:.1f} ms"
    pix_full = cfg["n"] * cfg["h"] * cfg["w"]
    return {"value": round(pix_full / dt_full / 1e6, 3), "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{n_img} of {cfg['n']} images x {cfg['h']}x{cfg['w']} per step, full EMA parameter set, "
                      f"{steps} timed steps after {warmup} warm-up, {torch.get_num_threads()} threads{note}",
            "scale_factor": factor, "ms_per_step_sample": round(dt * 1e3, 2), "ms_per_step": round(dt_full * 1e3, 2)}


# ------------------------------------------------------------------------------------------------
def miou_sweep(b200ssl, args, rank, world, device):
    """BASELINE.json configs[4] / SURVEY 8(d) c5: 19-class confusion matrix over 10 000 synthetic 1024x2048 masks,
    streamed in chunks of 64 masks, split over the ranks (strong scaling); the int64 matrices are summed
    with the peer exchange.  int64 labels (torch.argmax's dtype) and uint8 labels."""
    M = MIOU
    C, h, w, chunk = M["c"], M["h"], M["w"], M["chunk"]
    peak, peak_src = measured_peak()
    per_rank = M["masks"] // world + (1 if rank < M["masks"] % world else 0)
    gen = torch.Generator(device=device).manual_seed(4321 + rank)
    pool_l, pool_p = [], []
    for _ in range(M["pool"]):                      # 3 distinct chunks (3.2 GB as int64 pairs) >> L2; cycled
        lab = torch.cat([coherent_labels(16, C, h, w, device, gen) for _ in range(chunk // 16)])
        other = torch.cat([coherent_labels(16, C, h, w, device, gen) for _ in range(chunk // 16)])
        wrong = torch.rand(chunk, h // 16, w // 16, device=device, generator=gen) < 0.2
        wrong = wrong.repeat_interleave(16, 1).repeat_interleave(16, 2)
        prd = torch.where(wrong, other, lab)
        lab[torch.rand(chunk, h // 16, w // 16, device=device, generator=gen).repeat_interleave(16, 1)
            .repeat_interleave(16, 2) < 0.02] = 255      # void regions
        pool_l.append(lab.contiguous())
        pool_p.append(prd.contiguous())
        del other, wrong
    peer = reducer = None
    if world > 1:
        peer = b200ssl.utils.make_peer_all_reduce(C * C, 0, device)
    out = {"workload": M["name"], "masks_total": M["masks"], "masks_per_rank": per_rank, "scaling": "strong",
           "chunk_masks": chunk, "distinct_chunks_cycled": M["pool"]}
    px_mask = h * w
    for tag, cast, bpp in (("int64", None, 16), ("uint8", torch.uint8, 2)):
        ls = pool_l if cast is None else [t.to(cast) for t in pool_l]
        ps = pool_p if cast is None else [t.to(cast) for t in pool_p]
        cm = torch.zeros((C, C), dtype=torch.int64, device=device)
        result = {}

        def sweep():
            cm.zero_()
            done, i = 0, 0
            while done < per_rank:
                m = min(chunk, per_rank - done)
                b200ssl.metrics.confusion_matrix(ls[i % len(ls)][:m], ps[i % len(ps)][:m], C, ignore_index=255, out=cm)
                done += m
                i += 1
            if peer is not None:
                result["cm_sum"] = peer.all_reduce(cm.reshape(-1), [], lazy=False)[0]
            else:
                result["cm_sum"] = cm.reshape(-1)

        sweep()                                      # warm-up sweep (also validates below)
        sync_all(device, world)
        ms = timed_regions(sweep, 1, 3, device, world)
        total_ms = sorted(ms)[1]
        # exactness: the reduced matrix equals the library all-reduce of the per-rank matrices, and its total
        # equals the number of non-void pixels
        cm_sum = result["cm_sum"].view(C, C).clone()
        want = cm.clone()
        valid = 0
        done, i = 0, 0
        while done < per_rank:
            m = min(chunk, per_rank - done)
            valid += int((ls[i % len(ls)][:m] != 255).sum())
            done += m
            i += 1
        vt = torch.tensor([valid], device=device, dtype=torch.int64)
        if world > 1:
            dist.all_reduce(want)
            dist.all_reduce(vt)
        ok = bool(torch.equal(cm_sum, want)) and int(cm_sum.sum()) == int(vt)
        flag = torch.tensor([1 if ok else 0], device=device, dtype=torch.int32)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        pixels = M["masks"] * px_mask
        gbps_gpu = (per_rank * px_mask * bpp) / 1e9 / (total_ms * 1e-3)
        rec = {"ms_total": round(total_ms, 3), "ms_runs": [round(x, 3) for x in ms],
               "value": round(pixels / (total_ms * 1e-3) / 1e6, 1), "unit": UNIT,
               "roofline": {"bound": "hbm", "alg_bytes_per_pixel": bpp, "achieved": round(gbps_gpu, 1), "peak": peak,
                            "unit": "GB/s", "frac": round(gbps_gpu / peak, 4), "peak_source": peak_src,
                            "kernel": "confusion_u8_kernel" if tag == "uint8" else "confusion_kernel"},
               "check": "exact" if int(flag) else "MISMATCH",
               "miou": round(float(b200ssl.metrics.miou_from_cm(cm_sum)), 6)}
        if world == 1:
            # stock ATen on the same GPU: the restated oracle bincount(l*C+p) per chunk
            def aten():
                acc = torch.zeros(C * C, dtype=torch.int64, device=device)
                for i in range(4):
                    l, p = pool_l[i % len(pool_l)].reshape(-1), pool_p[i % len(pool_p)].reshape(-1)
                    keep = l != 255
                    acc += torch.bincount(l[keep] * C + p[keep], minlength=C * C)
                return acc
            aten()
            torch.cuda.synchronize(device)
            t0 = time.perf_counter()
            aten()
            torch.cuda.synchronize(device)
            dt = time.perf_counter() - t0
            rec["aten_cuda_baseline"] = {"value": round(4 * chunk * px_mask / dt / 1e6, 1), "unit": UNIT,
                                         "how": "torch.bincount(l*C+p) over 4 chunks of 64 int64 masks on the same GPU"}
        out[tag] = rec
        del ls, ps
    if world == 1:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        l, p = pool_l[0][:8].cpu().reshape(-1), pool_p[0][:8].cpu().reshape(-1)
        t0 = time.perf_counter()
        keep = l != 255
        torch.bincount(l[keep] * C + p[keep], minlength=C * C)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": round(8 * px_mask / dt / 1e6, 1), "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": "8 masks of 1024x2048 (int64), torch.bincount(l*C+p) -- the reference has no "
                                         "confusion-matrix function (SURVEY a15)"}
    if peer is not None:
        peer.close()
    return out


# ------------------------------------------------------------------------------------------------
def workload_config(world, n_params=None, n_tensors=None):
    """The `config` object: identical for the b200 arm and the reference arm (same workload)."""
    W = WORKLOAD
    shapes = load_param_shapes(W["params"])
    return {"workload": W["name"], "pixels_per_gpu_step": W["n"] * W["h"] * W["w"], "classes": W["c"],
            "ema_params": sum(math.prod(s) for s in shapes), "ema_tensors": len(shapes), "lovasz": W["lovasz"],
            "sigma_range": list(W["sigma_range"]), "mask_proportion_range": list(W["p_range"])}


def run_b200(args, rank, world, local_rank):
    import b200ssl
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    which = [c for c in args.configs.split(",") if c]
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    head, inp, step, peer, reducer = measure_step_config(b200ssl, WORKLOAD, args, rank, world, device, True, clocks)
    clock_info = clocks.stop() if rank == 0 else None
    e2e = end_to_end(b200ssl, WORKLOAD, inp, step, peer, reducer, args, device, world)
    bad = 0
    if peer is not None:
        # a wait that timed out anywhere invalidates the run on every rank (checked collectively)
        try:
            peer.status()
        except Exception as e:   # noqa: BLE001
            print(f"[rank {rank}] {e}", file=sys.stderr)
            bad = 1
        flag = torch.tensor([bad], device=device, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        peer.close()
        if int(flag):
            raise RuntimeError("the NVLink peer exchange timed out on at least one rank; results are invalid")
    del inp, step
    torch.cuda.empty_cache()
    extra = {}
    for c in which:
        if c == "1":
            continue
        if c in CONFIGS:
            r, i2, s2, p2, _ = measure_step_config(b200ssl, CONFIGS[c], args, rank, world, device, False)
            if p2 is not None:
                p2.close()
            del i2, s2
            torch.cuda.empty_cache()
            extra[CONFIGS[c]["key"]] = r
        elif c == "4":
            extra[MIOU["key"]] = miou_sweep(b200ssl, args, rank, world, device)
            torch.cuda.empty_cache()
    if rank != 0:
        return None
    cfg = workload_config(world)
    cfg.update({"parallelism": f"dp{world}",
                "l2": "no flush: the 234 MB of step inputs (+268 MB outputs/workspace) exceed the 126 MB L2",
                "issue_mode": "LossPathStep(static_outputs=True): preallocated output ring, no allocator call per step"})
    line = {
        "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "timing": {"regions": head["regions"], "steps_per_region": head["steps_per_region"],
                   "timed_steps_total": head["regions"] * head["steps_per_region"],
                   "warmup_steps_run": head["warmup_steps_run"], "ms_per_step_regions": head["ms_per_step_regions"],
                   "value_is": "all timed steps of all regions / their summed device time (no region dropped)"},
        "clocks": clock_info, "e2e": e2e, "gpu_launches": head["gpu_launches_per_region"],
        "roofline": head["roofline"],
    }
    for k in ("collective", "collective_check", "region_ms_per_rank", "ema_share_of_step_bytes", "aten_cuda_baseline",
              "cpu_baseline"):
        if k in head:
            line[k] = head[k]
    if extra:
        line["configs"] = extra
    return line


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on this host's cores -- the
    UNMODIFIED reference functions staged in oracle/_ref (kind "reference"), else the restated port -- on
    the same workload as the b200 arm (all 16 images per step), rank 0 only."""
    if rank != 0:
        return None
    W = WORKLOAD
    fn, kind = reference_step_fn()
    inp = make_inputs(torch.device("cpu"), 0)
    torch.set_num_threads(os.cpu_count() or 1)
    # bounded: shrink the per-step image count only if (K+W) full steps would not end within ~4 minutes
    probe = cpu_inputs(inp, 1)
    run_reference_step(fn, W, probe)
    t0 = time.perf_counter()
    run_reference_step(fn, W, probe)
    t_img = time.perf_counter() - t0
    total_steps = max(args.steps, 1) + max(args.warmup, 1)
    n_img = int(max(1, min(W["n"], 240.0 / (total_steps * t_img))))
    base = cpu_baseline(W, inp, steps=max(args.steps, 1), warmup=max(args.warmup, 1), n_images=n_img)
    cfg = workload_config(world)
    return {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": base["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg, "sample_images_per_step": n_img,
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
    ap.add_argument("--configs", default=os.environ.get("B200SSL_BENCH_CONFIGS", "1,2,3,4"),
                    help="comma list of BASELINE.json configs to measure besides the headline configs[1]")
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
