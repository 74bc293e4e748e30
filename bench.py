#!/usr/bin/env python
"""Benchmark of the CCST AdaIN style-transfer hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA path)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path

Metric (BASELINE.json): stylised images/s @512^2.  One *step* = one batch of the CCST Overall
transfer (config 3): `style_transfer(vgg, decoder, content[32,3,512,512], overall style stats,
alpha=1)` -- VGG-19 encoder to relu4_1, AdaIN, decoder.  With N GPUs every rank owns its own
batches (the path shards by image, no data-path collective) -> weak scaling.

Printed JSON (one line, rank 0):
  value       images/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e         same metric through the public batch loop `ccst_b200.drivers.overall_transfer` with
              pinned HOST buffers (uint8 HWC images in, uint8 HWC images out; the fp32-tensor form
              beside it): H2D of the batch and D2H of the stylised images inside the timed region
  sustained   the same step in a >= 3 s loop (the headline region is a short burst; long runs are
              power-capped), with the clocks sampled during it
  roofline    tcgen05 convolution kernels (dominant: ~99 % of the FLOPs) vs the measured bf16 BURST
              peak (each launch is event-timed on its own); extra `roofline_stats` / `roofline_adain`
              objects for the HBM-bound operators
  configs     every other configuration of BASELINE.json in the same run: config 1 (batch 6), config 2
              (overall statistics of a 2048-image client, sharded over the ranks, ONE NCCL all-reduce of
              the Welford moments, its cost reported), config 4 (single-style transfer, per batch and per
              image), config 5 (96^2 patches, batch 1024 per GPU)
  gpu_eager_baseline  the reference's own nn.Sequential forward (stock PyTorch eager / cuDNN) on the same
              B200 with the same inputs -- informational, the comparison SURVEY section 2 names
  cpu_baseline the oracle port of the reference path (PyTorch CPU, all host cores) on a bounded
              sample of the same workload
Weights are random-init (seeded) VGG-19/decoder, images synthetic: no dataset/checkpoint offline.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "stylised_images_per_sec_512"
UNIT = "images/s"
BATCH = 32
SIZE = 512
FLOP_PER_IMG = 253.072e9  # SURVEY.md §8d (enc 126.538 + dec 126.534 GFLOP @512^2)
FLOP_PER_IMG_96 = 8.897e9
ENC_FLOP_PER_IMG = 126.538e9
FIRST_FLOP_PER_IMG = 2 * 27 * 64 * 512 * 512  # conv1_1 (+ folded 1x1): CUDA cores in the f16x3 engine
CLIENT_IMAGES = 2048      # config 2: PACS art_painting-sized client


def workload_config(batch, world):
    """`config` of the JSON line -- identical for both arms (the reference arm runs a bounded sample of it)."""
    return {"workload": "CCST Overall K=3 transfer step (config 3): style_transfer batch 32 @512x512, "
                        "random-init VGG-19 relu4_1 encoder + decoder, overall style stats, alpha=1",
            "batch_per_gpu": batch, "image": [3, SIZE, SIZE], "parallelism": f"image-sharded x{world}",
            "l2": "two input batches alternate; per-step activations (>1 GB) exceed the 126 MB L2"}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d["bf16_tflops_sustained"],
                    src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


def load_profile_json(name):
    p = os.path.join(ROOT, "profiles", name)
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


class ClockSampler:
    """SM clock / throttle reasons sampled through NVML every 10 ms during the timed region."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, index: int):
        self.index = index
        self.sm, self.power, self.mask = [], [], 0
        self.max_sm = None
        self._stop = threading.Event()
        self.t = None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES when it is a plain list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.index
            if vis and all(v.strip().isdigit() for v in vis.split(",")):
                idx = int(vis.split(",")[self.index])
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)

            def loop():
                while not self._stop.is_set():
                    try:
                        self.sm.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                        self.power.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0)
                        self.mask |= pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                    except Exception:
                        pass
                    time.sleep(0.01)

            self.t = threading.Thread(target=loop, daemon=True)
            self.t.start()
        except Exception:
            self.t = None
        return self

    def stop(self):
        if self.t is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        self._stop.set()
        self.t.join(timeout=2)
        sm = sorted(self.sm)
        reasons = sorted(name for bit, name in self.REASONS.items() if self.mask & bit)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_sm, "reasons": reasons,
                "samples": len(sm), "power_w_max": max(self.power) if self.power else None}


def cpu_reference_throughput(steps, warmup, sample_imgs, seed=0):
    """The reference path on the host: oracle port (PyTorch CPU conv/ATen, all cores)."""
    import torch

    from ccst_b200 import synth
    from oracle import ccst_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    vgg, dec = synth.make_models(seed)
    g = torch.Generator().manual_seed(7)
    stat = [torch.randn((1, 512, 1, 1), generator=g).abs(), torch.rand((1, 512, 1, 1), generator=g) + 0.1]
    x = synth.images(sample_imgs, SIZE, SIZE, 123)
    with torch.no_grad():
        for _ in range(warmup):
            O.style_transfer(vgg, dec, x, stat, 1.0)
        t0 = time.perf_counter()
        for _ in range(steps):
            O.style_transfer(vgg, dec, x, stat, 1.0)
        dt = time.perf_counter() - t0
    return sample_imgs * steps / dt, dt / steps, cores


def run_reference(args, rank, world):
    if rank != 0:
        return
    sample = 2
    ips, sec_per_step, cores = cpu_reference_throughput(args.steps, args.warmup, sample)
    # one full batch-32 step through the same code (outside the K timed steps): the per-image rate of
    # the CPU does not depend on the batch -- one image already saturates the cores
    b32, _, _ = cpu_reference_throughput(1, 0, BATCH)
    same_batch = {"batch": BATCH, "value": round(b32, 4), "unit": UNIT}
    line = {
        "impl": "reference", "metric": METRIC, "value": round(ips, 4), "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(sec_per_step * 1e3, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(BATCH, world),
        "cpu_baseline": {"value": round(ips, 4), "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} images of the batch per step (bounded sample of the batch-{BATCH} workload; "
                                   "per-image CPU throughput is batch-independent, see same_batch_as_gpu_arm), "
                                   "oracle port of the reference path on PyTorch CPU, all host cores",
                         "same_batch_as_gpu_arm": same_batch},
        "e2e": {"value": round(ips, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def time_op(fn, iters, torch):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters  # ms


def run_ours(args, rank, world, local_rank):
    import torch

    import ccst_b200
    from ccst_b200 import _lib, drivers, overall, synth

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a B200 GPU; ccst_b200 has no CPU fallback "
                           "(use --impl reference for the CPU arm)")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist_

        dist = dist_
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    precision = args.precision
    vgg, dec = synth.make_models(0)
    eng = ccst_b200.engine_for(vgg, dec, dev)

    # every rank owns different images (seed offset by rank); two batches alternate so that a step
    # never finds its input in L2 (each batch's activations are > 1 GB anyway, L2 is 126 MB)
    host = [synth.images(args.batch, SIZE, SIZE, 1000 + 10 * rank + i).pin_memory() for i in range(2)]
    dev_in = [h.to(dev) for h in host]
    g = torch.Generator().manual_seed(7)
    stat = [torch.randn((1, 512, 1, 1), generator=g).abs().to(dev),
            (torch.rand((1, 512, 1, 1), generator=g) + 0.1).to(dev)]
    out = torch.empty((args.batch, 3, SIZE, SIZE), dtype=torch.float32, device=dev)
    host_out = torch.empty((args.batch, 3, SIZE, SIZE), dtype=torch.float32).pin_memory()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def step(i, prec=None):
        eng.transfer(dev_in[i & 1], stat, 1.0, prec or precision, out=out)

    def timed_steps(prec, n):
        for i in range(3):
            step(i, prec)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(n):
            step(i, prec)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    # ---------------- device-resident throughput ----------------
    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = _lib.lib().ccst_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(args.steps):
        step(i)
    ev1.record()
    barrier()
    launches = _lib.lib().ccst_launch_count() - l0
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    clocks = sampler.stop() if rank == 0 else None
    value = world * args.batch * args.steps / (ms_total / 1e3)
    eng.check_saturation()  # f16 range guard: raises if any store clamped

    # ---------------- end to end through the public API (host buffers) ----------------
    # The user-level loop of CCST_OverallStyleTransfer.py:149-167 (`data.to(device)` ->
    # `style_transfer` -> `output.cpu()`), via ccst_b200.drivers.overall_transfer: every step uploads
    # its batch from pinned host memory and downloads the stylised batch to pinned host memory;
    # uploads/downloads of neighbouring steps overlap the compute on separate streams.
    def e2e_loop(batches_of, nsteps, **kw):
        def run(n):
            seen = 0
            for _, out_host in drivers.overall_transfer(eng, (batches_of[i & 1] for i in range(n)), stat, 1.0,
                                                        precision, **kw):
                seen += out_host.shape[0]  # the result is in host memory here (what save_image would read)
            return seen

        run(3)
        barrier()
        t0 = time.perf_counter()
        seen = run(nsteps)
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        assert seen == args.batch * nsteps
        return world * args.batch * nsteps / dt

    e2e_value = e2e_loop(host, args.steps)
    # the same loop with the image I/O fused around the path: uint8 HWC batches in (the
    # loader's images before ToTensor), uint8 HWC batches out (what save_image encodes)
    host_u8 = [(h.permute(0, 2, 3, 1) * 255).round().to(torch.uint8).contiguous().pin_memory() for h in host]
    e2e_u8_value = e2e_loop(host_u8, args.steps, u8=True)

    # same loop without overlap: the reference's own structure, one blocking call per step
    def e2e_step_serial(i):
        x = host[i & 1].to(dev, non_blocking=True)
        y = ccst_b200.style_transfer(vgg, dec, x, stat, 1.0, precision=precision)
        host_out.copy_(y, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for i in range(2):
        e2e_step_serial(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        e2e_step_serial(i)
    barrier()
    e2e_serial_value = args.batch * args.steps / (time.perf_counter() - t0)
    img_bytes = args.batch * 3 * SIZE * SIZE * 4

    # The box's pinned-copy ceiling: every rank moves one fp32 batch up and one down per iteration, concurrently
    # on two streams, nothing else running -- what the host side (PCIe switches, memory controllers) allows the
    # fp32 host-tensor form of the loop above at this rank count, whatever the GPUs do in between.
    def host_copy_ceiling(iters=12):
        s_up, s_dn = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

        def go(n):
            for i in range(n):
                with torch.cuda.stream(s_up):
                    dev_in[i & 1].copy_(host[i & 1], non_blocking=True)
                with torch.cuda.stream(s_dn):
                    host_out.copy_(out, non_blocking=True)
            s_up.synchronize()
            s_dn.synchronize()

        go(2)
        barrier()
        t0 = time.perf_counter()
        go(iters)
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        gbs = world * iters * 2 * img_bytes / dt / 1e9
        return {"aggregate_GBps_h2d_plus_d2h": round(gbs, 1), "per_gpu_GBps_each_way": round(gbs / world / 2, 1),
                "images_per_s_ceiling": round(world * iters * args.batch / dt, 1),
                "how": "all ranks copy one pinned fp32 batch up and one down per iteration (2 x %.0f MB), two streams, "
                       "no compute; max over ranks" % (img_bytes / 1e6)}

    copy_ceiling = host_copy_ceiling()

    # (rank 0, right after the timed regions: the GPU is in the same thermal / power state as for `value`)
    line = None
    if rank == 0:
        # ---------------- per-kernel roofline (separate profiled pass, events per launch) ----------
        eng.profile(True)
        agg = {}
        psteps = min(args.steps, 5)
        for i in range(psteps):
            step(i)
            for rec in eng.profile_read():
                a = agg.setdefault(rec["kind"], dict(ms=0.0, flops=0.0, bytes=0.0, n=0))
                a["ms"] += rec["ms"]
                a["flops"] += rec["flops"]
                a["bytes"] += rec["bytes"]
                a["n"] += 1
        eng.profile(False)
        conv_kind = 2 if precision == "fp32" else 1
        conv = agg.get(conv_kind, dict(ms=1e-9, flops=0, n=1))
        step_ms_prof = sum(a["ms"] for a in agg.values()) / psteps
        conv_tflops = conv["flops"] / (conv["ms"] * 1e-3) / 1e12
        # algorithmic FLOPs of the same launches (253.072 GFLOP/image minus conv1_1's 0.906)
        alg_flops_step = (FLOP_PER_IMG - 0.906e9) * args.batch
        conv_tflops_alg = alg_flops_step / (conv["ms"] / psteps * 1e-3) / 1e12
        traffic, traffic_src = None, None
        td = load_profile_json("r02_ncu_conv_traffic.json") or load_profile_json("r01_ncu_conv_traffic.json")
        if precision != "fp32" and td and td.get("batch") == args.batch:
            traffic = td["dram_bytes_per_launch_avg"]
            traffic_src = td["source"]
        roofline = {
            "kernel": "tcgen05.mma + TMA implicit-GEMM 3x3 convs (conv_umma_kernel / conv_smerge_kernel / "
                      "conv_ups4_kernel / conv_last_rows_kernel, %d launches/step)" % round(conv["n"] / psteps)
            if precision != "fp32" else "conv_ffma_kernel (fp32 validation mode)",
            "bound": "tensor", "achieved": round(conv_tflops, 2), "peak": peaks["tf_burst"], "unit": "TFLOP/s",
            "frac": round(conv_tflops / peaks["tf_burst"], 4),
            "peak_source": f"{peaks['src']} bf16 BURST peak (every launch is timed on its own with a CUDA-event pair)",
            "achieved_algorithmic": round(conv_tflops_alg, 2),
            "frac_algorithmic": round(conv_tflops_alg / peaks["tf_burst"], 4),
            "frac_of_sustained_peak": round(conv_tflops / peaks["tf_sust"], 4),
            "flops_per_launch_avg": conv["flops"] / max(conv["n"], 1),
            "ms_per_launch_avg": round(conv["ms"] / max(conv["n"], 1), 4),
            "share_of_step": round(conv["ms"] / psteps / step_ms_prof, 4), "traffic": traffic,
            "traffic_source": traffic_src,
            "flops_basis": "`achieved` counts EXECUTED flops: the three convs that follow a nearest-x2 upsample run as "
                           "four 2x2 phase convolutions (16 instead of 36 tap-GEMMs per source pixel), so a step "
                           "executes 220.9 of the 253.07 algorithmic GFLOP/image; `achieved_algorithmic` divides the "
                           "algorithmic FLOPs of the same launches by the same time",
            "timing": f"cudaEvent pair around every launch on the launch stream, {psteps}-step pass after the timed region",
            "per_kind_ms_per_step": {str(k): round(a["ms"] / psteps, 4) for k, a in sorted(agg.items())},
            "kinds": "0 conv1_1, 1 tcgen05 convs, 2 ffma convs, 3 pool, 4 adain/stats pass, 5 layout convert, "
                     "6 AdaIN folded into dec1 (coefficients + per-image weights)",
        }

        # ---------------- HBM-bound operators on [32,512,64,64] fp32 ----------------
        feat = torch.randn((BATCH, 512, 64, 64), device=dev).relu_()
        feat2 = torch.randn((BATCH, 512, 64, 64), device=dev).relu_()  # alternate: 268 MB each > L2
        mu, sg = stat
        cnt = [0]

        def f_stats():
            cnt[0] += 1
            ccst_b200.calc_mean_std(feat if cnt[0] & 1 else feat2)

        def f_adain():
            cnt[0] += 1
            ccst_b200.adaIN_StyleStat_ContentFeat(feat if cnt[0] & 1 else feat2, [mu, sg])

        ms_stats = time_op(f_stats, 20, torch)
        ms_adain = time_op(f_adain, 20, torch)
        nbytes = feat.numel() * 4
        b_stats = nbytes + 8 * BATCH * 512
        b_adain = 2 * nbytes + 8 * 512
        ops_traffic = load_profile_json("r02_ncu_ops_traffic.json") or {}

        def op_traffic(key):
            e = ops_traffic.get(key)
            return (e["dram_bytes"], e["source"]) if e else (None, None)

        t_s, t_s_src = op_traffic("calc_mean_std")
        t_a, t_a_src = op_traffic("adain_stat")
        roofline_stats = {"kernel": "plane_bulk_kernel<0> (TMA bulk-staged single-pass Welford; calc_mean_std [32,512,64,64] fp32)",
                          "bound": "hbm", "achieved": round(b_stats / ms_stats / 1e6, 1), "peak": peaks["hbm"], "unit": "GB/s",
                          "frac": round(b_stats / ms_stats / 1e6 / peaks["hbm"], 4),
                          "traffic": t_s, "traffic_source": t_s_src, "algorithmic_bytes": b_stats,
                          "ms": round(ms_stats, 5), "peak_source": peaks["src"] + " copy bandwidth",
                          "note": "includes torch.empty of the outputs and the ctypes call per launch"}
        roofline_adain = {"kernel": "plane_bulk_kernel<2> (statistics + re-normalisation in one HBM pass; adaIN_StyleStat_ContentFeat [32,512,64,64] fp32)",
                          "bound": "hbm", "achieved": round(b_adain / ms_adain / 1e6, 1), "peak": peaks["hbm"],
                          "unit": "GB/s", "frac": round(b_adain / ms_adain / 1e6 / peaks["hbm"], 4),
                          "traffic": t_a, "traffic_source": t_a_src, "algorithmic_bytes": b_adain,
                          "ms": round(ms_adain, 5), "peak_source": peaks["src"] + " copy bandwidth"}
        del feat, feat2

        # ---------------- same step with the other operand type (single GPU, informative) ----------
        other = {}
        ACCURACY = {"bf16": "1.3e-2 .. 1.6e-2 max-abs (misses the 1e-2 bar; strict xfail in the tests)",
                    "fp16": "2e-3 max-abs (bar 1e-2)",
                    "fp16x3": "2.5e-6 max-abs vs the fp32 reference (fp32 mode's bar 1e-4): split f16 operands, 4x the MMAs",
                    "bf16x3": "3e-5 max-abs (fp32 mode's bar 1e-4): split bf16 operands, fp32 exponent range, 4x the MMAs"}
        for alt in ("bf16", "fp16", "fp16x3", "bf16x3"):
            if alt != precision and precision != "fp32":
                ms_alt = timed_steps(alt, args.steps if len(alt) == 4 else max(3, args.steps // 4))
                other[alt] = {"ms_per_step": round(ms_alt, 4), "images_per_s_per_gpu": round(args.batch / ms_alt * 1e3, 2),
                              "image_error": ACCURACY[alt]}

    barrier()

    # ---------------- the other BASELINE.json configurations (all ranks take part) ----------------
    configs = {}
    if not args.no_configs:
        configs = run_configs(args, torch, dist, dev, rank, world, eng, vgg, dec, stat, peaks, barrier, max_over_ranks)

    # ---------------- sustained: the same step for >= 3 s (all ranks, power-capped regime) -------
    sustained = None
    if not args.no_sustained:
        # (long loops run at lower clocks than the burst above, so this lasts longer than sustain_s)
        n_sus = max(args.steps, int(1.25 * args.sustain_s * 1e3 / (ms_total / args.steps)) + 1)
        sam = ClockSampler(local_rank).start() if rank == 0 else None
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(n_sus):
            step(i)
        b.record()
        barrier()
        ms_sus = max_over_ranks(a.elapsed_time(b))
        ck = sam.stop() if sam is not None else {}
        sustained = {"steps": n_sus, "seconds": round(ms_sus / 1e3, 3),
                     "value": round(world * args.batch * n_sus / (ms_sus / 1e3), 2), "unit": UNIT,
                     "ms_per_step": round(ms_sus / n_sus, 4), "sm_mhz": ck.get("sm_mhz"),
                     "reasons": ck.get("reasons"), "power_w_max": ck.get("power_w_max")}

    if rank == 0:
        # ---------------- the reference's own modules through stock PyTorch on this GPU -----------
        eager = None
        if not args.no_eager:
            eager = gpu_eager_baseline(torch, vgg, dec, dev_in, stat, args, value / world)

        # ---------------- CPU baseline beside it (bounded sample) ----------------
        cpu = None
        if not args.no_cpu_baseline:
            ips, sec, cores = cpu_reference_throughput(steps=2, warmup=1, sample_imgs=2)
            cpu = {"value": round(ips, 4), "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "2 steps x 2 images @512x512 (batch 32 scaled down), oracle port of the "
                             "reference style_transfer on PyTorch CPU, 1 warm-up"}

        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms_total / args.steps, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp16": "f16", "bf16": "bf16", "fp32": "f32"}[precision], "data": "synthetic",
            "dtype_note": ("fp32 FFMA validation mode (CUDA cores), not the product path" if precision == "fp32" else
                           "tcgen05 kind::f16 MMA, %s operands, fp32 accumulation in TMEM (f16 and bf16 run at the same "
                           "tensor peak; f16 meets the 1e-2 image tolerance, bf16 operands do not -- see DESIGN.md "
                           "Numerics)" % ("f16" if precision == "fp16" else "bf16")),
            "other_precisions": other,
            "config": workload_config(args.batch, world),
            "tflops_per_gpu": round(value / world * FLOP_PER_IMG / 1e12, 2),
            # headline e2e: the batch loop fed with what the reference's loader actually holds (uint8 HWC
            # images, before ToTensor) and returning what save_image encodes (uint8 HWC) -- ToTensor and
            # the quantisation run on the GPU, H2D + D2H inside the timed region.  The strict
            # drop-in form (fp32 NCHW host tensors both ways, 4x the PCIe bytes) is reported beside it.
            "e2e": {"value": round(e2e_u8_value, 2), "unit": UNIT, "h2d_bytes_per_step": img_bytes // 4,
                    "d2h_bytes_per_step": img_bytes // 4,
                    "api": "ccst_b200.drivers.overall_transfer(engine, pinned uint8 HWC host batches, style_stat, "
                           "u8=True): the batch loop of CCST_OverallStyleTransfer.py:149-167 with ToTensor "
                           "(cjm_util/data_helper.py:45) and save_image's quantisation (:167) on the GPU "
                           "(ccst_style_transfer_u8); H2D/compute/D2H double-buffered on 3 streams",
                    "fp32_host_tensors": {"value": round(e2e_value, 2), "h2d_bytes_per_step": img_bytes,
                                          "d2h_bytes_per_step": img_bytes,
                                          "api": "drivers.overall_transfer(engine, pinned fp32 NCHW host batches, "
                                                 "style_stat): data.to(device) -> style_transfer -> output.cpu() "
                                                 "as in the reference, overlapped",
                                          "host_copy_ceiling": copy_ceiling},
                    "serial_per_gpu": round(e2e_serial_value, 2),
                    "serial_api": "x.to(device); ccst_b200.style_transfer(vgg, decoder, x, style_stat, alpha); out.cpu() "
                                  "per step with fp32 tensors, no overlap (rank 0)"},
            "gpu_launches": int(launches),
            "sustained": sustained,
            "roofline": roofline, "roofline_stats": roofline_stats, "roofline_adain": roofline_adain,
            "configs": configs, "gpu_eager_baseline": eager,
            "cpu_baseline": cpu, "clocks": clocks,
        }
    barrier()
    if dist is not None:
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)


def run_configs(args, torch, dist, dev, rank, world, eng, vgg, dec, stat, peaks, barrier, max_over_ranks):
    """Configs 1, 2, 4, 5 of BASELINE.json, each a short timed region on every rank (max over ranks)."""
    import random

    import ccst_b200
    from ccst_b200 import drivers, overall, synth

    precision = args.precision
    res = {}

    def dev_timed(fn, iters, warm=2):
        for i in range(warm):
            fn(i)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(iters):
            fn(i)
        b.record()
        barrier()
        return max_over_ranks(a.elapsed_time(b)) / iters

    # ---- config 1: AdaIN single-style transfer, batch 6 @512^2 (the reference's CPU-runnable case)
    x6 = [synth.images(6, SIZE, SIZE, 2000 + 10 * rank + i).to(dev) for i in range(2)]
    o6 = torch.empty((6, 3, SIZE, SIZE), dtype=torch.float32, device=dev)
    ms = dev_timed(lambda i: eng.transfer(x6[i & 1], stat, 1.0, precision, out=o6), 20)
    res["config1_batch6_512"] = {
        "workload": "style_transfer batch 6 @512x512 (BASELINE config 1 on the GPU), device-resident",
        "value": round(world * 6 / ms * 1e3, 1), "unit": UNIT, "ms_per_step": round(ms, 4),
        "tensor_frac_algorithmic": round(6 * FLOP_PER_IMG / (ms * 1e-3) / 1e12 / peaks["tf_burst"], 4)}
    del x6, o6

    # ---- config 2: overall style statistics of one 2048-image client, sharded over the ranks ----
    # (mean_std_computation_effcientMem.py:117-137): contiguous image ranges per rank, uint8 HWC host
    # batches of 32 uploaded under the encoder, ONE all-reduce of the 1+2C fp64 moments (+ image count)
    begin, end = overall.shard_range(CLIENT_IMAGES, rank, world)
    hb = [torch.randint(0, 256, (BATCH, SIZE, SIZE, 3), dtype=torch.uint8,
                        generator=torch.Generator().manual_seed(3000 + rank + i)).pin_memory() for i in range(2)]

    def client_batches(first, last):
        i = 0
        for b0 in range(first, last, BATCH):
            n = min(BATCH, last - b0)
            yield hb[i & 1][:n]
            i += 1

    group = None  # default process group (NCCL) when world > 1
    c2 = {"workload": f"overall style statistics of a {CLIENT_IMAGES}-image client @512x512, batch 32, images "
                      f"sharded over {world} rank(s), uint8 host batches uploaded in the timed region, one "
                      "all-reduce(sum) of the 1025 fp64 Welford moments + image count"}
    # three engines: the drivers' default (f16x3: split f16 operands, promoted fp32 accumulation; meets the
    # 1e-5 statistics bar) on the whole client, the plain 16-bit tensor-core encoder (does not meet it) and the
    # fp32 CUDA-core engine on a 64-image-per-rank sample
    STAT_NOTE = {"fp16x3": "< 1e-5 relative (measured 2e-6, tests/test_gpu_net.py::test_overall_statistics_loop[fp16x3])",
                 "fp32": "< 1e-5 relative (measured 6e-7, tests/test_gpu_net.py::test_overall_statistics_loop[fp32])"}
    KEY = {"fp16x3": "f16x3_engine_default", "fp32": "fp32_cuda_core_engine"}
    for prec, n_img in ((overall.STATS_PRECISION, CLIENT_IMAGES), (precision, CLIENT_IMAGES),
                        ("fp32", min(CLIENT_IMAGES, 64 * world))):
        key = KEY.get(prec, "tensor_core_encoder_16bit")
        if key in c2:
            continue
        b_, e_ = overall.shard_range(n_img, rank, world)
        drivers.overall_statistics(eng, client_batches(b_, min(e_, b_ + BATCH)), prec, group)  # warm-up
        barrier()
        t0 = time.perf_counter()
        mean, std, seen = drivers.overall_statistics(eng, client_batches(b_, e_), prec, group)
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        assert seen == n_img, (seen, n_img)
        c2[key] = {"precision": prec, "images": n_img, "value": round(n_img / dt, 1), "unit": UNIT,
                   "seconds": round(dt, 4),
                   "tensor_frac_algorithmic": round(n_img / world * ENC_FLOP_PER_IMG / dt / 1e12 / peaks["tf_burst"], 4)
                   if prec != "fp32" else None,
                   "tensor_frac_executed": round(n_img / world * 4 * (ENC_FLOP_PER_IMG - FIRST_FLOP_PER_IMG) / dt / 1e12
                                                 / peaks["tf_burst"], 4) if prec == "fp16x3" else None,
                   "statistics_vs_reference": STAT_NOTE.get(prec, "~1e-3 relative (16-bit encoder): above the 1e-5 bar")}
    c2["value"] = c2[KEY["fp16x3"]]["value"]
    c2["unit"] = UNIT
    # the collective on its own: all-reduce of the 1026-double payload, CUDA events, 50 repetitions
    payload = torch.zeros((1026,), dtype=torch.float64, device=dev)
    if dist is not None:
        for _ in range(5):
            dist.all_reduce(payload)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(50):
            dist.all_reduce(payload)
        b.record()
        torch.cuda.synchronize()
        c2["allreduce_us"] = round(max_over_ranks(a.elapsed_time(b)) / 50 * 1e3, 2)
        c2["allreduce"] = "NCCL all_reduce(SUM) of 1026 fp64 (8.2 KB) per client, once; latency-bound"
    else:
        c2["allreduce_us"] = 0.0
        c2["allreduce"] = "single rank: no collective"
    res["config2_overall_stats_2048"] = c2
    del hb

    # ---- config 4: single-style transfer (CCST_SingleStyleTransfer.py:176-223), batch 32 @512^2 ----
    # through the public batch loops with pinned host batches: per BATCH one random style image (the
    # reference's semantics) and per IMAGE (BASELINE.json's wording); style encode + statistics included
    hc = [synth.images(BATCH, SIZE, SIZE, 4000 + 10 * rank + i).pin_memory() for i in range(2)]
    styles = [synth.images(1, SIZE, SIZE, 4100 + k) for k in range(8)]
    styles_ragged = [synth.images(1, 512, 512 + 128 * (k % 3 - 1), 4200 + k) for k in range(6)]  # 4:3 / 1:1 / 3:4 mix
    nb = max(4, min(args.steps, 10))

    def run_loop(fn, sts):
        seen = 0
        for _, o in fn(eng, (hc[i & 1] for i in range(nb)), sts, 1.0, precision, seed=1):
            seen += o.shape[0]
        return seen

    c4 = {"workload": "single-style transfer, batch 32 @512x512, pinned fp32 host batches in/out, style image "
                      "encode + statistics inside the loop"}
    for key, fn, sts in (("per_batch_style", drivers.single_transfer, styles_ragged),
                         ("per_image_style", drivers.single_transfer_per_image, styles)):
        run_loop(fn, sts)
        barrier()
        t0 = time.perf_counter()
        seen = run_loop(fn, sts)
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        c4[key] = {"value": round(world * seen / dt, 1), "unit": UNIT, "batches": nb}
    res["config4_single_style"] = c4
    del hc

    # ---- config 5: Camelyon17-like 96x96 patches, batch 1024 per GPU, 4 target styles -----------
    n5 = 1024
    x5 = [synth.images(n5, 96, 96, 5000 + 10 * rank + i).to(dev) for i in range(2)]
    o5 = torch.empty((n5, 3, 96, 96), dtype=torch.float32, device=dev)
    gs = torch.Generator().manual_seed(11)
    stats4 = [[torch.randn((1, 512, 1, 1), generator=gs).abs().to(dev), (torch.rand((1, 512, 1, 1), generator=gs) + 0.1).to(dev)]
              for _ in range(4)]
    ms = dev_timed(lambda i: eng.transfer(x5[i & 1], stats4[i & 3], 1.0, precision, out=o5), 8)
    res["config5_camelyon_96"] = {
        "workload": "Overall K=4 transfer at 96x96 patches, batch 1024 per GPU, 4 styles in turn, device-resident",
        "value": round(world * n5 / ms * 1e3, 1), "unit": UNIT, "ms_per_step": round(ms, 4),
        "tensor_frac_algorithmic": round(n5 * FLOP_PER_IMG_96 / (ms * 1e-3) / 1e12 / peaks["tf_burst"], 4),
        "note": "24x24 / 12x12 maps fill 75 % / 56 % of the 8x16 pixel tiles"}
    del x5, o5
    eng.check_saturation()
    return res


def gpu_eager_baseline(torch, vgg, dec, dev_in, stat, args, ours_ips):
    """The reference's nn.Sequential modules (net.py:6-92 as built by ccst_b200.net) run by stock PyTorch
    eager (cuDNN / ATen) on this GPU + the reference AdaIN formula (function.py:26-33), same inputs."""
    import copy

    from oracle import ccst_oracle as O

    dev = dev_in[0].device
    res = {"what": "stock PyTorch eager forward of the same nn.Sequential encoder/decoder + the reference's "
                   "AdaIN (function.py:26-33) on the same B200, batch %d @512x512, CUDA events" % args.batch}
    try:
        v, d = copy.deepcopy(vgg).to(dev), copy.deepcopy(dec).to(dev)
        modes = (("fp32_tf32_off", dict(tf32=False, amp=False, cl=False)),
                 ("fp32_tf32_on", dict(tf32=True, amp=False, cl=False)),
                 ("bf16_autocast_channels_last", dict(tf32=True, amp=True, cl=True)))
        for name, m in modes:
            torch.backends.cudnn.allow_tf32 = m["tf32"]
            torch.backends.cuda.matmul.allow_tf32 = m["tf32"]
            torch.backends.cudnn.benchmark = True
            vv, dd = (v.to(memory_format=torch.channels_last), d.to(memory_format=torch.channels_last)) if m["cl"] else (v, d)
            xs = [x.contiguous(memory_format=torch.channels_last) if m["cl"] else x for x in dev_in]

            def fwd(i):
                with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=m["amp"]):
                    return O.style_transfer(vv, dd, xs[i & 1], stat, 1.0)

            for i in range(2):
                fwd(i)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 4
            a.record()
            for i in range(n):
                fwd(i)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / n
            ips = args.batch / ms * 1e3
            res[name] = {"value": round(ips, 1), "unit": UNIT, "ms_per_step": round(ms, 3),
                         "ours_over_eager": round(ours_ips / ips, 2)}
        torch.backends.cudnn.allow_tf32 = True
    except Exception as e:  # informational: never take the bench line down
        res["error"] = repr(e)[:200]
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--precision", default="fp16", choices=["fp16", "bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true")
    ap.add_argument("--no-eager", action="store_true")
    ap.add_argument("--no-sustained", action="store_true")
    ap.add_argument("--sustain-s", type=float, default=3.0)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29400 + os.getpid() % 500), __file__] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
