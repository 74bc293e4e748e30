#!/usr/bin/env python
"""Device time of the stand-alone NCHW fp32 operators (calc_mean_std, AdaIN, Welford accumulate)
through the raw C ABI (no per-call torch allocation), CUDA events, inputs larger than L2.

    python tools/op_bench.py [--iters 20]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ccst_b200 import _lib

PEAK = 6650.0  # fallback of /opt/skills/guides/B200_PROFILING.md, used only without MEASURED_PEAKS.json
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    PEAK = json.load(open(p))["hbm_gbs"]


def timeit(fn, iters):
    for _ in range(3):
        fn(0)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--shapes", default="", help="e.g. 64x512x28x28,32x512x32x32 (default: the five profile shapes)")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    L = _lib.lib()
    _lib.check(L.ccst_set_device(0))
    st = torch.cuda.current_stream().cuda_stream
    rows = []
    shapes = [(32, 512, 64, 64), (6, 512, 64, 64), (1024, 512, 12, 12), (64, 512, 28, 28), (32, 512, 32, 32),
              (128, 512, 32, 32)]  # the last: the reference's own default (256x256 images, batch 128)
    if a.shapes:
        shapes = [tuple(int(v) for v in sh.split("x")) for sh in a.shapes.split(",")]
    for (n, c, h, w) in shapes:
        xs = [torch.randn((n, c, h, w), device=dev).relu_() for _ in range(2)]  # alternate buffers
        out = torch.empty_like(xs[0])
        mean = torch.empty((n * c,), device=dev)
        std = torch.empty_like(mean)
        mu = torch.randn((c,), device=dev)
        sg = torch.rand((c,), device=dev) + 0.1
        state = torch.zeros((2 + 2 * c,), dtype=torch.float64, device=dev)
        scratch = torch.empty((2 * n * c,), device=dev)
        nbytes = xs[0].numel() * 4
        t = timeit(lambda i: L.ccst_stats_nchw_f32(xs[i & 1].data_ptr(), n * c, h * w, 1e-5, 1, mean.data_ptr(),
                                                   std.data_ptr(), st), a.iters)
        rows.append(("calc_mean_std", (n, c, h, w), t, nbytes + 8 * n * c))
        t = timeit(lambda i: L.ccst_adain_stat_nchw_f32(xs[i & 1].data_ptr(), n, c, h * w, mu.data_ptr(), sg.data_ptr(),
                                                        0, 1.0, 1e-5, out.data_ptr(), st), a.iters)
        rows.append(("adain_stat", (n, c, h, w), t, 2 * nbytes + 8 * c))
        t = timeit(lambda i: L.ccst_welford_accumulate_nchw_f32(xs[i & 1].data_ptr(), n, c, h * w, state.data_ptr(),
                                                                scratch.data_ptr(), st), a.iters)
        rows.append(("welford_accumulate", (n, c, h, w), t, nbytes + 16 * n * c))
        del xs, out
    print(f"{'op':20s} {'shape':22s} {'MB':>7s} {'ms':>8s} {'GB/s':>8s} {'frac of ' + str(PEAK):>14s}")
    for name, shape, ms, by in rows:
        gbs = by / ms / 1e6
        print(f"{name:20s} {str(shape):22s} {by / 1e6:7.1f} {ms:8.4f} {gbs:8.1f} {gbs / PEAK:14.3f}")


if __name__ == "__main__":
    main()
