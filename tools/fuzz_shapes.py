#!/usr/bin/env python
"""Random-shape sweep of style_transfer / style_transfer_u8 / encoder statistics: every tensor-core engine against the
fp32 CUDA-core engine of the same library (itself pinned to the oracle by the tests), on ragged sizes that exercise
partial tiles, odd tile counts of the CTA-pair kernels, 1-pixel-wide remainders, the TMA / cp.async forms of conv1_1
and the fused uint8 loader.  usage: python tools/fuzz_shapes.py [cases] [seed] [big]"""
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ccst_b200
from ccst_b200 import drivers, synth  # noqa: F401

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
big = len(sys.argv) > 3 and sys.argv[3] == "big"  # sizes up to 640 (several tile rows / columns of every kernel)
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
dev = torch.device("cuda:0")
vgg, dec = synth.make_models(0)
eng = ccst_b200.engine_for(vgg, dec, dev)
g = torch.Generator().manual_seed(7)
# statistics of a real style image, so that the stylised images stay in the [0, 1] range the image bars refer to
stat = [t.to(dev) for t in ccst_b200.drivers.single_style_stat(eng, synth.images(1, 96, 128, 9).to(dev), "fp32")]
TOL = {"fp16": 1e-2, "fp16x3": 1e-4, "bf16x3": 2e-4}
worst = {k: 0.0 for k in TOL}
worst_u8 = {k: 0 for k in TOL}
bad = 0
for i in range(cases):
    n = rng.choice([1, 1, 2, 3, 5])
    h = rng.choice([rng.randint(16, 40), rng.randint(16, 200), 16 * rng.randint(1, 12)])
    if big:
        h = rng.choice([rng.randint(200, 640), 8 * rng.randint(30, 80)])
    w = rng.choice([rng.randint(16, 40), rng.randint(16, 300), 16 * rng.randint(1, 18), 128, 144, 256])
    if big:
        w = rng.choice([rng.randint(200, 640), 16 * rng.randint(13, 40), 512])
        n = rng.choice([1, 2])
    alpha = rng.choice([1.0, 0.5])
    x = synth.images(n, h, w, 100 + i).to(dev)
    ref = ccst_b200.style_transfer(vgg, dec, x, stat, alpha, precision="fp32")
    x_u8 = (x.permute(0, 2, 3, 1) * 255).round().to(torch.uint8).contiguous()
    ref_u8 = ccst_b200.style_transfer_u8(vgg, dec, x_u8, stat, alpha, precision="fp32").int()
    line = f"case {i:3d}: n={n} {h}x{w} a={alpha} ref [{ref.min().item():.2f},{ref.max().item():.2f}]"
    for prec, tol in TOL.items():
        out = ccst_b200.style_transfer(vgg, dec, x, stat, alpha, precision=prec)
        err = (out - ref).abs().max().item()
        o8 = ccst_b200.style_transfer_u8(vgg, dec, x_u8, stat, alpha, precision=prec).int()
        d8 = (o8 - ref_u8).abs().max().item()
        worst[prec] = max(worst[prec], err)
        worst_u8[prec] = max(worst_u8[prec], d8)
        ok = err < tol and torch.isfinite(out).all() and d8 <= (3 if prec == "fp16" else 1)
        line += f" | {prec} {err:.2e} u8 {d8}{'' if ok else '  <-- FAIL'}"
        bad += 0 if ok else 1
    print(line, flush=True)
print("worst float", worst, "worst u8 levels", worst_u8, "failures", bad)
sys.exit(1 if bad else 0)
