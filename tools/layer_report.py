#!/usr/bin/env python
"""Per-launch device time of one style_transfer step (CUDA events recorded by the library).

    python tools/layer_report.py [--batch 32] [--size 512] [--precision fp16] [--iters 5]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ccst_b200
from ccst_b200 import synth

KIND = {0: "conv_first", 1: "conv_umma", 2: "conv_ffma", 3: "pool", 4: "adain/stats", 5: "convert", 6: "adain_fold"}
NAMES = ["conv1_1", "conv1_2", "conv2_1", "conv2_2", "conv3_1", "conv3_2", "conv3_3", "conv3_4", "conv4_1",
         "adain", "dec1", "dec2", "dec3", "dec4", "dec5", "dec6", "dec7", "dec8", "dec9"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--precision", default="fp16")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--encoder", action="store_true", help="the statistics loop's step (encode + Welford fold) instead of style_transfer")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    vgg, dec = synth.make_models(0)
    eng = ccst_b200.engine_for(vgg, dec, dev)
    x = synth.images(a.batch, a.size, a.size, 1).to(dev)
    g = torch.Generator().manual_seed(7)
    stat = [torch.randn((1, 512, 1, 1), generator=g).abs().to(dev), (torch.rand((1, 512, 1, 1), generator=g) + 0.1).to(dev)]
    state = ccst_b200.function.WelfordState(512, dev)
    step = (lambda: eng.accumulate(x, state, a.precision)) if a.encoder else (lambda: eng.transfer(x, stat, 1.0, a.precision))
    for _ in range(3):
        step()
    eng.profile(True)
    acc = None
    for _ in range(a.iters):
        step()
        recs = eng.profile_read()
        if acc is None:
            acc = recs
        else:
            for r, q in zip(acc, recs):
                r["ms"] += q["ms"]
    eng.profile(False)
    tot = sum(r["ms"] for r in acc) / a.iters
    names = NAMES if len(acc) == len(NAMES) else (NAMES[:9] + ["stats"] if len(acc) == 10 else [f"L{i}" for i in range(len(acc))])
    print(f"precision={a.precision} batch={a.batch} size={a.size}: {tot:.3f} ms/step, {a.batch / tot * 1e3:.1f} img/s")
    print(f"{'layer':10s} {'kind':12s} {'ms':>8s} {'share':>6s} {'TFLOP/s':>9s} {'GB/s':>8s}")
    for nm, r in zip(names, acc):
        ms = r["ms"] / a.iters
        tf = r["flops"] / ms / 1e9 if r["flops"] else 0.0
        gb = r["bytes"] / ms / 1e6
        print(f"{nm:10s} {KIND[r['kind']]:12s} {ms:8.4f} {ms / tot * 100:5.1f}% {tf:9.1f} {gb:8.1f}")


if __name__ == "__main__":
    main()
