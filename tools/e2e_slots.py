#!/usr/bin/env python
"""End-to-end images/s of the overlapped batch loop (drivers.TransferPipeline) vs the number of
pipeline slots, fp32 and uint8 host batches.    python tools/e2e_slots.py [--steps 20]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ccst_b200
from ccst_b200 import drivers, synth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--batch", type=int, default=32)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    vgg, dec = synth.make_models(0)
    eng = ccst_b200.engine_for(vgg, dec, dev)
    host = [synth.images(a.batch, 512, 512, 1000 + i).pin_memory() for i in range(2)]
    host_u8 = [(h.permute(0, 2, 3, 1) * 255).round().to(torch.uint8).contiguous().pin_memory() for h in host]
    g = torch.Generator().manual_seed(7)
    stat = [torch.randn((1, 512, 1, 1), generator=g).abs().to(dev), (torch.rand((1, 512, 1, 1), generator=g) + 0.1).to(dev)]
    for u8 in (False, True):
        for slots in (2, 3, 4):
            def run(n):
                pipe = drivers.TransferPipeline(eng, "fp16", slots=slots, u8=u8)
                src = host_u8 if u8 else host
                seen = 0
                for _, o in pipe.run((src[i & 1] for i in range(n)), lambda i, x: stat, 1.0):
                    seen += o.shape[0]
                return seen
            run(4)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            n = run(a.steps)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            print(f"u8={u8} slots={slots}: {n / dt:.1f} img/s ({dt / a.steps * 1e3:.3f} ms/step)")


if __name__ == "__main__":
    main()
