#!/usr/bin/env python
"""Race hunt at the bench size: the same batch-32 @512^2 step many times, fp32-tensor and uint8 paths,
both operand types; every output must be bit-identical to the first.   python tools/stress_determinism.py [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ccst_b200
from ccst_b200 import synth

dev = torch.device("cuda:0")
vgg, dec = synth.make_models(0)
eng = ccst_b200.engine_for(vgg, dec, dev)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
x = synth.images(32, 512, 512, 21).to(dev)
x8 = (x.permute(0, 2, 3, 1) * 255).round().to(torch.uint8).contiguous()
g = torch.Generator().manual_seed(7)
stat = [torch.randn((1, 512, 1, 1), generator=g).abs().to(dev), (torch.rand((1, 512, 1, 1), generator=g) + 0.1).to(dev)]
bad = 0
for prec in ("fp16", "bf16"):
    ref = eng.transfer(x, stat, 1.0, prec).clone()
    ref8 = eng.transfer_u8(x8, stat, 1.0, prec).clone()
    assert torch.isfinite(ref).all()
    d = d8 = 0
    for i in range(reps):
        # interleave other work so that buffers / L2 state differ between repetitions
        if i % 3 == 0:
            eng.encode(x[: 1 + i % 5], prec)
        d += int(not torch.equal(eng.transfer(x, stat, 1.0, prec), ref))
        d8 += int(not torch.equal(eng.transfer_u8(x8, stat, 1.0, prec), ref8))
    print(f"{prec}: transfer distinct {d}/{reps}, transfer_u8 distinct {d8}/{reps}", flush=True)
    bad += d + d8
print("OK" if bad == 0 else "MISMATCH")
sys.exit(1 if bad else 0)
