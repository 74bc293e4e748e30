#!/usr/bin/env python
"""One small style_transfer (2 x 3 x 64 x 64) through one engine, for compute-sanitizer runs.
usage: compute-sanitizer --tool memcheck python tools/san_small.py fp32|fp16|bf16|fp16x3|bf16x3
(the tensor-core engines also run the uint8 entry point: ToTensor fused into conv1_1, quantised store in the last conv)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ccst_b200
from ccst_b200 import synth

prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
dev = torch.device("cuda:0")
vgg, dec = synth.make_models(0)
x = synth.images(2, 64, 64, 5).to(dev)
g = torch.Generator().manual_seed(7)
stat = [torch.randn((1, 512, 1, 1), generator=g).abs().to(dev), (torch.rand((1, 512, 1, 1), generator=g) + 0.1).to(dev)]
out = ccst_b200.style_transfer(vgg, dec, x, stat, 1.0, precision=prec)
torch.cuda.synchronize()
print(prec, "ok", tuple(out.shape), float(out.mean()))
if prec != "fp32":
    x_u8 = (x.permute(0, 2, 3, 1) * 255).round().to(torch.uint8).contiguous()
    o8 = ccst_b200.style_transfer_u8(vgg, dec, x_u8, stat, 1.0, precision=prec)
    st = ccst_b200.function.WelfordState(512, dev)
    ccst_b200.engine_for(vgg, dec, dev).accumulate_u8(x_u8, st, prec)
    torch.cuda.synchronize()
    print(prec, "u8 ok", tuple(o8.shape), float(o8.float().mean()), float(st.buf[0]))
