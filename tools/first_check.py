#!/usr/bin/env python
"""Tiny conv1_1 / encoder check against the oracle (debugging aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ccst_b200
from ccst_b200 import synth
from oracle import ccst_oracle as O

dev = torch.device("cuda:0")
vgg, dec = synth.make_models(0)
eng = ccst_b200.engine_for(vgg, dec, dev)
for (n, h, w) in ((1, 40, 40), (2, 64, 136), (1, 128, 256)):
    x = synth.images(n, h, w, 3)
    with torch.no_grad():
        ref = O.encode_relu4_1(vgg, x)
    out = eng.encode(x.to(dev), "fp16")
    torch.cuda.synchronize()
    print((n, h, w), "max err", (out.cpu() - ref).abs().max().item(), "ref max", ref.abs().max().item(), flush=True)
