#!/usr/bin/env python
"""Run encode / style_transfer several times and report whether the outputs are bit-identical
(debugging aid for races; toggle kernels with CCST_FIRST_WS / CCST_LAST_GATHER / CCST_CTA_PAIR)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ccst_b200
from ccst_b200 import synth

dev = torch.device("cuda:0")
vgg, dec = synth.make_models(0)
eng = ccst_b200.engine_for(vgg, dec, dev)
x = synth.images(4, 512, 512, 21).to(dev)
g = torch.Generator().manual_seed(7)
stat = [torch.randn((1, 512, 1, 1), generator=g).abs().to(dev), (torch.rand((1, 512, 1, 1), generator=g) + 0.1).to(dev)]
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
prec = sys.argv[2] if len(sys.argv) > 2 else "fp16"
feats = [eng.encode(x, prec).clone() for _ in range(reps)]
print("encode distinct:", sum(1 for f in feats[1:] if not torch.equal(f, feats[0])), "of", reps - 1, flush=True)
decs = [eng.decode(feats[0], prec).clone() for _ in range(reps)]
print("decode distinct:", sum(1 for f in decs[1:] if not torch.equal(f, decs[0])), "of", reps - 1, flush=True)
outs = [ccst_b200.style_transfer(vgg, dec, x, stat, 1.0, precision=prec).clone() for _ in range(reps)]
print("transfer distinct:", sum(1 for f in outs[1:] if not torch.equal(f, outs[0])), "of", reps - 1,
      "max diff", max((f - outs[0]).abs().max().item() for f in outs[1:]), flush=True)
ad = [ccst_b200.adaIN_StyleStat_ContentFeat(feats[0], stat).clone() for _ in range(reps)]
print("adain(nchw) distinct:", sum(1 for f in ad[1:] if not torch.equal(f, ad[0])), "of", reps - 1, flush=True)
