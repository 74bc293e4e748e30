#!/usr/bin/env python
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ccst_b200
from ccst_b200 import synth
dev = torch.device("cuda:0")
vgg, dec = synth.make_models(0)
eng = ccst_b200.engine_for(vgg, dec, dev)
x = synth.images(4, 512, 512, 21).to(dev)
reps = int(os.environ.get('REPS', '60'))
feats = [eng.encode(x, "fp16").clone() for _ in range(reps)]
ref = feats[-1]
bad = [i for i, f in enumerate(feats) if not torch.equal(f, ref)]
info = ""
if bad:
    d = (feats[bad[0]] - ref).abs()
    nz = d.nonzero()
    info = f"first bad {bad[0]}: {nz.shape[0]} elems differ, max {d.max().item():.4g}, n {sorted(set(nz[:,0].tolist()))} c range {nz[:,1].min().item()}-{nz[:,1].max().item()} y {nz[:,2].min().item()}-{nz[:,2].max().item()} x {nz[:,3].min().item()}-{nz[:,3].max().item()}"
print(os.environ.get("TAG", ""), "encode runs differing from the last:", bad, info, flush=True)
feat = ref
decs = [eng.decode(feat, "fp16").clone() for _ in range(reps)]
bad = [i for i, f in enumerate(decs) if not torch.equal(f, decs[-1])]
print(os.environ.get("TAG", ""), "decode runs differing from the last:", bad, flush=True)
