#!/bin/bash
# stage ablation of conv_ups4_kernel (dec8) in a CCST_DEV build: bit 1 epilogue does nothing, 2 no MMAs,
# 8 epilogue = TMEM loads + math only (no staging, no store), 16 no TMA store (staging kept)
for a in 0 1 2 8 16 10 18; do
  echo -n "ABLATE=$a: "
  CCST_ABLATE=$a timeout 300 python tools/layer_report.py 2>&1 | grep "^dec8"
done
