#!/bin/bash
# stage ablation of conv_smerge_kernel (conv1_2, dec7) in a CCST_DEV build: bit 1 epilogue does nothing, 2 no MMAs,
# 8 epilogue = TMEM loads only (no math / staging / store), 16 no TMA store
for a in 0 1 2 8 10 16; do
  echo -n "ABLATE=$a: "
  CCST_ABLATE=$a timeout 300 python tools/layer_report.py 2>&1 | grep -E "^(conv1_2|dec7)" | awk '{printf "%s %s  ", $1, $3} END {print ""}'
done
