#!/bin/bash
# per-layer stage ablation of the whole step in a CCST_DEV build: 0 = everything, 1 = epilogues do nothing (MMA + load side
# alone), 2 = no MMAs (load + epilogue side alone).  Kernels without hooks (conv1_1, dec9) are unaffected.
for a in 0 1 2; do
  echo "=== ABLATE=$a"
  CCST_ABLATE=$a timeout 300 python tools/layer_report.py 2>&1 | tail -19 | awk '{printf "%s %s | ", $1, $3} END {print ""}'
done
