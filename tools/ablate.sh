#!/bin/bash
# pipeline-stage ablation of the tcgen05 conv kernels (timing only; outputs are garbage).  The switch
# exists only in a development build of the library:   python -m ccst_b200.build --dev --force
# CCST_ABLATE bit 1 = epilogue does nothing, 2 = no MMAs, 4 = no activation (A) loads
for a in 0 1 2 4 3 5 6 7; do
  echo "=== ABLATE=$a"
  CCST_ABLATE=$a timeout 300 python tools/layer_report.py 2>&1 | tail -20 | awk '{printf "%s %s | ", $1, $3} END {print ""}'
done
