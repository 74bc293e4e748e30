#!/bin/bash
# correctness + per-layer time of the linear-slab tile geometries (CCST_GEO64 / CCST_GEO128 = 0,1,2)
mkdir -p gpurun_out
for cfg in ${TESTS:-"0 0" "1 1" "2 2"}; do
  set -- $cfg
  echo "=== tests GEO64=$1 GEO128=$2"
  CCST_GEO64=$1 CCST_GEO128=$2 timeout 300 python -m pytest tests/test_gpu_net.py -q -m gpu --timeout 120 -x \
      -k "single_conv and (bf16 or fp16)" 2>&1 | tail -3
done
for cfg in ${LAYERS:-"0 0" "1 0" "2 0" "0 1" "0 2"}; do
  set -- $cfg
  echo "=== layers GEO64=$1 GEO128=$2"
  CCST_GEO64=$1 CCST_GEO128=$2 timeout 300 python tools/layer_report.py 2>&1 | tail -20 | awk '{printf "%s %s | ", $1, $3} END {print ""}'
done
