#!/bin/bash
# First-contact / regression run on a B200 box (under gpurun).  Every stage runs in its own process
# with its own timeout so a trapping kernel cannot take the later stages down with it.
# usage: tools/gpu_check.sh [stage ...]   (default: all)
mkdir -p gpurun_out
STAGES=${@:-"info ops conv32 conv16 net smoke bench"}
run() { # name timeout cmd...
  local name=$1 to=$2; shift 2
  echo "=== $name: $*"
  timeout $to "$@" > gpurun_out/$name.log 2>&1
  local rc=$?
  echo "=== $name rc=$rc"; tail -n ${TAILN:-25} gpurun_out/$name.log
  return $rc
}
for s in $STAGES; do
  case $s in
    info)   run info 60 nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total,power.limit --format=csv ;;
    ops)    run ops 600 python -m pytest tests/test_gpu_ops.py -q -m gpu --timeout 120 ;;
    conv32) run conv32 600 python -m pytest tests/test_gpu_net.py -q -m gpu --timeout 120 -k "single_conv and fp32" ;;
    conv16) run conv16 600 python -m pytest tests/test_gpu_net.py -q -m gpu --timeout 120 -k "single_conv and (bf16 or fp16)" ;;
    net)    run net 900 python -m pytest tests/test_gpu_net.py -q -m gpu --timeout 300 -k "not single_conv" ;;
    drivers) run drivers 600 python -m pytest tests/test_gpu_drivers.py -q -m gpu --timeout 300 ;;
    smoke)  run smoke 300 python __graft_entry__.py smoke ;;
    bench)  run bench 900 python bench.py --steps 5 --warmup 3 ;;
    bench32) run bench32 900 python bench.py --steps 2 --warmup 1 --precision fp32 --batch 4 --no-cpu-baseline ;;
    layers) run layers 300 python tools/layer_report.py ;;
    layersb) run layersb 300 python tools/layer_report.py --precision bf16 ;;
    all)    run all 1200 python -m pytest tests -q -m gpu --timeout 300 ;;
  esac
done
