#!/bin/bash
# focused ncu capture of selected kernels of one step (batch 8).  usage: tools/ncu_capture2.sh <tag>
TAG=${1:-x}
mkdir -p gpurun_out
NCU="ncu --clock-control none --import-source on --set full"
$NCU -k regex:'conv_umma_kernel|conv_first_umma' -s 54 -c 18 -o gpurun_out/prof_conv_$TAG \
    python tools/layer_report.py --iters 1 --batch 8 > gpurun_out/ncu_conv_$TAG.log 2>&1
ls -la gpurun_out/*.ncu-rep
