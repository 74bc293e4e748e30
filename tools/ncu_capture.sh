#!/bin/bash
# ncu captures (one GPU, under gpurun).  usage: tools/ncu_capture.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
NCU="ncu --clock-control none --import-source on"
# launch list with device times for one step (shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -s 63 -c 21 --csv --log-file gpurun_out/launches_$TAG.csv \
    python tools/layer_report.py --iters 1 > gpurun_out/ncu_launch_$TAG.log 2>&1
# full capture of the conv kernels of one step (17 tcgen05 convs + conv1_1), batch 8 to keep replays short
$NCU --set full -k regex:conv_umma_kernel -s 51 -c 17 -o gpurun_out/prof_conv_$TAG \
    python tools/layer_report.py --iters 1 --batch 8 > gpurun_out/ncu_conv_$TAG.log 2>&1
$NCU --set full -k regex:conv_first_umma -s 3 -c 1 -o gpurun_out/prof_first_$TAG \
    python tools/layer_report.py --iters 1 --batch 8 > gpurun_out/ncu_first_$TAG.log 2>&1
# HBM-bound operators at [32,512,64,64]
$NCU --set full -k regex:'stats_regs_kernel|adain_regs_kernel' -s 6 -c 3 -o gpurun_out/prof_ops_$TAG \
    python tools/op_bench.py --iters 1 > gpurun_out/ncu_ops_$TAG.log 2>&1
ls -la gpurun_out/*.ncu-rep
