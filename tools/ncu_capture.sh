#!/bin/bash
# ncu captures (one GPU, under gpurun).  usage: tools/ncu_capture.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
NCU="ncu --clock-control none --import-source on"
# (1) launch list of the bench command itself: per-launch device times (cold-cache, serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1
# (2) full capture of every conv launch of one style_transfer step (batch 32, the bench workload)
$NCU --set full -k regex:'conv_umma_kernel|conv_smerge_kernel|conv_first_umma|conv_last_umma' -s 53 -c 19 -f -o gpurun_out/prof_conv_$TAG \
    python tools/layer_report.py --iters 1 --batch 32 > gpurun_out/ncu_conv_$TAG.log 2>&1
# (3) HBM-bound operators at [32,512,64,64] (reference-layout ops) and the arena AdaIN kernels
$NCU --set full -k regex:'plane_bulk_kernel|stats_regs_kernel|adain_regs_kernel|merge_planes' -s 8 -c 4 -f -o gpurun_out/prof_ops_$TAG \
    python tools/op_bench.py --iters 1 > gpurun_out/ncu_ops_$TAG.log 2>&1
$NCU --set full -k regex:'nhwc_stats_partial|adain_nhwc' -s 9 -c 3 -f -o gpurun_out/prof_nhwc_$TAG \
    python tools/layer_report.py --iters 1 --batch 32 > gpurun_out/ncu_nhwc_$TAG.log 2>&1
ls -la gpurun_out/*.ncu-rep
