#!/bin/bash
# full ncu capture of kernels matching a regex inside one style_transfer step (batch 8 by default).
# usage: tools/ncu_full.sh <regex> <tag> [skip] [count] [batch]
RE=$1; TAG=$2; SKIP=${3:-0}; CNT=${4:-2}; B=${5:-8}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"$RE" -s $SKIP -c $CNT -o gpurun_out/prof_$TAG -f \
    python tools/layer_report.py --iters 1 --batch $B > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out/prof_$TAG.ncu-rep
