#!/bin/bash
# One gpurun call worth of ncu evidence.  usage: tools/ncu_capture.sh <tag>
# gpurun copies back at most 64 MiB, so the reports are summarised ON THE BOX (tools/ncu_summary.py,
# tools/ncu_traffic.py) and the big per-step conv report is dropped afterwards unless KEEP_REP=1.
TAG=${1:-r01}
mkdir -p gpurun_out
NCU="ncu --clock-control none"
# (1) launch list of the bench command itself: per-launch device times (cold-cache, serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1
# (2) full capture of every conv launch of one style_transfer step (batch 32, the bench workload)
$NCU --set full -k regex:'conv_umma_kernel|conv_smerge_kernel|conv_first_umma|conv_last|conv_ups4' -s 53 -c 19 -f -o gpurun_out/prof_conv_$TAG \
    python tools/layer_report.py --iters 1 --batch 32 > gpurun_out/ncu_conv_$TAG.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_conv_$TAG.ncu-rep > gpurun_out/ncu_conv_summary_$TAG.txt 2>&1
python tools/ncu_traffic.py gpurun_out/prof_conv_$TAG.ncu-rep 32 > gpurun_out/ncu_conv_traffic_$TAG.json 2>gpurun_out/ncu_traffic_$TAG.err
[ "$KEEP_REP" = "1" ] || rm -f gpurun_out/prof_conv_$TAG.ncu-rep
# (3) HBM-bound operators at [32,512,64,64] (reference-layout ops) and the arena AdaIN kernels
$NCU --set full --import-source on -k regex:'plane_bulk_kernel|stats_regs_kernel|adain_regs_kernel|merge_planes' -s 8 -c 4 -f -o gpurun_out/prof_ops_$TAG \
    python tools/op_bench.py --iters 1 > gpurun_out/ncu_ops_$TAG.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_ops_$TAG.ncu-rep > gpurun_out/ncu_ops_summary_$TAG.txt 2>&1
$NCU --set full --import-source on -k regex:'nhwc_stats_partial|adain_nhwc|stats_nhwc|nhwc_tiles' -s 6 -c 2 -f -o gpurun_out/prof_nhwc_$TAG \
    python tools/layer_report.py --iters 1 --batch 32 > gpurun_out/ncu_nhwc_$TAG.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_nhwc_$TAG.ncu-rep > gpurun_out/ncu_nhwc_summary_$TAG.txt 2>&1
du -sh gpurun_out
