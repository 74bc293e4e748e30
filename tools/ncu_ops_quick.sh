#!/bin/bash
# quick ncu sections for every stand-alone operator launch of op_bench (one iteration per op)
mkdir -p gpurun_out
ncu --clock-control none --section SpeedOfLight --section WarpStateStats --section SchedulerStats --section LaunchStats --section Occupancy --section MemoryWorkloadAnalysis \
    -k regex:'plane_bulk_kernel|stats_regs|adain_regs|merge_planes|stats_stream' -c 80 -f -o gpurun_out/prof_opsq \
    python tools/op_bench.py --iters 1 > gpurun_out/ncu_opsq.log 2>&1
ls -la gpurun_out/prof_opsq.ncu-rep
