#!/bin/bash
# A/B of two builds of the library on the SAME box, interleaved: tools/probes/ab/A.so, B.so (git-ignored)
# usage: tools/ab.sh [rounds] [layer_report args]
R=${1:-3}; shift
cp ccst_b200/libccst_b200.so /tmp/keep.so
for i in $(seq $R); do
  for v in A B; do
    cp tools/probes/ab/$v.so ccst_b200/libccst_b200.so
    python tools/layer_report.py "$@" > /tmp/ab_$v.log 2>&1
    echo "$v: $(head -1 /tmp/ab_$v.log | cut -d: -f2) | $(grep -E '^(conv1_1|conv1_2|dec7|dec8|dec9)' /tmp/ab_$v.log | awk '{printf "%s %s  ", $1, $3}')"
  done
done
cp /tmp/keep.so ccst_b200/libccst_b200.so
