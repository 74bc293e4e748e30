#!/bin/bash
# A/B of two builds (tools/probes/ab/A.so, B.so) on the same box: the bench's 20-step value and its >= 4 s sustained loop
R=${1:-2}
cp ccst_b200/libccst_b200.so /tmp/keep.so
for i in $(seq $R); do
  for v in A B; do
    cp tools/probes/ab/$v.so ccst_b200/libccst_b200.so
    python bench.py --steps 20 --warmup 5 --no-configs --no-eager --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        print('$v', 'value', d['value'], 'ms', d['ms_per_step'], 'sustained', d['sustained']['value'], d['sustained']['sm_mhz'], 'MHz', d['sustained']['power_w_max'], 'W')
"
  done
done
cp /tmp/keep.so ccst_b200/libccst_b200.so
