#!/bin/bash
# ncu timing of the arena AdaIN kernels inside one step. usage: tools/ncu_adain.sh
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'nhwc|adain' -c 12 --csv --log-file gpurun_out/adain_launches.csv python tools/layer_report.py --iters 1 > gpurun_out/ncu_adain.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/adain_launches.csv')))
hdr=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
H=rows[hdr]
for r in rows[hdr+1:]:
    d=dict(zip(H,r))
    print(d['Kernel Name'][:40], d['Metric Name'], d['Metric Value'], d['Metric Unit'])
PY
