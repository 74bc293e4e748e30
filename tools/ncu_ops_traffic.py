#!/usr/bin/env python
"""DRAM traffic per launch of the NCHW operators at [32,512,64,64] from an `ncu --set full` report of
`tools/op_bench.py --iters 1` (first 12 launches: calc_mean_std x4, AdaIN x4, Welford accumulate x4)
-> JSON for bench.py (roofline_stats.traffic / roofline_adain.traffic).
usage: tools/ncu_ops_traffic.py gpurun_out/prof_ops_TAG.ncu-rep > profiles/r02_ncu_ops_traffic.json"""
import csv
import io
import json
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
ik = hdr.index("Kernel Name")
ir, iw, idur = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tscale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}
res = {}
for r in data:
    name = r[ik]
    flat = name.replace("(int)", "").replace(" ", "")
    if "plane_bulk_kernel<0," in flat:
        key = "calc_mean_std"
    elif "plane_bulk_kernel<2," in flat:
        key = "adain_stat"
    elif "welford_bulk_kernel" in flat or "plane_bulk_kernel<1," in flat:
        key = "welford_accumulate"
    else:
        continue
    b = float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]
    res[key] = {"kernel": name.split("(")[0][-60:] + name[name.find("<"):name.find(">") + 1][:40], "dram_bytes": b,
                "dur_us_under_ncu": float(r[idur]) * tscale.get(units[idur], 1.0),
                "shape": [32, 512, 64, 64],
                "source": f"ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum, last "
                          f"launch of the kind in {sys.argv[1].split('/')[-1]}"}
print(json.dumps(res, indent=1))
