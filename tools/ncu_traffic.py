#!/usr/bin/env python
"""DRAM traffic per launch of the conv kernels from an `ncu --set full` report -> JSON for bench.py.
usage: tools/ncu_traffic.py gpurun_out/prof_conv_TAG.ncu-rep BATCH > profiles/r01_ncu_conv_traffic.json"""
import csv
import io
import json
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
ik = hdr.index("Kernel Name")
ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
launches = []
for r in data:
    if "conv_first" in r[ik] or "adain_fold" in r[ik]:
        continue  # the roofline object covers the 3x3 tcgen05 convs after conv1_1
    b = float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]
    launches.append({"kernel": r[ik].split("(")[0][-60:], "dram_bytes": b})
print(json.dumps({"batch": int(sys.argv[2]), "launches": len(launches),
                  "dram_bytes_per_launch_avg": sum(l["dram_bytes"] for l in launches) / max(len(launches), 1),
                  "source": f"ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, {sys.argv[1].split('/')[-1]}",
                  "per_launch": launches}, indent=1))
