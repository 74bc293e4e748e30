#!/usr/bin/env python
"""One line per kernel launch from an .ncu-rep: duration, tensor pipe %, L2 %, DRAM bytes, smem %.
usage: tools/ncu_summary.py file.ncu-rep"""
import csv
import io
import re
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, data = rows[0], rows[2:]
want = [("dur_us", "gpu__time_duration.sum"),
        ("tensor%", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        ("sm%", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("l2%", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("l1/smem%", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("dramR_MB", "dram__bytes_read.sum"), ("dramW_MB", "dram__bytes_write.sum"),
        ("regs", "launch__registers_per_thread"), ("grid", "launch__grid_size"),
        ("smclk", "sm__cycles_active.avg")]
cols = [(n, hdr.index(m)) for n, m in want if m in hdr]
units = {n: rows[1][i] for n, i in cols}
print("units:", units)
ik = hdr.index("Kernel Name")
for r in data:
    name = r[ik]
    m = re.search(r"(\w+_kernel)<(.*?)>\(", name)
    short = (m.group(1) + "<" + m.group(2) + ">") if m else name[:50]
    short = short.replace("(int)", "").replace("(bool)", "").replace("__half", "h").replace("__nv_bfloat16", "b")
    print(f"{short:42s}", "  ".join(f"{n}={r[i]}" for n, i in cols))
