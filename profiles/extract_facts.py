#!/usr/bin/env python
"""Reads the `ncu --set full` captures of a round (profiles/rNN/*.ncu-rep or gpurun_out/) and writes
profiles/roofline_traffic.json: per workload, DRAM bytes and warp instructions per launch of the dominant kernel.
usage: python profiles/extract_facts.py <dir with prof_<key>.ncu-rep files> [round tag]"""
import csv
import glob
import json
import os
import subprocess
import sys

src = sys.argv[1]
tag = sys.argv[2] if len(sys.argv) > 2 else ""
out = {}
for rep in sorted(glob.glob(os.path.join(src, "prof_*.ncu-rep"))):
    key = os.path.basename(rep)[len("prof_"):-len(".ncu-rep")]
    rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
    hdr = rows[0]

    def col(name):
        i = hdr.index(name)
        unit, val = rows[1][i], float(rows[2][i].replace(",", ""))
        scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0, "inst": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9}.get(unit, 1.0)
        return val * scale

    out[key] = {"dram_bytes": int(col("dram__bytes_read.sum") + col("dram__bytes_write.sum")),
                "warp_instructions": int(col("smsp__inst_executed.sum")),
                "kernel": rows[2][hdr.index("Kernel Name")], "ncu_duration_s": col("gpu__time_duration.sum"),
                "source": f"profiles/{tag}/prof_{key}.summary.txt" if tag else os.path.basename(rep)}
# printed, not written: profiles/roofline_traffic.json (read by bench.py) is keyed by workload and curated by hand from this
print(json.dumps(out, indent=1))
