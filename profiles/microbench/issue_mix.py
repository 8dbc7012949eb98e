#!/usr/bin/env python
"""Builds issue_mix.cu, counts the SASS instructions of each mode's main loop (per march step) and runs it on the GPU."""
import os
import re
import subprocess
import sys

here = os.path.dirname(os.path.abspath(__file__))
exe = os.path.join(here, "issue_mix")
if "--no-build" not in sys.argv:
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-o", exe, os.path.join(here, "issue_mix.cu")])
sass = subprocess.run(["cuobjdump", "-sass", exe], capture_output=True, text=True, check=True).stdout
counts = {}
mode, ins = None, []


def finish():
    if mode is None:
        return
    back = [(a, int(re.search(r"0x([0-9a-f]+)", b).group(1), 16)) for a, b in ins if re.search(r"\bBRA", b) and re.search(r"0x([0-9a-f]+)", b)
            and int(re.search(r"0x([0-9a-f]+)", b).group(1), 16) < a]
    end, start = max(back, key=lambda t: t[0] - t[1])      # the longest back edge = the main loop
    body = [b for a, b in ins if start <= a <= end]
    steps = round(sum(1 for b in body if b.startswith("FFMA")) / (22.0 if mode in (2, 3) else 21.0))   # FFMA per step by construction
    counts[mode] = (len(body), steps)


for line in sass.splitlines():
    m = re.search(r"Function : .*mixILi(\d)E", line)
    if m:
        finish()
        mode, ins = int(m.group(1)), []
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);", line)
    if m and mode is not None:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
finish()
per_step = [counts[k][0] / max(counts[k][1], 1) for k in range(5)]
print("SASS loop instructions / steps per mode:", {k: counts[k] for k in sorted(counts)})
if "--count-only" not in sys.argv:
    subprocess.check_call([exe] + [f"{v:.3f}" for v in per_step])
