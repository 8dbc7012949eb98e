#!/usr/bin/env python
"""Driver for ncu captures: sets one workload up and launches the render kernel a few times in one mapping.
usage: python profiles/prof_one.py <cfg2|cfg3A|cfg3C|cfg4A|cfg4C> <linear|tiled|frame> [launches]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402

W = bench.Workload
WORK = {"cfg2": W(1920, 1080, 32, 0, 0, "B"), "cfg2_n8": W(1920, 1080, 8, 0, 0, "B"), "cfg3A": W(1920, 1080, 8, 64, 1, "A"),
        "cfg3C": W(1920, 1080, 8, 64, 1, "C"), "cfg4A": W(3840, 2160, 8, 128, 2, "A"), "cfg4C": W(3840, 2160, 8, 128, 2, "C")}
wl = WORK[sys.argv[1]]
mapping = sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 2
torch.cuda.set_device(0)
R = bench.Runner(torch, wl, 0)        # launches make_rays + ONE linear render_rays (skip it with ncu -s 1)
out = torch.empty_like(R.d_rgba)
for _ in range(n):
    if mapping == "frame":
        R.ctx.render_frame(R.cam, R.d_depth, wl.width, wl.height, out, None)
    else:
        R.render_rays(grid=(mapping == "tiled"))
torch.cuda.synchronize()
print("done", sys.argv[1:], flush=True)
