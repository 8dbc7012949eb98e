#!/usr/bin/env python
"""NVLink byte counters of the fused render + delivery kernel (for `ncu --metrics nvltx__bytes...`): ONE process, TWO GPUs.
GPU 0 renders a 1920x1080x32 frame (camera B) with b200atmo_render_frame_peers and stores the pixels straight into a buffer
that lives on GPU 1 (peer access, the same store path as the multi-process symmetric-memory case) — float4 and half4 tiles.
ncu can replay this kernel (the stores are idempotent), which it cannot do for a multi-rank job.
usage: ncu --metrics nvltx__bytes.sum,nvltx__bytes_data_user.sum,nvltx__bytes_data_protocol.sum,nvlrx__bytes.sum,gpu__time_duration.sum \\
           -k regex:render_frame_kernel --devices 0 python profiles/nvlink_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from godot_atmosphere_shader_b200 import abi, sharding  # noqa: E402

assert torch.cuda.device_count() >= 2, "needs two GPUs"
torch.cuda.set_device(0)
wl = bench.Workload(1920, 1080, 32, 0, 0, "B")
R = bench.Runner(torch, wl, 0)
w, h = wl.width, wl.height
remote32 = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda:1")
remote16 = torch.zeros((h, w, 4), dtype=torch.float16, device="cuda:1")
probe = torch.ones(4, device="cuda:0")
remote32.view(-1)[:4].copy_(probe)          # a cross-device copy makes torch enable peer access 0 <-> 1
torch.cuda.synchronize()
local = torch.empty((h, w, 4), dtype=torch.float32, device="cuda:0")
R.ctx.render_frame(R.cam, R.d_depth, w, h, local, None)
for fmt, buf in ((abi.COLOR_RGBA32F, remote32), (abi.COLOR_RGBA16F, remote16)):
    t = sharding.peer_targets([buf.data_ptr()], rgba_format=fmt)
    R.ctx.render_frame_peers(R.cam, R.d_depth, w, h, t)
    torch.cuda.synchronize()
    want = local if fmt == abi.COLOR_RGBA32F else local.to(torch.float16)
    ok = torch.equal(buf.to("cuda:0"), want)
    print(f"format {fmt}: {w * h * (16 if fmt == 0 else 8)} algorithmic bytes over NVLink, remote buffer matches: {ok}")
    assert ok
print("done")
