#!/usr/bin/env python
"""Does the END of a cloud kernel cost a fixed time (few long blocks running alone)? CPU experiment: the lane-occupancy model
(warp_model.cpp) prices every warp of the 3840x2160 cfg4 frame (8x4 tiles), a block costs what its slowest warp costs, and a list
scheduler plays the GPU: 148 SMs x 9 resident blocks, blocks dispatched in blockIdx order to the first free slot. Compared
dispatch orders of the block ROWS: top-down (shipped), centre-out, heaviest-row-first (needs costs of the previous frame).
usage: python profiles/microbench/tail_model.py [gpus]"""
import ctypes as C
import heapq
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from godot_atmosphere_shader_b200 import scenes  # noqa: E402
from oracle import pyoracle as O  # noqa: E402
from tests import helpers as Hh  # noqa: E402

here = os.path.dirname(os.path.abspath(__file__))
so = os.path.join(here, "libwarp_model.so")
subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-Wno-unknown-pragmas", "-x", "c++", "-shared", "-o", so,
                       os.path.join(here, "warp_model.cpp")])
L = C.CDLL(so)
w, h, steps = 3840, 2160, 128
cache = os.path.join(here, "tail_model_costs.npy")
if os.path.exists(cache):
    cost = np.load(cache)
else:
    p = scenes.demo_params()
    cam = scenes.camera_a(w, h)
    depth = scenes.synth_depth(cam, p, w, h)
    shape, cube, bn = scenes.shape_texture(64, seed=1), scenes.coverage_cubemap(256, seed=1), scenes.blue_noise_tile()
    lut = O.bake_lut(p)
    otex = O.Textures(lut=lut, shape=shape, cube_faces=cube, blue_noise=bn)
    od, dj, fr = O.make_rays(p, cam, otex, depth, w, h)
    hs = Hh.HostsimScene(lut, shape, cube, bn)
    idx = np.arange(w * h).reshape(h, w)
    warps = idx.reshape(h // 4, 4, w // 8, 8).transpose(0, 2, 1, 3).reshape(-1, 32)      # warp (ty, tx) = 8x4 pixel tile
    sel = warps.reshape(-1)
    s_od, s_dj = np.ascontiguousarray(od[sel]), np.ascontiguousarray(dj[sel])
    costs = (C.c_double * 8)(24, 46, 52, 22, 24, 8, 14, 40)
    out = (C.c_double * 16)()
    per_warp = np.zeros(len(warps), np.float64)
    L.warp_model_run(C.byref(p), C.byref(fr), hs.lut_pad.ctypes.data_as(C.c_void_p), hs.cube_pad.ctypes.data_as(C.c_void_p), C.c_int(hs.cube_res),
                     hs.shape_pad.ctypes.data_as(C.c_void_p), C.c_int(hs.nx), C.c_int(hs.ny), C.c_int(hs.nz), C.c_int(steps),
                     s_od.ctypes.data_as(C.c_void_p), s_dj.ctypes.data_as(C.c_void_p), C.c_size_t(len(warps)), costs, out,
                     per_warp.ctypes.data_as(C.c_void_p))
    cost = per_warp.reshape(h // 4, w // 8) + 400.0          # + prologue / scatter / epilogue of every warp
    np.save(cache, cost)
# block (by, bx) = 16x8 pixels = warps (2by..2by+1, 2bx..2bx+1); lifetime ~ its slowest warp
blk = np.maximum(np.maximum(cost[0::2, 0::2], cost[0::2, 1::2]), np.maximum(cost[1::2, 0::2], cost[1::2, 1::2]))
rows, cols = blk.shape
print(f"{rows} block rows x {cols} blocks; block cost (slowest warp, model instr): mean {blk.mean():.0f}, p99 {np.quantile(blk, 0.99):.0f}, max {blk.max():.0f}")


def makespan(order_rows, slots):
    """list scheduling: blocks in dispatch order go to the slot that frees first; time unit = model instructions of a warp"""
    free = [0.0] * slots
    heapq.heapify(free)
    end = 0.0
    for r in order_rows:
        for c in blk[r]:
            t = heapq.heappop(free)
            t += c
            end = max(end, t)
            heapq.heappush(free, t)
    return end


def centre_out(n):
    c = n // 2
    out = [c]
    for d in range(1, n):
        if c + d < n:
            out.append(c + d)
        if c - d >= 0:
            out.append(c - d)
    return out


for gpus in ([1, 2, 4, 8] if len(sys.argv) < 2 else [int(sys.argv[1])]):
    mine = list(range(0, rows, gpus))            # rank 0 of the interleaved shard (8-row tiles g, g+G, ..)
    slots = 148 * 9
    ideal = blk[mine].sum() / slots
    row_cost = blk[mine].sum(axis=1)
    orders = {"top-down (shipped)": mine, "centre-out": [mine[i] for i in centre_out(len(mine))],
              "heaviest row first": [mine[i] for i in np.argsort(-row_cost)]}
    print(f"N={gpus}: ideal (total / slots) {ideal:.0f}; " + "; ".join(f"{k}: {makespan(v, slots):.0f} (+{100 * (makespan(v, slots) / ideal - 1):.1f} %)" for k, v in orders.items()))

# ---- finer-grained alternatives -------------------------------------------------------------------------------------------
def makespan_blocks(costs_in_order, slots):
    free = [0.0] * slots
    heapq.heapify(free)
    end = 0.0
    for c in costs_in_order:
        t = heapq.heappop(free) + c
        end = max(end, t)
        heapq.heappush(free, t)
    return end


print("\nfiner-grained alternatives (N = GPUs, interleaved 8-row tiles, rank 0):")
rng = np.random.default_rng(0)
for gpus in (1, 8):
    mine = list(range(0, rows, gpus))
    b = blk[mine]                                   # [rows_mine, cols]
    ideal = b.sum() / (148 * 9)
    flat = b.reshape(-1)
    res = {"top-down": makespan_blocks(flat, 148 * 9),
           "blocks shuffled": makespan_blocks(rng.permutation(flat), 148 * 9),
           "blocks heaviest first (LPT, needs per-block costs)": makespan_blocks(np.sort(flat)[::-1], 148 * 9)}
    # 64-thread blocks (2 warps: 16x4 pixels) and 32-thread blocks (1 warp), same register budget => 18 / 36 resident per SM
    wrows = np.concatenate([[2 * r, 2 * r + 1] for r in mine])
    cw = cost[wrows]                                # warp costs of my rows [2*rows_mine, 2*cols]
    b64 = np.maximum(cw[:, 0::2], cw[:, 1::2]).reshape(-1)
    res["64-thread blocks, top-down"] = makespan_blocks(b64, 148 * 18)
    res["32-thread blocks, top-down"] = makespan_blocks(cw.reshape(-1), 148 * 36)
    ideal_w = cw.sum() / (148 * 36)
    print(f"N={gpus}: ideal {ideal:.0f} (warp-granular ideal {ideal_w:.0f}); " + "; ".join(f"{k}: {v:.0f} (+{100 * (v / ideal - 1):.1f} %)" for k, v in res.items()))
