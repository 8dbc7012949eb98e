#!/usr/bin/env python
"""Lane-occupancy model of the cloud march (CPU; see warp_model.cpp). Usage:
    python profiles/microbench/warp_model.py [--width 3840 --height 2160 --cloud-steps 128 --camera A --every 24]
Prints the priced strategies for warps of 32 consecutive rays of a row (ray kernel) and of 8x4 pixel tiles (frame kernel)."""
import argparse
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from godot_atmosphere_shader_b200 import scenes  # noqa: E402
from oracle import pyoracle as O  # noqa: E402
from tests import helpers as Hh  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--cloud-steps", type=int, default=128)
    ap.add_argument("--camera", default="A")
    ap.add_argument("--every", type=int, default=24, help="sample every k-th warp")
    # section costs in issued warp-instructions: base step, cube fetch, shape fetch + density, hit tail, light-step base,
    # light tail, queue push, queue pop + ordered accumulate
    ap.add_argument("--costs", default="24,46,52,22,24,8,14,40")
    a = ap.parse_args()
    here = os.path.dirname(os.path.abspath(__file__))
    so = os.path.join(here, "libwarp_model.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-Wno-unknown-pragmas", "-x", "c++", "-shared",
                           "-o", so, os.path.join(here, "warp_model.cpp")])
    L = C.CDLL(so)
    w, h = a.width, a.height
    p = scenes.demo_params()
    cam = {"A": scenes.camera_a, "B": lambda w, h: scenes.camera_b(w, h, p), "C": lambda w, h: scenes.camera_c(w, h, p)}[a.camera](w, h)
    depth = scenes.synth_depth(cam, p, w, h)
    shape, cube, bn = scenes.shape_texture(64, seed=1), scenes.coverage_cubemap(256, seed=1), scenes.blue_noise_tile()
    lut = O.bake_lut(p)
    otex = O.Textures(lut=lut, shape=shape, cube_faces=cube, blue_noise=bn)
    od, dj, fr = O.make_rays(p, cam, otex, depth, w, h)
    hs = Hh.HostsimScene(lut, shape, cube, bn)
    costs = (C.c_double * 8)(*[float(x) for x in a.costs.split(",")])
    idx = np.arange(w * h).reshape(h, w)
    layouts = {
        "row (ray kernel: 32 consecutive rays)": idx.reshape(-1, 32),
        "8x4 tile (frame kernel)": idx[: h // 4 * 4, : w // 8 * 8].reshape(h // 4, 4, w // 8, 8).transpose(0, 2, 1, 3).reshape(-1, 32),
    }
    for name, warps in layouts.items():
        sel = warps[:: a.every].reshape(-1)
        s_od, s_dj = np.ascontiguousarray(od[sel]), np.ascontiguousarray(dj[sel])
        out = (C.c_double * 16)()
        L.warp_model_run(C.byref(p), C.byref(fr), hs.lut_pad.ctypes.data_as(C.c_void_p), hs.cube_pad.ctypes.data_as(C.c_void_p),
                         C.c_int(hs.cube_res), hs.shape_pad.ctypes.data_as(C.c_void_p), C.c_int(hs.nx), C.c_int(hs.ny), C.c_int(hs.nz),
                         C.c_int(a.cloud_steps), s_od.ctypes.data_as(C.c_void_p), s_dj.ctypes.data_as(C.c_void_p),
                         C.c_size_t(len(sel) // 32), costs, out, None)
        pt, lq, ideal, ls, l1, l2, l3, ws, wh, items, batches = out[:11]
        print(f"== {a.width}x{a.height} camera {a.camera}, {a.cloud_steps} cloud steps x 6 light steps, warp = {name}; {len(sel) // 32} warps sampled")
        print(f"   marching lanes per warp-step {ls / max(ws, 1):.2f}/32; in shell {l1 / max(ls, 1):.3f}, shape fetched {l2 / max(ls, 1):.3f}, density>0 {l3 / max(ls, 1):.4f} of lane-steps")
        print(f"   warp-steps with a hit {wh / max(ws, 1):.3f}; lanes hit in those {l3 / max(wh, 1):.2f}/32; light items {items:.0f} in {batches:.0f} batches ({items / max(batches, 1):.1f}/batch)")
        print(f"   warp-steps with no lane in the shell {out[11] / max(ws, 1):.3f}")
        print(f"   issued warp-instr (model): per-thread {pt:.4g}  light-queue {lq:.4g} ({lq / pt:.3f}x)  ideal {ideal:.4g} ({ideal / pt:.3f}x)")


if __name__ == "__main__":
    main()
