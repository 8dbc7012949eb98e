#!/usr/bin/env python
"""GPU parity report: for every shader variant and two cameras, max / p99.9 relative error of the CUDA path against the
fp32 oracle, next to the fp32 oracle's own distance from its fp64 twin, and whether the oracle equals the reference's own
shader sources compiled as C++ (oracle/_ref) bit for bit. Run on the B200 box:
    python profiles/parity_report.py > gpurun_out/parity_report.txt
`gate` = max over all values of |err| / (1e-4*|want| + 2e-6): the fraction of the test tolerance actually used (< 1 passes).
`p99.9 rel` uses max(|want|, 1e-3) as denominator (colours live in [0,1])."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from godot_atmosphere_shader_b200 import abi, context, scenes  # noqa: E402
from oracle import pyoracle as O  # noqa: E402
from oracle import pyref as R  # noqa: E402

VARIANTS = [("no_clouds N=8", 0, 8, 0, 0), ("scatter N=32", 0, 32, 0, 0), ("scatter N=64", 0, 64, 0, 0),
            ("clouds 8+32 cheap", 0, 8, 32, 1), ("clouds_high 8+64 cheap", 0, 8, 64, 1), ("clouds_high_rm 8+64x6", 0, 8, 64, 2),
            ("cfg4 8+128x6", 0, 8, 128, 2), ("v1 N=16", 1, 16, 0, 0), ("v1_clouds 16+32", 1, 16, 32, 1)]


def stats(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    err = np.abs(a - b)
    rel = err / np.maximum(np.abs(b), 1e-3)
    return (err / (1e-4 * np.abs(b) + 2e-6)).max(), np.quantile(rel, 0.999), err.max()


def main():
    w, h = 480, 270
    shape, cube, bn = scenes.shape_texture(64, 1), scenes.coverage_cubemap(256, 1), scenes.blue_noise_tile()
    ctx = context.AtmosphereContext(0)
    ctx.upload_shape3d(shape); ctx.upload_coverage_cube(cube); ctx.upload_blue_noise(bn)
    print(f"# {torch.cuda.get_device_name(0)}; frame {w}x{h}; textures: shape 64^3, cube 6x256^2; tolerance gate: 1e-4*|want| + 2e-6")
    print(f"{'variant':26s} {'cam':3s} {'hit%':>5s} | {'CUDA vs oracle32: gate':>23s} {'p99.9 rel':>9s} {'max_abs':>9s} | {'oracle32 vs oracle64: gate':>27s} {'p99.9 rel':>9s} | discard   | oracle32 == compiled reference")
    for name, model, ns, nc, lm in VARIANTS:
        for cam_name in ("A", "B"):
            p = scenes.demo_params()
            if model == abi.SCATTER_V1:
                p.density = 0.02
            cam = scenes.camera_a(w, h) if cam_name == "A" else scenes.camera_b(w, h, p)
            depth = scenes.synth_depth(cam, p, w, h)
            ctx.set_params(p); ctx.set_variant(ns, nc, lm, model)
            rgba = torch.empty((h, w, 4), dtype=torch.float32, device="cuda")
            disc = torch.empty((h, w), dtype=torch.uint8, device="cuda")
            ctx.render_frame(cam, torch.from_numpy(depth).cuda(), w, h, rgba, disc)
            torch.cuda.synchronize()
            tex = O.Textures(lut=O.bake_lut(p), shape=shape, cube_faces=cube, blue_noise=bn)
            var = O.variant(ns, nc, lm, model)
            ref, rdisc = O.render_frame(p, var, cam, tex, depth, w, h, threads=0)
            r64, _ = O.render_frame(p, var, cam, tex, depth, w, h, dtype=np.float64, threads=0)
            pin = "n/a"
            if R.available():
                cref, cdisc = R.render_frame(p, var, cam, tex, depth, w, h, threads=0)
                pin = "bit-exact" if (np.array_equal(cdisc, rdisc) and np.array_equal(cref.view(np.uint32), ref.view(np.uint32))) else "MISMATCH"
            g = stats(rgba.cpu().numpy(), ref)
            o = stats(ref, r64)
            print(f"{name:26s} {cam_name:3s} {100 * (rdisc == 0).mean():5.1f} | {g[0]:23.3f} {g[1]:9.2e} {g[2]:9.2e} | {o[0]:27.2f} {o[1]:9.2e} | "
                  f"{'bit-exact' if np.array_equal(disc.cpu().numpy(), rdisc) else 'MISMATCH'} | {pin}")
    ctx.set_params(scenes.demo_params())
    lut_ok = np.array_equal(ctx.download_lut(), O.bake_lut(scenes.demo_params()))
    print("LUT bake bit-exact (CUDA vs oracle):", lut_ok)
    if R.available():
        print("LUT bake bit-exact (oracle vs compiled optical_depth.gdshader):",
              np.array_equal(O.bake_lut(scenes.demo_params()).view(np.uint32), R.bake_lut(scenes.demo_params()).view(np.uint32)))


if __name__ == "__main__":
    main()
