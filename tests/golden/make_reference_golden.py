#!/usr/bin/env python
"""Generates tests/golden/reference_golden_v1.npz — golden vectors produced by THE REFERENCE ITSELF.

Source of the numbers: oracle/_ref/libatmo_ref.so, i.e. the reference's own GDShader files under /root/reference compiled
as C++ by oracle/ref/build_ref.py (purely syntactic rewrite). That library can only be built where /root/reference exists
(this container); these fixtures carry its outputs everywhere else: each of the 7 shipped entry shaders run by name with
its own #defines on two small frames (cameras A and B), plus the baked LUT of optical_depth.gdshader (sub-sampled values
and a SHA-256 of all 65 536 floats). Inputs are regenerated from seeds by the test.
Re-run only if the reference changes:   python tests/golden/make_reference_golden.py
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from godot_atmosphere_shader_b200 import abi, scenes  # noqa: E402
from godot_atmosphere_shader_b200.planet_atmosphere import SHADER_VARIANTS  # noqa: E402
from oracle import pyoracle as O  # noqa: E402
from oracle import pyref as R  # noqa: E402

W, H = 40, 24
SHAPE_N, CUBE_RES = 16, 16


def scene(shader, cam_name):
    """Seeded inputs shared by the generator and tests/test_reference_golden.py."""
    model = SHADER_VARIANTS[shader][0]
    p = scenes.demo_params()
    p.sphere_depth_factor = 0.125
    a = 0.37
    p.cloud_coverage_rotation[:] = (np.cos(a), np.sin(a), -np.sin(a), np.cos(a))
    if model == abi.SCATTER_V1:
        p.density = 0.02
    cam = scenes.camera_a(W, H, orbit_deg=25.0) if cam_name == "A" else scenes.camera_b(W, H, p)
    depth = scenes.synth_depth(cam, p, W, H)
    shape = scenes.shape_texture(SHAPE_N, seed=7)
    cube = scenes.coverage_cubemap(CUBE_RES, seed=7)
    return p, cam, depth, shape, cube, scenes.blue_noise_tile()


def main():
    assert R.reference_present(), "needs /root/reference (the fixtures are the reference's own outputs)"
    R.build(force=True)
    out = {}
    for shader, (model, ns, nc, light) in sorted(SHADER_VARIANTS.items()):
        for cam_name in ("A", "B"):
            p, cam, depth, shape, cube, bn = scene(shader, cam_name)
            tex = O.Textures(lut=R.bake_lut(p), shape=shape, cube_faces=cube, blue_noise=bn)
            rgba, disc = R.render_frame(p, O.variant(ns, nc, light, model), cam, tex, depth, W, H, shader=shader)
            out[f"{shader}/{cam_name}/rgba"] = rgba
            out[f"{shader}/{cam_name}/discard"] = disc
    for name, p in (("demo", scenes.demo_params()), ("template", scenes.template_params())):
        lut = R.bake_lut(p)
        out[f"lut/{name}/sample"] = lut[::16, ::16].copy()
        out[f"lut/{name}/sha256"] = np.frombuffer(hashlib.sha256(lut.tobytes()).digest(), dtype=np.uint8)
    out["entry_defines"] = np.array([R.entry_shaders()[s] for s in sorted(SHADER_VARIANTS)], dtype=np.int32)
    path = os.path.join(ROOT, "tests", "golden", "reference_golden_v1.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
