#!/usr/bin/env python
"""Generates tests/golden/atmo_golden_v1.npz — golden vectors for the atmosphere hot path.

The reference has no tests / fixtures and cannot run here (GDShader needs the Godot engine + Vulkan), so these
vectors are produced by the ORACLE (oracle/atmo_oracle.hpp, fp32 instantiation, g++ -O2 -ffp-contract=off) from
seeded inputs; the fp64 instantiation of the same template is stored next to them as the rounding-error bound.
They pin (a) the oracle against compiler / platform drift and accidental edits, (b) the CUDA path against a
committed artefact. Re-run only when the oracle's DEFINITION changes:   python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from godot_atmosphere_shader_b200 import abi, scenes  # noqa: E402
from oracle import pyoracle as O  # noqa: E402
from tests import helpers as Hh  # noqa: E402

VARIANTS = {
    "no_clouds": (abi.SCATTER_V2, 8, 0, abi.LIGHT_NONE),
    "scatter32": (abi.SCATTER_V2, 32, 0, abi.LIGHT_NONE),
    "clouds": (abi.SCATTER_V2, 8, 32, abi.LIGHT_CHEAP),
    "clouds_high": (abi.SCATTER_V2, 8, 64, abi.LIGHT_CHEAP),
    "clouds_high_rm": (abi.SCATTER_V2, 8, 64, abi.LIGHT_RAYMARCHED),
    "v1_clouds": (abi.SCATTER_V1, 16, 32, abi.LIGHT_CHEAP),
}
N_RAYS = 640
SHAPE_N, CUBE_RES = 16, 16


def inputs():
    p = scenes.demo_params()
    p.sphere_depth_factor = 0.125
    a = 0.37
    p.cloud_coverage_rotation[:] = (np.cos(a), np.sin(a), -np.sin(a), np.cos(a))
    shape = scenes.shape_texture(SHAPE_N, seed=7)
    cube = scenes.coverage_cubemap(CUBE_RES, seed=7)
    od, dj, fr = Hh.random_rays(N_RAYS, p, seed=2024)
    return p, shape, cube, od, dj, fr


def params_for(p, variant):
    q = p.copy()
    if variant[0] == abi.SCATTER_V1:
        q.density = 0.02  # v1 is chaotic at the demo density (tests/test_gpu_parity.py::_params_for)
    return q


def main():
    p, shape, cube, od, dj, fr = inputs()
    out = {"origin_depth": od, "dir_jitter": dj, "frame": np.frombuffer(bytes(fr), dtype=np.float32).copy(),
           "params": np.frombuffer(bytes(p), dtype=np.float32).copy(), "shape": shape, "cube": cube}
    lut = O.bake_lut(p)
    out["lut_sha256"] = np.frombuffer(hashlib.sha256(lut.tobytes()).digest(), dtype=np.uint8).copy()
    out["lut_sub"] = lut[::8, ::8].copy()
    out["cube_padded_sha256"] = np.frombuffer(hashlib.sha256(O.cube_build_padded(cube).tobytes()).digest(), dtype=np.uint8).copy()
    for name, v in VARIANTS.items():
        q = params_for(p, v)
        tex = O.Textures(lut=O.bake_lut(q), shape=shape, cube_faces=cube)
        var = O.variant(v[1], v[2], v[3], v[0])
        rgba, disc = O.render_rays(q, var, fr, tex, od, dj)
        rgba64, _ = O.render_rays(q, var, fr, tex, od, dj, dtype=np.float64)
        out[f"rgba_{name}"] = rgba
        out[f"rgba64_{name}"] = rgba64
        out[f"discard_{name}"] = disc
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "atmo_golden_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
