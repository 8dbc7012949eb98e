"""Golden vectors produced by the reference's own shader sources (compiled as C++ in the build container, see
tests/golden/make_reference_golden.py): they hold everywhere, also where oracle/_ref cannot be built.
CPU: the hand-written oracle reproduces them bit for bit. GPU: the CUDA path matches them within the north_star tolerance."""
import hashlib
import os

import numpy as np
import pytest

from godot_atmosphere_shader_b200 import abi, scenes
from godot_atmosphere_shader_b200.planet_atmosphere import SHADER_VARIANTS
from oracle import pyoracle as O
from tests import helpers as Hh
from tests.golden.make_reference_golden import H, W, scene

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_golden_v1.npz"))


def test_fixture_is_complete():
    assert len([k for k in GOLD.files if k.endswith("/rgba")]) == 14
    assert GOLD["entry_defines"].shape == (7, 4)
    for row, name in zip(GOLD["entry_defines"], sorted(SHADER_VARIANTS)):
        model, ns, nc, light = SHADER_VARIANTS[name]
        assert tuple(row[:3]) == (int(model == abi.SCATTER_V1), ns, nc) and int(row[3]) == int(light == abi.LIGHT_RAYMARCHED)


@pytest.mark.parametrize("name", ["demo", "template"])
def test_oracle_lut_equals_the_reference_bake(name):
    p = scenes.demo_params() if name == "demo" else scenes.template_params()
    lut = O.bake_lut(p)
    assert np.array_equal(lut[::16, ::16].view(np.uint32), GOLD[f"lut/{name}/sample"].view(np.uint32))
    assert hashlib.sha256(lut.tobytes()).digest() == GOLD[f"lut/{name}/sha256"].tobytes()


@pytest.mark.parametrize("shader", sorted(SHADER_VARIANTS))
@pytest.mark.parametrize("cam_name", ["A", "B"])
def test_oracle_equals_the_reference_frames(shader, cam_name):
    model, ns, nc, light = SHADER_VARIANTS[shader]
    p, cam, depth, shape, cube, bn = scene(shader, cam_name)
    tex = O.Textures(lut=O.bake_lut(p), shape=shape, cube_faces=cube, blue_noise=bn)
    got, disc = O.render_frame(p, O.variant(ns, nc, light, model), cam, tex, depth, W, H)
    assert np.array_equal(disc, GOLD[f"{shader}/{cam_name}/discard"])
    assert np.array_equal(got.view(np.uint32), GOLD[f"{shader}/{cam_name}/rgba"].view(np.uint32))


@pytest.mark.gpu
@pytest.mark.parametrize("shader", sorted(SHADER_VARIANTS))
def test_cuda_path_matches_the_reference_frames(cuda_ctx_factory, shader):
    import torch
    ctx = cuda_ctx_factory()
    model, ns, nc, light = SHADER_VARIANTS[shader]
    for cam_name in ("A", "B"):
        p, cam, depth, shape, cube, bn = scene(shader, cam_name)
        ctx.set_params(p)
        ctx.set_variant(ns, nc, light, model)
        ctx.upload_blue_noise(bn)
        ctx.upload_shape3d(shape)
        ctx.upload_coverage_cube(cube)
        rgba = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
        disc = torch.empty((H, W), dtype=torch.uint8, device="cuda")
        ctx.render_frame(cam, torch.from_numpy(depth).cuda(), W, H, rgba, disc)
        torch.cuda.synchronize()
        assert np.array_equal(disc.cpu().numpy(), GOLD[f"{shader}/{cam_name}/discard"])
        Hh.assert_rgba_close(rgba.cpu().numpy(), GOLD[f"{shader}/{cam_name}/rgba"], what=f"{shader}/{cam_name} vs reference golden")
    if model == abi.SCATTER_V2:
        assert hashlib.sha256(ctx.download_lut().tobytes()).digest() == GOLD["lut/demo/sha256"].tobytes()
