"""NoiseCubemap generator (noise_cubemap.gd:101-155): oracle KATs, host build of the device code, node mirror,
and (GPU) bit-exact parity of b200atmo_generate_noise_cubemap."""
import ctypes as C

import numpy as np
import pytest

from godot_atmosphere_shader_b200 import abi, scenes
from godot_atmosphere_shader_b200.noise_cubemap import FastNoiseLiteParams, NoiseCubemap
from oracle import pyoracle as O
from tests import helpers as Hh

NOISE = abi.B200AtmoNoise(1337, 0.01, 5, 2.0, 0.5)


def test_noise_known_answers():
    # gradient noise vanishes on the integer lattice; value at the cell centre is a sum of +-0.5 gradient terms / 8
    for p in ((0.0, 0.0, 0.0), (3.0, 4.0, 5.0), (-2.0, 7.0, -1.0)):
        assert O.noise3(*p, seed=1) == 0.0
    v = O.noise3(0.5, 0.5, 0.5, seed=1)
    assert abs(v) <= 1.0 and (v * 8) == int(v * 8)
    vals = [O.noise3(x * 0.37 + 0.1, x * 0.11, -x * 0.23, seed=9) for x in range(200)]
    assert -1.0 <= min(vals) < -0.2 and 0.2 < max(vals) <= 1.0
    assert O.noise3(0.3, 0.4, 0.5, seed=1) != O.noise3(0.3, 0.4, 0.5, seed=2)


def test_generator_mapping_and_quantisation():
    """octaves=1, frequency tiny -> noise ~ 0 -> density 0.5 -> uint8(0.5*255) = 127 (truncation, engine L8 store)."""
    flat = O.noise_cubemap(abi.B200AtmoNoise(0, 1e-9, 1, 2.0, 0.5), 4, (1, 1, 1))
    assert (flat == 127).all()
    # direction mapping: a noise that depends on direction only must be seamless across faces -> the apron of the
    # seamless layout (copies of the NEIGHBOUR face's edge texels) continues each face as smoothly as its interior
    faces = O.noise_cubemap(NOISE, 32, (100, 200, 100))
    pad = O.cube_build_padded(faces).astype(int)
    interior_step = np.abs(np.diff(faces.astype(int), axis=2)).max()
    for apron, edge in ((pad[:, 1:-1, 0], pad[:, 1:-1, 1]), (pad[:, 1:-1, -1], pad[:, 1:-1, -2]),
                        (pad[:, 0, 1:-1], pad[:, 1, 1:-1]), (pad[:, -1, 1:-1], pad[:, -2, 1:-1])):
        assert np.abs(apron - edge).max() <= interior_step + 2
    # texel (x, y) of side s is the noise at cube_texel_directions * scale: check against a direct evaluation
    dirs = scenes.cube_texel_directions(8)
    one = O.noise_cubemap(abi.B200AtmoNoise(5, 0.02, 1, 2.0, 0.5), 8, (50, 60, 70))
    for s, y, x in ((0, 0, 0), (1, 3, 5), (2, 7, 1), (3, 2, 2), (4, 6, 6), (5, 1, 4)):
        d = np.float32(dirs[s, y, x])
        n = O.noise3(float(np.float32(d[0] * np.float32(50)) * np.float32(0.02)), float(np.float32(d[1] * np.float32(60)) * np.float32(0.02)),
                     float(np.float32(d[2] * np.float32(70)) * np.float32(0.02)), seed=5)
        assert abs(int(one[s, y, x]) - int((0.5 + 0.5 * n) * 255)) <= 1


def test_device_code_host_build_is_bit_exact():
    for res, scale in ((16, (100, 100, 100)), (33, (100, 200, 100))):
        want = O.noise_cubemap(NOISE, res, scale)
        got = np.empty_like(want)
        Hh.hostsim().hostsim_noise_cubemap(C.byref(NOISE), res, (C.c_float * 3)(*[float(v) for v in scale]), got.ctypes.data_as(C.c_void_p))
        assert np.array_equal(got, want)


def test_atlas_layout():
    faces = np.arange(6 * 2 * 2, dtype=np.uint8).reshape(6, 2, 2)
    atlas = O.cubemap_atlas(faces)
    assert atlas.shape == (4, 6)
    for side in range(6):
        x, y = side % 3, side // 3
        assert np.array_equal(atlas[2 * y:2 * y + 2, 2 * x:2 * x + 2], faces[side])


class FakeCtx:
    def __init__(self):
        self.calls = []

    def generate_noise_cubemap(self, noise, res, scale, download=True, set_as_coverage=False):
        self.calls.append((noise.seed, res, tuple(scale), download, set_as_coverage))
        return O.noise_cubemap(noise, res, scale) if download else None


def test_noise_cubemap_resource_mirror():
    ctx = FakeCtx()
    nc = NoiseCubemap(ctx)
    assert nc.resolution == 256 and nc.scale == (100.0, 100.0, 100.0)      # noise_cubemap.gd:25,38
    nc.resolution = 100000
    assert nc.resolution == 4096                                            # clampi(value, 1, 4096)
    nc.resolution = 0
    assert nc.resolution == 1
    nc.resolution = 8
    nc.scale = (100, 200, 100)
    assert ctx.calls == []                                                  # deferred: nothing generated yet
    fired = []
    nc.changed_callbacks.append(lambda: fired.append(1))
    nc.flush()
    assert len(ctx.calls) == 1 and ctx.calls[0][1:3] == (8, (100.0, 200.0, 100.0)) and fired == [1]
    nc.flush()
    assert len(ctx.calls) == 1                                              # nothing scheduled
    nc.noise.seed = 42
    nc.noise.emit_changed()                                                 # noise.changed -> _on_noise_changed
    im = nc.generate_importable_image()
    assert len(ctx.calls) == 2 and ctx.calls[1][0] == 42
    assert im.shape == (16, 24) and np.array_equal(im, O.cubemap_atlas(nc.get_faces()))
    assert np.array_equal(nc.get_layer_data(3), nc.get_faces()[3])
    nc.noise = FastNoiseLiteParams(seed=7)
    nc.bind_as_coverage()
    assert ctx.calls[-1] == (7, 8, (100.0, 200.0, 100.0), False, True)


@pytest.mark.gpu
@pytest.mark.parametrize("res,scale", [(1, (100, 100, 100)), (64, (100, 200, 100)), (256, (100, 200, 100))])
def test_gpu_generator_bit_exact(cuda_ctx_factory, res, scale):
    ctx = cuda_ctx_factory()
    got = ctx.generate_noise_cubemap(NOISE, res, scale)
    assert np.array_equal(got, O.noise_cubemap(NOISE, res, scale))


@pytest.mark.gpu
def test_gpu_generated_cube_installs_as_coverage(cuda_ctx_factory):
    ctx = cuda_ctx_factory()
    faces = ctx.generate_noise_cubemap(NOISE, 32, (100, 200, 100), download=True, set_as_coverage=True)
    assert np.array_equal(ctx.download_cube_padded(), O.cube_build_padded(faces))
    nc = NoiseCubemap(ctx, FastNoiseLiteParams(seed=1337))
    nc.resolution = 16
    assert np.array_equal(nc.get_faces(), O.noise_cubemap(abi.B200AtmoNoise(1337, 0.01, 5, 2.0, 0.5), 16, (100, 100, 100)))
