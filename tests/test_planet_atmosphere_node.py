"""Host logic of the PlanetAtmosphere / OpticalDepthBaker mirror (planet_atmosphere.gd, optical_depth_baker.gd),
with a recording stand-in for the device context; the GPU variant at the bottom drives the real C-ABI."""
import math
import warnings

import numpy as np
import pytest

from godot_atmosphere_shader_b200 import abi, scenes
from godot_atmosphere_shader_b200.planet_atmosphere import (MODE_FAR, MODE_NEAR, SHADER_VARIANTS, OpticalDepthBaker,
                                                            PlanetAtmosphere, srgb_to_linear)


class FakeCtx:
    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        def rec(*a, **k):
            self.calls.append((name, a, k))
        return rec

    def count(self, name):
        return sum(1 for c in self.calls if c[0] == name)


def make_node():
    ctx = FakeCtx()
    return PlanetAtmosphere(ctx=ctx), ctx


def test_defaults_and_init():
    n, ctx = make_node()
    assert (n.planet_radius, n.atmosphere_height) == (1.0, 0.1)          # planet_atmosphere.gd:20,28
    assert tuple(n.params.sun_position) == (5000.0, 0.0, 0.0)            # :106
    assert n.params.clip_mode == 0.0 and n.mode == MODE_FAR              # :108, :58
    assert n.clouds_rotation_speed == 1.0 and n.force_fullscreen is False
    assert n.extra_cull_margin == pytest.approx(1.1)                     # :241-242
    assert ctx.count("upload_blue_noise") == 1                           # :107
    assert ("set_variant", (8, 0, abi.LIGHT_NONE, abi.SCATTER_V2), {}) in ctx.calls   # default shader :13-14


def test_setters_clamp_and_trigger_rebake():
    n, ctx = make_node()
    n.set_custom_shader("planet_atmosphere_no_clouds.gdshader")          # has u_optical_depth_texture -> baking on
    baker = n._optical_depth_baker
    assert baker is not None and baker._state == OpticalDepthBaker.STATE_REQUEST_BAKE
    n._process(0.016)   # frame 1: _setup_bake
    assert baker._state == OpticalDepthBaker.STATE_PENDING_RENDER and ctx.count("bake_optical_depth") == 1
    assert not n._optical_depth_ready
    n._process(0.016)   # frame 2: baked signal
    assert baker._state == OpticalDepthBaker.STATE_IDLE and n._optical_depth_ready and not baker.processing
    n.planet_radius = -5.0                                               # maxf(new_radius, 0.0) :233
    assert n.planet_radius == 0.0 and baker._state == OpticalDepthBaker.STATE_REQUEST_BAKE
    n._process(); n._process()
    n.set_atmosphere_height(0.3)
    assert n.atmosphere_height == 0.3 and n.extra_cull_margin == pytest.approx(0.3)
    assert baker._state == OpticalDepthBaker.STATE_REQUEST_BAKE
    n._process(); n._process()
    bakes = ctx.count("bake_optical_depth")
    n.set_atmosphere_height(0.3)                                         # unchanged: early return :246-247
    assert baker._state == OpticalDepthBaker.STATE_IDLE
    n.set("shader_params/u_scattering_strength", 3.0)                    # not in _shader_params_affecting_optical_depth
    assert baker._state == OpticalDepthBaker.STATE_IDLE
    n.set("shader_params/u_density", 0.7)                                # :217-218
    assert baker._state == OpticalDepthBaker.STATE_REQUEST_BAKE and n.params.density == pytest.approx(0.7)
    n._process(); n._process()
    assert ctx.count("bake_optical_depth") == bakes + 1


def test_shader_params_surface():
    n, _ = make_node()
    names = {p["name"] for p in n._get_property_list()}
    assert "shader_params/u_density" in names and "shader_params/u_scattering_wavelengths" in names
    assert not any(x.endswith(("u_planet_radius", "u_sun_position", "u_optical_depth_texture", "u_clip_mode")) for x in names)
    assert not any("cloud" in x for x in names)                          # no clouds in the default shader
    n.set_custom_shader("planet_atmosphere_clouds_high")
    names = {p["name"] for p in n._get_property_list()}
    assert "shader_params/u_cloud_density_scale" in names and "shader_params/u_cloud_coverage_cubemap" in names
    assert "shader_params/u_cloud_coverage_rotation" not in names and "shader_params/u_world_to_model_matrix" not in names
    n.set_custom_shader("planet_atmosphere_v1_clouds")
    names = {p["name"] for p in n._get_property_list()}
    assert "shader_params/u_day_color0" in names and "shader_params/u_scattering_strength" not in names
    # _get falls back to the shader default when unset (:206-207)
    assert n.get("shader_params/u_cloud_top") == 0.5
    n.set("shader_params/u_cloud_top", 0.6)
    assert n.get("shader_params/u_cloud_top") == 0.6 and n.params.cloud_top == pytest.approx(0.6)
    # source_color uniforms are converted sRGB -> linear before upload; get returns what was set
    n.set_custom_shader("planet_atmosphere_no_clouds")
    n.set_shader_parameter("u_atmosphere_modulate", (1.0, 0.5, 0.0))
    assert n.get_shader_parameter("u_atmosphere_modulate") == (1.0, 0.5, 0.0)
    assert list(n.params.atmosphere_modulate) == pytest.approx([1.0, float(srgb_to_linear(0.5)), 0.0])
    assert float(srgb_to_linear(0.5)) == pytest.approx(0.21404114)


def test_deprecated_accessors_warn():
    n, _ = make_node()
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        n.set_shader_param("u_density", 0.4)
        assert n.get_shader_param("u_density") == 0.4
    assert len(w) == 2 and "deprecated" in str(w[0].message)


def test_variant_table_matches_entry_shaders():
    assert SHADER_VARIANTS["planet_atmosphere_no_clouds"] == (abi.SCATTER_V2, 8, 0, abi.LIGHT_NONE)
    assert SHADER_VARIANTS["planet_atmosphere_clouds"] == (abi.SCATTER_V2, 8, 32, abi.LIGHT_CHEAP)
    assert SHADER_VARIANTS["planet_atmosphere_clouds_high"] == (abi.SCATTER_V2, 8, 64, abi.LIGHT_CHEAP)
    assert SHADER_VARIANTS["planet_atmosphere_clouds_high_rm"] == (abi.SCATTER_V2, 8, 64, abi.LIGHT_RAYMARCHED)
    assert SHADER_VARIANTS["planet_atmosphere_v1_no_clouds"][:2] == (abi.SCATTER_V1, 16)
    n, ctx = make_node()
    n.custom_shader = "res://addons/zylann.atmosphere/shaders/planet_atmosphere_clouds_high_m.gdshader"  # README spelling
    assert ctx.calls[-2] == ("set_variant", (8, 64, abi.LIGHT_RAYMARCHED, abi.SCATTER_V2), {}) or \
        ("set_variant", (8, 64, abi.LIGHT_RAYMARCHED, abi.SCATTER_V2), {}) in ctx.calls
    n.custom_shader = (abi.SCATTER_V2, 32, 128, abi.LIGHT_RAYMARCHED)   # BASELINE scale-up step counts
    assert ("set_variant", (32, 128, abi.LIGHT_RAYMARCHED, abi.SCATTER_V2), {}) in ctx.calls
    with pytest.raises(ValueError):
        n.custom_shader = "no_such_shader"
    n.custom_shader = None
    assert n._variant() == SHADER_VARIANTS["planet_atmosphere_no_clouds"]


def test_process_mode_switch_sun_and_rotation():
    n, _ = make_node()
    n.planet_radius, n.atmosphere_height = 100.0, 8.0
    clip = 1.75 * (100 + 8 + 0.1) * 1.1
    n._process(0.0, camera_position=(0, 0, clip * 1.01), camera_near=0.1, now=0.0)
    assert n.mode == MODE_FAR and n.params.clip_mode == 0.0 and n._far_mesh_size == pytest.approx(clip)
    n._process(0.0, camera_position=(0, 0, clip * 0.99), camera_near=0.1, now=0.0)
    assert n.mode == MODE_NEAR and n.params.clip_mode == 1.0
    n._process(0.0, camera_position=(0, 0, 1e6), now=0.0)
    assert n.mode == MODE_FAR
    n.force_fullscreen = True
    n._process(0.0, camera_position=(0, 0, 1e6), now=0.0)
    assert n.mode == MODE_NEAR
    assert n._get_configuration_warnings() == ["The path to the sun is not assigned."]

    class Sun:
        global_transform = np.eye(4)
    Sun.global_transform[:3, 3] = (1.0, 2.0, 478.677)
    n.sun_path = Sun
    assert n._get_configuration_warnings() == []
    T = np.eye(4); T[:3, 3] = (10.0, 0.0, 0.0)
    n.global_transform = T
    n.clouds_rotation_speed = 90.0
    n._process(0.0, camera_position=(0, 0, 0), now=1.0)                  # t = 1 s -> 90 degrees
    assert tuple(n.params.sun_position) == pytest.approx((1.0, 2.0, 478.677))
    w2m = np.array(n.params.world_to_model[:]).reshape(4, 4).T
    assert w2m[:3, 3] == pytest.approx([-10.0, 0.0, 0.0])                # global_transform.inverse()
    c0x, c0y, c1x, c1y = n.params.cloud_coverage_rotation[:]
    assert (c0x, c0y, c1x, c1y) == pytest.approx((0.0, 1.0, -1.0, 0.0), abs=1e-6)  # Transform2D().rotated(pi/2)


@pytest.mark.gpu
def test_node_renders_through_the_c_abi():
    import torch

    from oracle import pyoracle as O
    from tests import helpers as Hh
    n = PlanetAtmosphere(0)
    try:
        n.planet_radius, n.atmosphere_height = 100.0, 8.0
        n._ready()
        n.custom_shader = "planet_atmosphere_clouds"
        demo = scenes.demo_params()
        for u, f in (("u_density", "density"), ("u_scattering_strength", "scattering_strength"),
                     ("u_cloud_density_scale", "cloud_density_scale"), ("u_cloud_top", "cloud_top"),
                     ("u_cloud_shape_invert", "cloud_shape_invert"), ("u_cloud_shape_factor", "cloud_shape_factor"),
                     ("u_cloud_shape_scale", "cloud_shape_scale")):
            n.set(f"shader_params/{u}", getattr(demo, f))
        shape, cube, bn = Hh.demo_textures()
        n.set("shader_params/u_cloud_shape_texture", shape)
        n.set("shader_params/u_cloud_coverage_cubemap", cube)
        n.sun_path = np.array(demo.sun_position[:])
        w, h = 160, 90
        cam0 = scenes.camera_a(w, h)
        for _ in range(2):
            n._process(0.016, camera_position=cam0._meta["inv_view"][:3, 3], now=0.0)
        cam = n.make_camera(np.linalg.inv(cam0._meta["P"]), cam0._meta["inv_view"])
        depth = scenes.synth_depth(cam0, demo, w, h)
        rgba = torch.empty((h, w, 4), dtype=torch.float32, device="cuda")
        disc = torch.empty((h, w), dtype=torch.uint8, device="cuda")
        n.render(cam, torch.from_numpy(depth).cuda(), w, h, rgba, disc)
        torch.cuda.synchronize()
        p = n.params
        tex = O.Textures(lut=O.bake_lut(p), shape=shape, cube_faces=cube, blue_noise=scenes.blue_noise_tile())
        ref, rdisc = O.render_frame(p, O.variant(8, 32, abi.LIGHT_CHEAP), cam, tex, depth, w, h, threads=0)
        assert np.array_equal(disc.cpu().numpy(), rdisc)
        Hh.assert_rgba_close(rgba.cpu().numpy(), ref, what="node")
    finally:
        n.free()
