"""Known-answer tests that pin the ORACLE (oracle/atmo_oracle.hpp) to the shader source.

The reference ships no tests or golden vectors: every expected value below is derived by
hand from the cited GDShader lines, so that the oracle — the thing every GPU parity test is measured against —
is itself checked against something other than itself. Paths are relative to addons/zylann.atmosphere/shaders/.
"""
import math

import numpy as np
import pytest

from godot_atmosphere_shader_b200 import abi, scenes
from oracle import pyoracle as O
from tests import helpers as Hh


# K1 — include/util.gdshaderinc:20-40
def test_k1_ray_sphere():
    assert O.ray_sphere((0, 0, -5), 1.0, (0, 0, 0), (0, 0, -1)) == (4.0, 6.0)
    assert O.ray_sphere((0, 5, -5), 1.0, (0, 0, 0), (0, 0, -1)) == (1000000.0, 1000000.0)  # miss sentinel, :36
    assert O.ray_sphere((0, 0, 0), 2.0, (0, 0, 0), (0, 0, -1)) == (-2.0, 2.0)  # origin at the centre
    # tangent ray: h == 0 -> both roots equal -> callers treat it as a miss (main:150)
    x, y = O.ray_sphere((0, 1, -5), 1.0, (0, 0, 0), (0, 0, -1))
    assert x == y == 5.0
    # sphere behind the origin: both roots negative
    x, y = O.ray_sphere((0, 0, 5), 1.0, (0, 0, 0), (0, 0, -1))
    assert (x, y) == (-6.0, -4.0)
    assert O.ray_sphere((0, 0, -5), 1.0, (0, 0, 0), (0, 0, -1), dtype=np.float64) == (4.0, 6.0)


# include/util.gdshaderinc:5-17 (used for the MODE_FAR proxy-cube coverage)
def test_ray_box_intersection():
    assert O.ray_box((0, 0, -5), (0, 0, 1), (1, 1, 1)) == (4.0, 6.0)
    assert O.ray_box((0, 0, 0), (0, 0, 1), (1, 2, 3)) == (-3.0, 3.0)            # origin inside
    assert O.ray_box((0, 5, -5), (1e-3, 1e-3, 1), (1, 1, 1)) == (-1.0, -1.0)    # miss sentinel vec2(-1.0)
    assert O.ray_box((0, 0, 5), (0, 0, 1), (1, 1, 1)) == (-1.0, -1.0)           # box behind the ray
    # known flaw of the routine, kept: an exactly axis-parallel ray OUTSIDE a slab gives -inf/NaN bounds; GPU min/max
    # ignore the NaN, so the slab is skipped and a hit is reported (measure-zero set of rays)
    assert O.ray_box((0, 5, -5), (0, 0, 1), (1, 1, 1)) == (4.0, 6.0)
    tn, tf = O.ray_box((-3, -3, -3), (1, 1, 1), (1, 1, 1))                       # un-normalised direction: t in its units
    assert (tn, tf) == (2.0, 4.0)


# K2 — include/atmosphere_common.gdshaderinc:12-24
def test_k2_density():
    p = abi.default_params()
    p.planet_radius, p.atmosphere_height, p.density = 10.0, 4.0, 3.0
    assert O.atmosphere_density(p, 10.0) == 3.0          # at the ground
    assert O.atmosphere_density(p, 12.0) == 3.0 / 8.0    # half way: (1/2)^3 * rho
    assert O.atmosphere_density(p, 14.0) == 0.0          # top
    assert O.atmosphere_density(p, 99.0) == 0.0          # above: clamps to exactly 0
    assert O.atmosphere_density(p, 5.0) == 3.0           # below ground: clamps to rho


# K3 — optical_depth.gdshader:17-31,45-69: a (nearly) straight-up ray is a left Riemann sum of a cubic
def test_k3_lut_vertical_ray_closed_form():
    p = scenes.demo_params()
    lut64 = O.bake_lut(p, dtype=np.float64)
    lut32 = O.bake_lut(p)
    R, H, rho = 100.0, 8.0, 0.5
    i = 255  # column nearest cos(theta)=+1: dy = 2*(255.5/256)-1
    dy = 2.0 * (i + 0.5) / 256 - 1.0
    dx = math.sqrt(1.0 - dy * dy)
    for j in (0, 17, 128, 255):
        v = (j + 0.5) / 256
        py = R + H * v
        # exact evaluation of the same 64-step sum in python floats (fp64)
        b = py * dy
        h = (R + H) ** 2 - ((py - b * dy) ** 2 + (b * dx) ** 2)
        ray_len = (-b + math.sqrt(h)) - max(-b - math.sqrt(h), 0.0)
        step = ray_len / 64
        od = 0.0
        for s in range(64):
            x, y = dx * step * s, py + dy * step * s
            hh = min(max((math.hypot(x, y) - R) / H, 0.0), 1.0)
            od += (1 - hh) ** 3 * rho * step * rho
        assert lut64[j, i] == pytest.approx(od, rel=1e-12)
        assert lut32[j, i] == pytest.approx(od, rel=2e-6)
        # and the closed form of SURVEY §8(c) K3 for a truly vertical ray bounds it within the tilt of column 255
        closed = rho * rho * H * (1 - v) ** 4 * (2080.0 ** 2 / 64.0 ** 4)
        assert lut64[j, i] == pytest.approx(closed, rel=2e-2)
    assert 2080.0 ** 2 / 64.0 ** 4 == 0.25787353515625


def test_lut_layout_and_rgba8_roundtrip():
    """Row = height ratio, column = 0.5+0.5*cos(theta) (optical_depth.gdshader:12-15); the RGBA8 viewport encoding
    (:33-43) reinterpreted as FORMAT_RF (optical_depth_baker.gd:75-77) is a lossless fp32 round trip."""
    p = scenes.demo_params()
    lut = O.bake_lut(p)
    assert lut.shape == (256, 256) and np.isfinite(lut).all() and (lut >= 0).all()
    assert (np.diff(lut[:, 200]) <= 1e-7).all()      # optical depth decreases with height (looking up)
    assert lut[10, 5] > lut[10, 250]                 # looking down through the planet's shell is thicker than looking up
    assert np.array_equal(O.bake_lut(p, via_rgba8=True).view(np.uint32), lut.view(np.uint32))
    for f in (0.0, 1.0, 50.831257, 7.5e-12, 3.4e38):
        b = O.encode_float(f)
        assert b == np.float32(f).tobytes() and O.decode_float(b) == float(np.float32(f))
    # LUT range for the demo parameters
    assert 1e-12 < lut.min() < 1e-8 and 45.0 < lut.max() < 55.0


def test_lut_sampling_convention():
    """texture() on the LUT: texel centres at (i+0.5)/256, clamp to edge, a+(b-a)*t lerps."""
    lut = np.arange(256 * 256, dtype=np.float32).reshape(256, 256)
    assert O.sample_lut(lut, 0.5 / 256, 0.5 / 256) == 0.0
    assert O.sample_lut(lut, 10.5 / 256, 3.5 / 256) == 3 * 256 + 10
    assert O.sample_lut(lut, 11.0 / 256, 3.5 / 256) == 3 * 256 + 10.5
    assert O.sample_lut(lut, 10.5 / 256, 4.0 / 256) == 3.5 * 256 + 10
    assert O.sample_lut(lut, 0.0, 0.0) == 0.0                # clamp to edge
    assert O.sample_lut(lut, 1.0, 1.0) == 255 * 256 + 255
    assert O.sample_lut(lut, -3.0, 7.0) == 255 * 256 + 0


# K4 / K5 — include/atmosphere_funcs_v2.gdshaderinc:32-101
def test_k4_zero_length_interval():
    p = scenes.demo_params()
    lut = O.bake_lut(p)
    rgba = O.compute_atmosphere_v2(p, lut, 8, (0, 0, 0), (0, 0, -1), (0, 0, -150), 42.0, 42.0, (0, 0, 1), 0.5)
    amb, mod = p.atmosphere_ambient_color, p.atmosphere_modulate
    for c in range(3):
        assert rgba[c] == np.float32(np.float32(amb[c]) * np.float32(mod[c]))
    assert rgba[3] == np.float32(np.float32(0.5) * np.float32(0.02))


def test_k5_alpha_is_one_minus_exp_of_view_optical_depth():
    p = scenes.demo_params()
    lut = O.bake_lut(p)
    C = (0.0, 0.0, -150.0)
    t0, t1 = O.ray_sphere(C, 108.0, (0, 0, 0), (0.2, 0.1, -0.97))
    d = np.array([0.2, 0.1, -0.97]); d /= np.linalg.norm(d)
    t0, t1 = O.ray_sphere(C, 108.0, (0, 0, 0), d)
    for n in (8, 32):
        rgba = O.compute_atmosphere_v2(p, lut, n, (0, 0, 0), d, C, t0, t1, (0, 0, 1), 0.0)
        step = (t1 - t0) / n
        od = 0.0
        for i in range(n):
            pos = d * (t0 + i * step) - np.array(C)
            h = min(max((np.linalg.norm(pos) - 100.0) / 8.0, 0.0), 1.0)
            od += (1 - h) ** 3 * 0.5 * 0.5 * step
        assert rgba[3] == pytest.approx(min(1 - math.exp(-od), 0.99), rel=2e-5)
    # jitter adds 0.02*jitter, clamped to 0.99 (:96)
    a0 = O.compute_atmosphere_v2(p, lut, 8, (0, 0, 0), d, C, t0, t1, (0, 0, 1), 0.0)[3]
    a1 = O.compute_atmosphere_v2(p, lut, 8, (0, 0, 0), d, C, t0, t1, (0, 0, 1), 1.0)[3]
    assert a1 == pytest.approx(min(a0 + 0.02, 0.99), abs=1e-7)


def test_scattering_coefficients():
    """(400/lambda)^4 * strength, funcs_v2:47-51; SURVEY §8(b2): (0.106622, 0.324442, 0.683013) for (700,530,440)."""
    assert (400 / 700) ** 4 == pytest.approx(0.106622, rel=1e-5)
    assert (400 / 530) ** 4 == pytest.approx(0.324442, rel=1e-5)
    assert (400 / 440) ** 4 == pytest.approx(0.683013, rel=1e-5)
    # a single thin sample: L = ld*step*exp(-(sun_od+view_od)*coef)*coef — channel ratios follow the coefficients
    p = scenes.demo_params()
    p.atmosphere_ambient_color[:] = (0, 0, 0)
    p.atmosphere_modulate[:] = (1, 1, 1)
    lut = np.zeros((256, 256), np.float32)  # no sun attenuation
    C = (0.0, 0.0, -104.0)
    rgba = O.compute_atmosphere_v2(p, lut, 1, (0, 0, 0), (0, 0, -1), C, 0.0, 1e-3, (0, 0, 1), 0.0)
    ld_step = (1 - 0.5) ** 3 * 0.25 * 1e-3
    for c, k in enumerate([(400 / 700) ** 4, (400 / 530) ** 4, (400 / 440) ** 4]):
        assert rgba[c] == pytest.approx(ld_step * k * math.exp(-ld_step * k), rel=1e-5)


# K6 — include/cloud_funcs.gdshaderinc:25-68
def test_k6_cloud_density():
    p = scenes.demo_params()  # bottom 101.6, top 104.8
    shape, cube, _ = Hh.demo_textures()
    tex = O.Textures(lut=O.bake_lut(p), shape=shape, cube_faces=cube)
    rng = np.random.default_rng(0)
    for _ in range(200):
        u = rng.normal(size=3); u /= np.linalg.norm(u)
        for r in (99.0, 101.0, 101.59, 104.81, 110.0):   # outside the shell: exactly 0 (height curve clamps)
            assert O.cloud_density(p, tex, u * r) == 0.0
    # constant textures: closed form, filter independent
    for cov8, shp8 in ((255, 255), (128, 64), (40, 200)):
        t2 = O.Textures(lut=tex.lut, shape=np.full((4, 4, 4), shp8, np.uint8), cube_faces=np.full((6, 4, 4), cov8, np.uint8))
        for r in (102.0, 103.2, 104.5):
            pos = np.array([0.3, 0.5, 0.81]); pos = pos / np.linalg.norm(pos) * r
            hr = (np.linalg.norm(np.float32(pos).astype(np.float64)) - 101.6) / 3.2
            hc = max(1 - (2 * hr - 1) ** 2, 0.0)
            cov = cov8 / 255 - 0.25 * hr + 0.0
            shp = 1.0 - (0.5 * (1 - 0.5) + (shp8 / 255) * 0.5)          # mix(0.5, tex, 0.5), inverted (demo: invert=1)
            want = min(max(((shp - 0.1) + (-1.2 * (1 - cov) + 1.5 * cov)) * hc * 50 - 20, 0.0), 1.0)
            assert O.cloud_density(p, t2, pos) == pytest.approx(want, abs=2e-3)
    # unset samplers read white (README.md:46)
    t3 = O.Textures(lut=tex.lut)
    t4 = O.Textures(lut=tex.lut, shape=np.full((2, 2, 2), 255, np.uint8), cube_faces=np.full((6, 2, 2), 255, np.uint8))
    pos = (0.0, 103.0, 0.5)
    assert O.cloud_density(p, t3, pos) == O.cloud_density(p, t4, pos)


def test_shape_texture_repeat_and_trilinear():
    p = scenes.demo_params()
    n = 4
    shape = np.zeros((n, n, n), np.uint8)
    shape[1, 2, 3] = 255
    tex = O.Textures(lut=O.bake_lut(p), shape=shape)
    centre = ((3 + 0.5) / n, (2 + 0.5) / n, (1 + 0.5) / n)
    assert O.sample_shape(p, tex, centre) == 1.0
    assert O.sample_shape(p, tex, (centre[0] + 5, centre[1] - 3, centre[2] + 1)) == 1.0      # repeat_enable
    assert O.sample_shape(p, tex, (centre[0] + 0.5 / n, centre[1], centre[2])) == pytest.approx(0.5)  # wraps to x=0
    assert O.sample_shape(p, tex, ((0 + 0.5) / n, centre[1], centre[2])) == 0.0
    assert O.sample_shape(p, tex, (centre[0] + 0.25 / n, centre[1] + 0.25 / n, centre[2])) == pytest.approx(0.75 * 0.75)


def test_cubemap_convention_matches_generator():
    """Face order / orientation of noise_cubemap.gd:110-128: sampling the direction of a texel centre returns that texel."""
    p = scenes.demo_params()
    res = 8
    rng = np.random.default_rng(3)
    faces = rng.integers(0, 256, size=(6, res, res), dtype=np.uint8)
    tex = O.Textures(lut=O.bake_lut(p), cube_faces=faces)
    dirs = scenes.cube_texel_directions(res)
    for f in range(6):
        for y in range(res):
            for x in range(res):
                got = O.sample_cube(p, tex, dirs[f, y, x] * 3.7)  # direction need not be normalised
                assert got == pytest.approx(faces[f, y, x] / 255.0, abs=2e-6), (f, y, x)


def test_cubemap_seamless_across_edges_and_corners():
    p = scenes.demo_params()
    res = 4
    rng = np.random.default_rng(4)
    faces = rng.integers(0, 256, size=(6, res, res), dtype=np.uint8)
    tex = O.Textures(lut=O.bake_lut(p), cube_faces=faces)
    eps = 1e-4
    # continuity across every edge type: approach the x=y edge, the x=z edge and the y=z edge from both sides
    for t in np.linspace(-0.9, 0.9, 7):
        for a, b in (((1.0, 1.0 - eps, t), (1.0 - eps, 1.0, t)), ((1.0, t, 1.0 - eps), (1.0 - eps, t, 1.0)),
                     ((t, 1.0, 1.0 - eps), (t, 1.0 - eps, 1.0)), ((-1.0, t, -1.0 + eps), (-1.0 + eps, t, -1.0))):
            assert O.sample_cube(p, tex, a) == pytest.approx(O.sample_cube(p, tex, b), abs=2e-3)
    # at a cube corner all three faces give the same value: the mean of the three corner texels (+ the rounded mean apron)
    c = [O.sample_cube(p, tex, v) for v in ((1.0, 1.0 - 1e-6, 1.0 - 1e-6), (1.0 - 1e-6, 1.0, 1.0 - 1e-6), (1.0 - 1e-6, 1.0 - 1e-6, 1.0))]
    assert max(c) - min(c) < 1e-4
    # the padded layout itself: interior preserved, apron texels are copies of real texels, corner = rounded mean
    pad = O.cube_build_padded(faces)
    assert np.array_equal(pad[:, 1:-1, 1:-1], faces)
    allv = set(faces.reshape(-1).tolist())
    for f in range(6):
        for k in range(1, res + 1):
            for v in (pad[f, 0, k], pad[f, -1, k], pad[f, k, 0], pad[f, k, -1]):
                assert int(v) in allv
        a, b, cc = int(pad[f, 0, 1]), int(pad[f, 1, 0]), int(pad[f, 1, 1])
        assert int(pad[f, 0, 0]) == (2 * (a + b + cc) + 3) // 6
    # each edge apron row equals the adjacent face's edge row/column (possibly reversed): check one known pair
    # +X face right apron (s>1) folds onto -Z (noise_cubemap.gd: +X s axis is -z): +X[j][res] == -Z[j][0]
    assert np.array_equal(pad[0, 1:-1, -1], faces[5, :, 0])
    assert np.array_equal(pad[0, 1:-1, 0], faces[4, :, -1])   # +X left apron == +Z right column


# K7 — include/util.gdshaderinc:61-69
def test_k7_blend_colors():
    x = (0.2, 0.4, 0.6, 0.5)
    assert O.blend_colors(x, (0.9, 0.9, 0.9, 0.0)) == pytest.approx(x, rel=1e-6)
    assert O.blend_colors((0.3, 0.3, 0.3, 0.0), (0.9, 0.9, 0.9, 0.0)) == (0.0, 0.0, 0.0, 0.0)
    r = O.blend_colors((1.0, 0.0, 0.0, 1.0), (0.0, 1.0, 0.0, 0.25))
    assert r == pytest.approx((0.75, 0.25, 0.0, 1.0))


# K8 — include/cloud_funcs.gdshaderinc:186-204 and :108 for the demo parameters
def test_k8_march_distance_cap_and_light_reach():
    p = scenes.demo_params()
    bottom, top, ground = 100 + 0.2 * 8, 100 + 0.6 * 8, 100.0
    space = 0.5 * math.sqrt(1 - (ground / top) ** 2) * bottom
    assert space == pytest.approx(15.19805544, rel=1e-6)
    assert 3 * space == pytest.approx(45.59416633, rel=1e-6)
    assert (top - bottom) * 0.15 == pytest.approx(0.48)
    # observable effect: with uniform unit-density clouds the march length is capped at `space` for an origin far above
    # the shell, so alpha = 1 - exp(-density_scale * min(length, space)) no matter how long the requested interval is
    tex = O.Textures(lut=O.bake_lut(p))  # unset samplers: white -> density clamps to 1 inside the shell
    o = (0.0, 0.0, 300.0)  # far outside: smoothstep(...) = 1 -> max_d = march_distance_space
    d = (0.0, 0.0, -1.0)
    # a ray straight down through the shell: t in [300-104.8, 300-101.6], requested interval much longer than the cap
    l1, a1 = O.raymarch_cloud(p, tex, 64, abi.LIGHT_CHEAP, o, d, 195.2, 195.2 + 100.0, 0.0, (0, 0, 1))
    l2, a2 = O.raymarch_cloud(p, tex, 64, abi.LIGHT_CHEAP, o, d, 195.2, 195.2 + space, 0.0, (0, 0, 1))
    assert (l1, a1) == (l2, a2)
    # the shell is 3.2 thick and density_scale = 2: steps of space/64 sample ~13.5 points inside at density<=1
    assert 0.9 < a1 <= 1.0


def test_cheap_light_pow16_and_planet_shadow():
    p = scenes.demo_params()
    tex = O.Textures(lut=O.bake_lut(p))
    pos = (0.0, 103.2, 0.0)  # height ratio 0.5
    # sun behind the viewer direction (dp < 0): pow(dp,16) term defined as 0; shadow: dot(up,-sun) = -1 -> smoothstep = 0
    l = O.cloud_light(p, tex, abi.LIGHT_CHEAP, pos, (0, 0, -1), (0, 1, 0), 0.0, 0.0)
    assert l == pytest.approx(0.5, rel=1e-5)
    # looking straight at the sun: + 1^16 * (1 - alpha)
    l = O.cloud_light(p, tex, abi.LIGHT_CHEAP, pos, (0, 1, 0), (0, 1, 0), 0.0, 0.25)
    assert l == pytest.approx(0.5 + 0.75, rel=1e-5)
    # night side: dot(normalize(pos), -sun) = 1 -> smoothstep = 1 -> light * 0.002  (:87,:164)
    l = O.cloud_light(p, tex, abi.LIGHT_CHEAP, pos, (0, 0, -1), (0, -1, 0), 0.0, 0.0)
    assert l == pytest.approx(0.5 * 0.002, rel=1e-4)


def test_raymarched_light_clear_sky_is_full_light():
    p = scenes.demo_params()
    zero = O.Textures(lut=O.bake_lut(p), shape=np.full((2, 2, 2), 255, np.uint8), cube_faces=np.zeros((6, 2, 2), np.uint8))
    # coverage 0 -> density 0 everywhere -> alpha 0 -> mix(1, light0, 0) = 1 (then the planet shadow factor)
    l = O.cloud_light(p, zero, abi.LIGHT_RAYMARCHED, (0.0, 103.2, 0.0), (0, 0, -1), (0, 1, 0), 0.0, 0.0)
    assert l == pytest.approx(1.0, rel=1e-6)


# K9 — SURVEY §8(c): magnitudes of the demo view (survey-time scratch values, 64^2 LUT proxy): same ballpark
def test_k9_demo_view_magnitudes():
    p = scenes.demo_params()
    lut = O.bake_lut(p)
    eye_to_centre = 157.92054
    C = (0.0, 0.0, -eye_to_centre)
    sun_dir = (0.0, 0.0, 1.0)  # sun behind the camera
    t0, t1 = O.ray_sphere(C, 108.0, (0, 0, 0), (0, 0, -1))
    g0, _ = O.ray_sphere(C, 100.0, (0, 0, 0), (0, 0, -1))
    r8 = O.compute_atmosphere_v2(p, lut, 8, (0, 0, 0), (0, 0, -1), C, t0, min(t1, g0), sun_dir, 0.5)
    r32 = O.compute_atmosphere_v2(p, lut, 32, (0, 0, 0), (0, 0, -1), C, t0, min(t1, g0), sun_dir, 0.5)
    assert r8 == pytest.approx((0.0585, 0.1249, 0.2303, 0.3281), rel=0.05)
    assert r32 == pytest.approx((0.0671, 0.1468, 0.2660, 0.3845), rel=0.05)


def test_fragment_quirks_are_replicated():
    """main:150 tangent/miss -> discard; main:160-162: opaque geometry nearer than the atmosphere entry gives a
    NEGATIVE march interval and the loop still runs; both oracles and the kernels keep that."""
    p = scenes.demo_params()
    tex = O.Textures(lut=O.bake_lut(p))
    fr = abi.B200AtmoFrame()
    fr.planet_center_view[:] = (0.0, 0.0, -300.0)
    fr.sun_center_view[:] = (0.0, 0.0, 5000.0)
    fr.inv_view[:] = abi.IDENTITY16
    od = np.array([[0, 0, 0, 50.0],      # depth 50 < atmosphere entry 192 -> t_end < t_begin
                   [0, 0, 0, 200.0],     # stopped by the ground (a sample AT the planet centre would be normalize(0) = NaN)
                   [0, 0, 0, 1e4]], np.float32)
    dj = np.array([[0, 0, -1, 0.5], [0, 0, -1, 0.5], [0, 1, 0, 0.5]], np.float32)
    rgba, disc = O.render_rays(p, O.variant(8), fr, tex, od, dj)
    assert disc.tolist() == [0, 0, 1]
    amb = np.float32(p.atmosphere_ambient_color[:]) * np.float32(p.atmosphere_modulate[:])
    assert rgba[0, :3] == pytest.approx(amb, rel=1e-4)            # outside the shell density is 0: ambient only
    assert rgba[0, 3] == pytest.approx(0.01, abs=1e-6)            # alpha = 0.02*jitter
    assert (rgba[1, :3] > amb).all() and rgba[1, 3] > 0.3
    assert (rgba[2] == 0).all()


def test_oracle_multithreading_is_deterministic():
    p = scenes.demo_params()
    shape, cube, bn = Hh.demo_textures()
    tex = O.Textures(lut=O.bake_lut(p), shape=shape, cube_faces=cube, blue_noise=bn)
    od, dj, fr = Hh.random_rays(30000, p, seed=2)
    a, da = O.render_rays(p, O.variant(8, 32, abi.LIGHT_CHEAP), fr, tex, od, dj, threads=1)
    b, db = O.render_rays(p, O.variant(8, 32, abi.LIGHT_CHEAP), fr, tex, od, dj, threads=0)
    assert np.array_equal(a, b) and np.array_equal(da, db)
