"""PINS THE ORACLE TO THE REFERENCE'S OWN SOURCE.

oracle/_ref/libatmo_ref.so is the reference's GDShader files (read from /root/reference at build time) compiled as C++
after a purely syntactic rewrite (oracle/ref/build_ref.py): every entry shader with its #defines, include tree and
vertex()/fragment() trampolines, plus the LUT bake shader. The hand-written oracle (oracle/atmo_oracle.hpp) must
reproduce it BIT FOR BIT — LUT, discard masks and fp32 RGBA — for every shipped shader, the BASELINE scale-ups, several
cameras and parameter sets. A mutation test shows the comparison has teeth.

What this does not pin (the reference leaves it to the engine / GPU): the rounding of GLSL built-ins and texture
filtering — both sides use the definitions documented in oracle/atmo_oracle.hpp.

Runs where the library can be built (/root/reference present) or was prebuilt (it travels with the repo snapshot)."""
import os
import shutil

import numpy as np
import pytest

from godot_atmosphere_shader_b200 import abi, scenes
from godot_atmosphere_shader_b200.planet_atmosphere import SHADER_VARIANTS
from oracle import pyoracle as O
from oracle import pyref as R
from tests import helpers as Hh

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref not built and /root/reference absent")


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def _assert_same(p, var, cam, tex, depth, w, h, shader=None, what=""):
    want, wdisc = R.render_frame(p, var, cam, tex, depth, w, h, shader=shader)
    got, gdisc = O.render_frame(p, var, cam, tex, depth, w, h)
    assert np.array_equal(gdisc, wdisc), f"{what}: discard masks differ"
    bad = _bits(got) != _bits(want)
    assert not bad.any(), f"{what}: {int(bad.sum())} of {bad.size} fp32 words differ, max |d| = {np.abs(got - want).max():.3e}"
    return want, wdisc


def test_entry_shader_defines_match_the_variant_tables():
    """The #defines of the 7 shipped entry shaders, read from the compiled sources, are what the node mirrors select."""
    entries = R.entry_shaders()
    assert set(entries) == set(SHADER_VARIANTS)
    for name, (lite, atmo_steps, cloud_steps, rm) in entries.items():
        model, ns, nc, light = SHADER_VARIANTS[name]
        assert model == (abi.SCATTER_V1 if lite else abi.SCATTER_V2), name
        assert (ns, nc) == (atmo_steps, cloud_steps), name
        assert light == (abi.LIGHT_NONE if cloud_steps == 0 else (abi.LIGHT_RAYMARCHED if rm else abi.LIGHT_CHEAP)), name


@pytest.mark.parametrize("params", ["demo", "template", "defaults", "thin"])
def test_lut_bake_bit_exact(params):
    p = {"demo": scenes.demo_params, "template": scenes.template_params, "defaults": abi.default_params,
         "thin": lambda: scenes.demo_params()}[params]()
    if params == "thin":
        p.planet_radius, p.atmosphere_height, p.density = 6371.0, 60.0, 0.013
    want = R.bake_lut(p)          # shader -> RGBA8 viewport -> FORMAT_RF
    got = O.bake_lut(p)
    assert np.array_equal(_bits(got), _bits(want))
    assert np.isfinite(want).all() and want.max() > 0


@pytest.mark.parametrize("shader", sorted(SHADER_VARIANTS))
@pytest.mark.parametrize("cam_name", ["A", "B"])
def test_every_shipped_shader_bit_exact(shader, cam_name):
    model, ns, nc, light = SHADER_VARIANTS[shader]
    w, h = 112, 63
    p = scenes.demo_params()
    if model == abi.SCATTER_V1:
        p.density = 0.02
    shape, cube, bn = Hh.demo_textures()
    tex = O.Textures(lut=O.bake_lut(p), shape=shape, cube_faces=cube, blue_noise=bn)
    cam = scenes.camera_a(w, h) if cam_name == "A" else scenes.camera_b(w, h, p)
    depth = scenes.synth_depth(cam, p, w, h)
    want, wdisc = _assert_same(p, O.variant(ns, nc, light, model), cam, tex, depth, w, h, shader=shader, what=shader)
    assert (wdisc == 0).any() and np.abs(want).max() > 0.01


@pytest.mark.parametrize("steps", [(32, 0, abi.LIGHT_NONE), (32, 64, abi.LIGHT_CHEAP), (8, 128, abi.LIGHT_RAYMARCHED), (64, 0, abi.LIGHT_NONE),
                                   (1, 1, abi.LIGHT_CHEAP)])
def test_baseline_scale_up_step_counts_bit_exact(steps):
    """BASELINE.json asks for 32 in-scatter / 128 cloud steps: the compiled reference runs them through the same code
    (the #define is a runtime variable there)."""
    ns, nc, light = steps
    w, h = 96, 54
    p = scenes.demo_params()
    shape, cube, bn = Hh.demo_textures()
    tex = O.Textures(lut=O.bake_lut(p), shape=shape, cube_faces=cube, blue_noise=bn)
    cam = scenes.camera_a(w, h, orbit_deg=40.0)
    depth = scenes.synth_depth(cam, p, w, h)
    _assert_same(p, O.variant(ns, nc, light), cam, tex, depth, w, h, what=str(steps))


def test_cameras_flags_and_parameter_corners_bit_exact():
    w, h = 80, 45
    shape, cube, bn = Hh.demo_textures()
    base = scenes.demo_params()
    cases = []
    # far away, inside the cloud layer, below ground level, looking away from the planet
    R0, H0 = base.planet_radius, base.atmosphere_height
    cases.append(("far", base, scenes.make_camera((0.0, 0.0, 1500.0), (0.0, 0.0, -1.0), aspect=w / h, far=4000.0)))
    cases.append(("in clouds", base, scenes.make_camera((0.0, R0 + 0.4 * H0, 0.0), (1.0, -0.1, 0.2), aspect=w / h)))
    cases.append(("underground", base, scenes.make_camera((0.0, R0 - 1.0, 0.0), (0.3, 1.0, 0.0), up=(0, 0, 1), aspect=w / h)))
    cases.append(("looking away", base, scenes.make_camera((0.0, 0.0, 157.0), (0.0, 0.2, 1.0), aspect=w / h)))
    q = base.copy()
    q.sphere_depth_factor = 0.6                       # planet_atmosphere_main.gdshaderinc:160
    cases.append(("sphere depth", q, scenes.camera_a(w, h)))
    q = base.copy()
    q.cloud_shape_invert, q.cloud_coverage_bias, q.cloud_blend = 0.0, 0.2, 0.9
    q.cloud_coverage_rotation[:] = (0.8, 0.6, -0.6, 0.8)
    cases.append(("cloud params", q, scenes.camera_b(w, h, q)))
    dp = scenes.camera_a(w, h)
    dp.double_precision = 1                            # `#ifdef DOUBLE_PRECISION`, main:118-125
    cases.append(("double precision", base, dp))
    moved = scenes.make_camera((30.0, 5.0, 150.0), (-0.2, 0.0, -1.0), aspect=w / h)
    T = np.eye(4)
    T[:3, 3] = (12.0, -3.0, 4.0)                       # MODEL_MATRIX: planet away from the origin
    moved.model[:] = scenes.flat_colmajor(T)
    q = base.copy()
    q.world_to_model[:] = scenes.flat_colmajor(np.linalg.inv(T))
    cases.append(("planet off origin", q, moved))
    for what, p, cam in cases:
        tex = O.Textures(lut=O.bake_lut(p), shape=shape, cube_faces=cube, blue_noise=bn)
        depth = scenes.synth_depth(cam, p, w, h, planet_center=(12.0, -3.0, 4.0) if what == "planet off origin" else (0.0, 0.0, 0.0))
        for var in (O.variant(8, 0, abi.LIGHT_NONE), O.variant(8, 32, abi.LIGHT_CHEAP), O.variant(8, 16, abi.LIGHT_RAYMARCHED)):
            _assert_same(p, var, cam, tex, depth, w, h, what=what)


def test_jitter_window_and_cloud_deck_camera_bit_exact():
    """(1) `ivec2(px) & ivec2(0xff)` (planet_atmosphere_main.gdshaderinc:168-169) with a 512 x 512 jitter texture and a frame
    wider than 256: the compiled shader reads only the top-left 256 x 256 texels, and so must the oracle. (2) camera C,
    looking down into the cloud deck (most pixels see cloud), for the cheap and the raymarched cloud light."""
    w, h = 300, 270
    p = scenes.demo_params()
    shape, cube, _ = Hh.demo_textures()
    big = np.random.default_rng(3).integers(0, 256, size=(512, 512), dtype=np.uint8)
    tex = O.Textures(lut=O.bake_lut(p), shape=shape, cube_faces=cube, blue_noise=big)
    cam = scenes.camera_b(w, h, p)
    depth = scenes.synth_depth(cam, p, w, h)
    want, _ = _assert_same(p, O.variant(8, 0, abi.LIGHT_NONE), cam, tex, depth, w, h, what="512^2 jitter texture")
    small = O.Textures(lut=tex.lut, shape=shape, cube_faces=cube, blue_noise=np.ascontiguousarray(big[:256, :256]))
    again, _ = O.render_frame(p, O.variant(8, 0, abi.LIGHT_NONE), cam, small, depth, w, h)
    assert np.array_equal(_bits(again), _bits(want))            # the other 3/4 of the texture are never read
    w, h = 96, 54
    cam = scenes.camera_c(w, h, p)
    depth = scenes.synth_depth(cam, p, w, h)
    tex = O.Textures(lut=O.bake_lut(p), shape=shape, cube_faces=cube, blue_noise=scenes.blue_noise_tile())
    plain, _ = O.render_frame(p, O.variant(8, 0, abi.LIGHT_NONE), cam, tex, depth, w, h)
    for var in (O.variant(8, 64, abi.LIGHT_CHEAP), O.variant(8, 64, abi.LIGHT_RAYMARCHED)):
        want, wdisc = _assert_same(p, var, cam, tex, depth, w, h, what="camera C")
        assert not wdisc.any() and (np.abs(want - plain).max(axis=-1) > 1e-6).mean() > 0.6


@pytest.mark.parametrize("seed", range(6))
def test_random_scenes_bit_exact(seed):
    """Random uniform blocks: planet scale over 3 decades, rotated + translated node, random cloud settings."""
    p, cam = Hh.random_scene(seed)
    w, h = 72, 48
    shape, cube, bn = Hh.demo_textures()
    tex = O.Textures(lut=O.bake_lut(p), shape=shape, cube_faces=cube, blue_noise=bn)
    depth = scenes.synth_depth(cam, p, w, h, planet_center=np.array(cam.model[:]).reshape(4, 4).T[:3, 3])
    for var in (O.variant(8, 0, abi.LIGHT_NONE), O.variant(8, 32, abi.LIGHT_CHEAP), O.variant(8, 16, abi.LIGHT_RAYMARCHED)):
        _assert_same(p, var, cam, tex, depth, w, h, what=f"seed {seed}")


@pytest.mark.parametrize("variant", [(8, 0, abi.LIGHT_NONE, abi.SCATTER_V2), (32, 0, abi.LIGHT_NONE, abi.SCATTER_V2), (8, 32, abi.LIGHT_CHEAP, abi.SCATTER_V2),
                                     (8, 16, abi.LIGHT_RAYMARCHED, abi.SCATTER_V2), (16, 32, abi.LIGHT_CHEAP, abi.SCATTER_V1)])
def test_fp64_twin_bit_exact(variant):
    """The oracle's T=double instantiation (the rounding-error bound quoted in every parity report) against the same
    reference sources compiled with `float` = double (oracle/_ref/libatmo_ref64.so): LUT and frames, bit for bit."""
    ns, nc, light, model = variant
    w, h = 80, 45
    p = scenes.demo_params()
    if model == abi.SCATTER_V1:
        p.density = 0.02
    lut64 = O.bake_lut(p, dtype=np.float64)
    assert np.array_equal(lut64.view(np.uint64), R.bake_lut(p, dtype=np.float64).view(np.uint64))
    shape, cube, bn = Hh.demo_textures()
    tex = O.Textures(lut=O.bake_lut(p), lut64=lut64, shape=shape, cube_faces=cube, blue_noise=bn)
    for cam in (scenes.camera_a(w, h, orbit_deg=10.0), scenes.camera_b(w, h, p)):
        depth = scenes.synth_depth(cam, p, w, h)
        var = O.variant(ns, nc, light, model)
        want, wdisc = R.render_frame(p, var, cam, tex, depth, w, h, dtype=np.float64)
        got, gdisc = O.render_frame(p, var, cam, tex, depth, w, h, dtype=np.float64)
        assert np.array_equal(gdisc, wdisc) and np.array_equal(got.view(np.uint64), want.view(np.uint64))
    # and without a double LUT both sides widen the fp32 one
    tex32 = O.Textures(lut=O.bake_lut(p), shape=shape, cube_faces=cube, blue_noise=bn)
    want, _ = R.render_frame(p, var, cam, tex32, depth, w, h, dtype=np.float64)
    got, _ = O.render_frame(p, var, cam, tex32, depth, w, h, dtype=np.float64)
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))


def test_unset_textures_bit_exact():
    """README.md:46: unset samplers cover the atmosphere uniformly (white) — both sides agree, and no blue noise = no jitter."""
    w, h = 64, 36
    p = scenes.demo_params()
    tex = O.Textures(lut=O.bake_lut(p))
    cam = scenes.camera_a(w, h)
    depth = scenes.synth_depth(cam, p, w, h)
    _assert_same(p, O.variant(8, 32, abi.LIGHT_CHEAP), cam, tex, depth, w, h, what="unset textures")


@pytest.mark.skipif(not R.reference_present(), reason="needs the reference tree to build a mutant")
def test_the_pin_has_teeth(tmp_path):
    """A copy of the reference with ONE constant changed (alpha jitter 0.02 -> 0.03, atmosphere_funcs_v2.gdshaderinc:96)
    compiles to a library the oracle no longer matches."""
    import ctypes as C

    from oracle.ref import build_ref
    mutant = tmp_path / "reference"
    shutil.copytree(os.path.join(R.REFERENCE, "addons", "zylann.atmosphere", "shaders"),
                    mutant / "addons" / "zylann.atmosphere" / "shaders")
    f = mutant / "addons" / "zylann.atmosphere" / "shaders" / "include" / "atmosphere_funcs_v2.gdshaderinc"
    src = f.read_text()
    assert src.count("jitter * 0.02") == 1
    f.write_text(src.replace("jitter * 0.02", "jitter * 0.03"))
    so = build_ref.build(str(mutant), verbose=False, out_dir=str(tmp_path / "out"))
    mlib = C.CDLL(so)
    w, h = 64, 36
    p = scenes.demo_params()
    _, _, bn = Hh.demo_textures()
    tex = O.Textures(lut=O.bake_lut(p), blue_noise=bn)
    cam = scenes.camera_b(w, h, p)
    depth = scenes.synth_depth(cam, p, w, h)
    var = O.variant(8, 0, abi.LIGHT_NONE)
    rgba = np.zeros((h, w, 4), np.float32)
    disc = np.zeros((h, w), np.uint8)
    ts = tex.struct()
    assert mlib.ref_render_frame(None, C.byref(p), C.byref(var), C.byref(cam), C.byref(ts), O._ptr(np.ascontiguousarray(depth, np.float32)),
                                     w, h, 0, h, 1, O._ptr(rgba), O._ptr(disc), 1) == 0
    got, _ = O.render_frame(p, var, cam, tex, depth, w, h)
    assert np.array_equal(_bits(got[..., :3]), _bits(rgba[..., :3]))          # colour untouched by the mutation
    assert (_bits(got[..., 3]) != _bits(rgba[..., 3])).mean() > 0.9            # alpha differs wherever jitter > 0


@pytest.mark.skipif(not R.reference_present(), reason="needs the reference tree")
def test_the_rewrite_is_syntax_only():
    """Token audit of oracle/ref/build_ref.py: for every entry shader and the bake shader, the multiset of numeric
    literals (values), identifiers and arithmetic operators of the rewritten text equals that of the original text, apart
    from the keywords the rewrite is documented to add or drop. No constant, name or operator appears or disappears."""
    import collections
    import re

    from oracle.ref import build_ref as B

    def strip_comments(t):
        t = re.sub(r"/\*.*?\*/", " ", t, flags=re.S)
        return re.sub(r"//[^\n]*", " ", t)

    tok = re.compile(r"[A-Za-z_]\w*|(?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?f?|0x[0-9a-fA-F]+|\d+u?|[-+*/<>=!?|^%\[\].&]")

    def tokens(t):
        out = collections.Counter()
        for x in tok.findall(strip_comments(t)):
            if re.fullmatch(r"(?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?f", x):
                x = x[:-1]                      # the f suffix is syntax; the VALUE must be unchanged
            out[x] += 1
        return out

    dropped_ok = {"uniform", "varying", "in", "out", "inout", "shader_type", "render_mode", "spatial", "canvas_item", "unshaded",
                  "blend_disabled", "source_color", "repeat_disable", "repeat_enable", "filter_nearest", "hint_depth_texture",
                  "discard", "ifdef", "endif", "DOUBLE_PRECISION", "height_curve"}
    added_ok = {"static", "int", "define", "if", "true", "return", "ref_discarded", "ref_double_precision", "height_curve_value",
                "ref_ATMOSPHERE_RAYMARCH_STEPS", "ref_CLOUDS_MAX_RAYMARCH_STEPS", "=", "&"}
    files = [n + ".gdshader" for n in B.ENTRY_SHADERS] + ["optical_depth.gdshader"]
    for f in files:
        path = os.path.join(R.REFERENCE, B.SHADERS, f)
        original = B.inline_includes(path, R.REFERENCE)
        rewritten, _ = B.rewrite(original)
        # inline_includes leaves '// >>> path' markers (comments) and removes the #include lines of the original files
        raw = ""
        seen = []

        def gather(p):
            for line in open(p, encoding="utf-8").read().splitlines():
                m = re.match(r'\s*#include\s+"([^"]+)"', line)
                if m:
                    gather(os.path.normpath(os.path.join(os.path.dirname(p), m.group(1))))
                else:
                    seen.append(line)
        gather(path)
        raw = "\n".join(seen)
        a, b = tokens(raw), tokens(rewritten)
        gone = a - b
        new = b - a
        assert set(gone) <= dropped_ok, f"{f}: tokens lost by the rewrite: {sorted(set(gone) - dropped_ok)}"
        assert set(new) <= added_ok, f"{f}: tokens introduced by the rewrite: {sorted(set(new) - added_ok)}"
        # every numeric literal survives with its value and multiplicity
        nums = lambda c: {k: v for k, v in c.items() if re.match(r"[\d.]", k) and k != "."}
        assert nums(a) == nums(b), f"{f}: numeric literals changed"
