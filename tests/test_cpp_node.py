"""C++ core of the PlanetAtmosphere node (godot_atmosphere_shader_b200/csrc/node/): host logic on CPU with a recording
stand-in of the C-ABI table (tests/cpp/test_planet_atmosphere_node.cpp), and — on the GPU — one frame rendered through
the C++ node, compared bit for bit with the Python mirror of the same node and, within tolerance, with the oracle."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from godot_atmosphere_shader_b200 import abi, scenes

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CPP = os.path.join(HERE, "cpp")
PKG = os.path.join(ROOT, "godot_atmosphere_shader_b200")


def _test_node_binary():
    if not os.path.exists(os.path.join(PKG, "libb200atmo_node.so")):
        subprocess.check_call(["bash", os.path.join(PKG, "csrc", "build.sh")], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", CPP, "-s"])
    return os.path.join(CPP, "test_node")


def test_cpp_node_host_logic():
    out = subprocess.run([_test_node_binary(), "logic"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "0 failed" in out.stdout and int(out.stdout.split()[0]) >= 80


def test_node_library_exports_only_cxx_core_and_links_the_c_abi():
    # the node core is host code above the C-ABI: it must import b200atmo_* from libb200atmo.so, not re-implement it
    nm = subprocess.run(["nm", "-D", os.path.join(PKG, "libb200atmo_node.so")], capture_output=True, text=True, check=True).stdout
    undefined = {l.split()[-1] for l in nm.splitlines() if " U " in l}
    assert {"b200atmo_create", "b200atmo_set_params", "b200atmo_render_frame", "b200atmo_bake_optical_depth"} <= undefined
    ldd = subprocess.run(["readelf", "-d", os.path.join(PKG, "libb200atmo_node.so")], capture_output=True, text=True, check=True).stdout
    assert "libb200atmo.so" in ldd and "$ORIGIN" in ldd


@pytest.mark.gpu
def test_cpp_node_renders_like_the_python_mirror_and_the_oracle(tmp_path):
    from godot_atmosphere_shader_b200.planet_atmosphere import PlanetAtmosphere
    from oracle import pyoracle as O
    from tests import helpers as Hh

    w, h = 160, 90
    demo = scenes.demo_params()
    shape, cube, bn = Hh.demo_textures()
    cam0 = scenes.camera_a(w, h)
    P, inv_view = cam0._meta["P"], cam0._meta["inv_view"]
    depth = scenes.synth_depth(cam0, demo, w, h)
    cam_pos = inv_view[:3, 3].astype(np.float32)
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(fin, "wb") as f:
        f.write(np.array([w, h, shape.shape[0], cube.shape[1]], dtype=np.int32).tobytes())
        f.write(bytes(demo))
        for m in (np.linalg.inv(P), inv_view, np.linalg.inv(inv_view)):
            f.write(np.array(scenes.flat_colmajor(m), dtype=np.float32).tobytes())
        f.write(cam_pos.tobytes())
        f.write(np.ascontiguousarray(depth, dtype=np.float32).tobytes())
        f.write(np.ascontiguousarray(shape).tobytes())
        f.write(np.ascontiguousarray(cube).tobytes())
        f.write(np.ascontiguousarray(bn).tobytes())
    out = subprocess.run([_test_node_binary(), "render", str(fin), str(fout)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    raw = open(fout, "rb").read()
    got = np.frombuffer(raw, dtype=np.float32, count=w * h * 4).reshape(h, w, 4)
    gdisc = np.frombuffer(raw, dtype=np.uint8, count=w * h, offset=w * h * 16).reshape(h, w)
    p_cpp = abi.B200AtmoParams.from_buffer_copy(raw[w * h * 17:w * h * 17 + ctypes.sizeof(abi.B200AtmoParams)])

    # the Python mirror, same calls in the same order
    n = PlanetAtmosphere(0, blue_noise=bn)
    try:
        n.planet_radius, n.atmosphere_height = demo.planet_radius, demo.atmosphere_height
        n._ready()
        n.custom_shader = "planet_atmosphere_clouds"
        for u, f in (("u_density", "density"), ("u_scattering_strength", "scattering_strength"),
                     ("u_cloud_density_scale", "cloud_density_scale"), ("u_cloud_top", "cloud_top"),
                     ("u_cloud_shape_invert", "cloud_shape_invert"), ("u_cloud_shape_factor", "cloud_shape_factor"),
                     ("u_cloud_shape_scale", "cloud_shape_scale")):
            n.set(f"shader_params/{u}", getattr(demo, f))
        n.set("shader_params/u_cloud_shape_texture", shape)
        n.set("shader_params/u_cloud_coverage_cubemap", cube)
        n.sun_path = np.array(demo.sun_position[:])
        for _ in range(2):
            n._process(0.016, camera_position=cam_pos, now=0.0)
        cam = n.make_camera(np.linalg.inv(P), inv_view)
        want = np.empty((h, w, 4), dtype=np.float32)
        wdisc = np.empty((h, w), dtype=np.uint8)
        n.render_host(cam, depth, w, h, want, wdisc)
        assert bytes(p_cpp) == bytes(n.params), "the two node mirrors uploaded different uniform blocks"
        assert n.mode == 0  # MODE_NEAR at the demo pose: 157.9 < 1.75 * 108.1 * 1.1
    finally:
        n.free()
    assert np.array_equal(gdisc, wdisc) and np.array_equal(got.view(np.uint32), want.view(np.uint32))
    tex = O.Textures(lut=O.bake_lut(p_cpp), shape=shape, cube_faces=cube, blue_noise=bn)
    ref, rdisc = O.render_frame(p_cpp, O.variant(8, 32, abi.LIGHT_CHEAP), cam, tex, depth, w, h, threads=0)
    assert np.array_equal(gdisc, rdisc)
    Hh.assert_rgba_close(got, ref, what="C++ node")
