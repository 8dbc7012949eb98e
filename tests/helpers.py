"""Shared test helpers: tolerances, scene set-up, padded texture layouts, the host-sim build of the device code."""
import ctypes as C
import os
import subprocess

import numpy as np

from godot_atmosphere_shader_b200 import abi, scenes
from oracle import pyoracle as O

HERE = os.path.dirname(os.path.abspath(__file__))

# Parity tolerance against the fp32 oracle (BASELINE.json north_star: 1e-4 relative per channel).
# ATOL is the absolute floor for values near zero: one fp32 rounding of a transmittance next to 1.0 is
# 6e-8 and the oracle accumulates up to 128 of them into alpha / cloud light (SURVEY.md §7 hard part 2).
RTOL = 1e-4
ATOL = 2e-6


def assert_rgba_close(got, want, rtol=RTOL, atol=ATOL, what=""):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (got.shape, want.shape)
    err = np.abs(got - want)
    bound = rtol * np.abs(want) + atol
    bad = ~(err <= bound)
    if bad.any():
        idx = np.unravel_index(np.argmax(np.where(bad, err / bound, 0)), err.shape)
        raise AssertionError(f"{what}: {int(bad.sum())}/{bad.size} values out of tolerance; worst at {idx}: got {got[idx]!r} "
                             f"want {want[idx]!r} (|err|={err[idx]:.3e}, bound={bound[idx]:.3e})")


def rel_err_stats(got, want, floor=1e-3):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    rel = np.abs(got - want) / np.maximum(np.abs(want), floor)
    return float(rel.max()), float(np.quantile(rel, 0.999))


# ------------------------------------------------------------------------------------------------
# padded fp32 texture layouts as the kernels see them (numpy restatement, used by host-sim tests)
# ------------------------------------------------------------------------------------------------
def lut_padded(lut: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(np.pad(lut.astype(np.float32), 1, mode="edge"))


def cube_padded_f32(faces: np.ndarray) -> np.ndarray:
    return (O.cube_build_padded(np.ascontiguousarray(faces)).astype(np.float32) / np.float32(255.0)).astype(np.float32)


def shape_padded_f32(shape: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray((np.pad(shape, 1, mode="wrap").astype(np.float32) / np.float32(255.0)).astype(np.float32))


class HostsimTextures(C.Structure):
    _fields_ = [("lut_pad", C.c_void_p), ("cube_pad", C.c_void_p), ("cube_res", C.c_int32), ("shape_pad", C.c_void_p),
                ("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32), ("blue_noise", C.c_void_p), ("bn_w", C.c_int32),
                ("bn_h", C.c_int32)]


_hostsim = None


def hostsim():
    global _hostsim
    if _hostsim is None:
        d = os.path.join(HERE, "hostsim")
        subprocess.check_call(["make", "-C", d, "-s"])
        _hostsim = C.CDLL(os.path.join(d, "libhostsim.so"))
        _hostsim.hostsim_sqrt_refined.restype = C.c_float
        _hostsim.hostsim_sqrt_refined.argtypes = [C.c_float]
        _hostsim.hostsim_div_refined.restype = C.c_float
        _hostsim.hostsim_div_refined.argtypes = [C.c_float, C.c_float]
    return _hostsim


class HostsimScene:
    """Keeps the padded arrays alive and runs the host build of shade_ray."""

    def __init__(self, lut, shape=None, cube_faces=None, blue_noise=None):
        self.lut_pad = lut_padded(lut)
        self.cube_pad = cube_padded_f32(cube_faces if cube_faces is not None else np.full((6, 1, 1), 255, np.uint8))
        self.cube_res = int(cube_faces.shape[1]) if cube_faces is not None else 1
        shp = shape if shape is not None else np.full((1, 1, 1), 255, np.uint8)
        self.shape_pad = shape_padded_f32(shp)
        self.nz, self.ny, self.nx = shp.shape
        self.blue = None if blue_noise is None else np.ascontiguousarray(blue_noise, dtype=np.uint8)

    def struct(self):
        t = HostsimTextures()
        t.lut_pad = self.lut_pad.ctypes.data
        t.cube_pad = self.cube_pad.ctypes.data
        t.cube_res = self.cube_res
        t.shape_pad = self.shape_pad.ctypes.data
        t.nx, t.ny, t.nz = self.nx, self.ny, self.nz
        if self.blue is not None:
            t.blue_noise = self.blue.ctypes.data
            t.bn_h, t.bn_w = self.blue.shape
        return t

    def render_rays(self, params, variant, frame, od, dj):
        od = np.ascontiguousarray(od, dtype=np.float32)
        dj = np.ascontiguousarray(dj, dtype=np.float32)
        n = od.shape[0]
        rgba = np.empty((n, 4), np.float32)
        disc = np.empty((n,), np.uint8)
        var = (C.c_int32 * 4)(variant.scatter_model, variant.scatter_steps, variant.cloud_steps, variant.light_mode)
        ts = self.struct()
        hostsim().hostsim_render_rays(C.byref(params), var, C.byref(frame), C.byref(ts), od.ctypes.data_as(C.c_void_p),
                                      dj.ctypes.data_as(C.c_void_p), C.c_size_t(n), rgba.ctypes.data_as(C.c_void_p),
                                      disc.ctypes.data_as(C.c_void_p))
        return rgba, disc

    def make_rays(self, params, cam, depth, w, h):
        dep = np.ascontiguousarray(depth, dtype=np.float32)
        od = np.empty((h * w, 4), np.float32)
        dj = np.empty((h * w, 4), np.float32)
        fr = abi.B200AtmoFrame()
        ts = self.struct()
        hostsim().hostsim_make_rays(C.byref(params), C.byref(cam), C.byref(ts), dep.ctypes.data_as(C.c_void_p), C.c_int(w),
                                    C.c_int(h), od.ctypes.data_as(C.c_void_p), dj.ctypes.data_as(C.c_void_p), C.byref(fr))
        return od, dj, fr


# ------------------------------------------------------------------------------------------------
# scene bundles
# ------------------------------------------------------------------------------------------------
_cache = {}


def demo_textures(shape_n=32, cube_res=64):
    key = ("tex", shape_n, cube_res)
    if key not in _cache:
        _cache[key] = (scenes.shape_texture(shape_n, seed=1), scenes.coverage_cubemap(cube_res, seed=1), scenes.blue_noise_tile())
    return _cache[key]


def random_rays(n, params, seed=0, inside_fraction=0.3):
    """Seeded mix of rays: outside looking at/near the planet, inside the atmosphere, grazing, missing."""
    rng = np.random.default_rng(seed)
    R, H = float(params.planet_radius), float(params.atmosphere_height)
    center = np.array([3.0, -2.0, -(R + H) * 1.5])  # planet centre in view space
    n_in = int(n * inside_fraction)
    o = np.zeros((n, 3))
    # inside: origin in the shell
    u = rng.normal(size=(n_in, 3))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    o[:n_in] = center + u * (R + H * rng.uniform(0.02, 0.98, size=(n_in, 1)))
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    # outside rays: aim at a point within 1.3 atmosphere radii of the centre
    tgt = center + rng.normal(size=(n - n_in, 3)) * (R + H) * 0.6
    dd = tgt - o[n_in:]
    d[n_in:] = dd / np.linalg.norm(dd, axis=1, keepdims=True)
    depth = np.where(rng.random(n) < 0.5, 1e4, rng.uniform(0.5, 3.0 * (R + H), size=n))
    jitter = rng.integers(0, 256, size=n) / 255.0
    od = np.concatenate([o, depth[:, None]], axis=1).astype(np.float32)
    d32 = d.astype(np.float32)
    d32 /= np.linalg.norm(d32.astype(np.float64), axis=1, keepdims=True).astype(np.float32)
    dj = np.concatenate([d32, jitter[:, None].astype(np.float32)], axis=1).astype(np.float32)
    fr = abi.B200AtmoFrame()
    fr.planet_center_view[:] = tuple(np.float32(center).tolist())
    sun = center + np.array([0.3, 0.5, 0.8]) * 5000.0
    fr.sun_center_view[:] = tuple(np.float32(sun).tolist())
    fr.inv_view[:] = abi.IDENTITY16
    return od, dj, fr


def random_scene(seed):
    """A seeded random uniform block + frame: planet scale over 3 decades, rotated/translated node transform, random
    cloud shell, blend, bias, invert, coverage rotation, sphere-depth blend. Returns (params, frame, rays...)."""
    from godot_atmosphere_shader_b200 import scenes as sc
    rng = np.random.default_rng(seed)
    p = sc.demo_params()
    R = float(10.0 ** rng.uniform(0.0, 3.0))
    H = R * float(rng.uniform(0.03, 0.25))
    p.planet_radius, p.atmosphere_height = R, H
    p.density = float(rng.uniform(0.5, 6.0)) / H              # optical thickness of order 1
    p.scattering_strength = float(rng.uniform(0.3, 3.0))
    p.scattering_wavelengths[:] = tuple(float(x) for x in (rng.uniform(620, 750), rng.uniform(500, 570), rng.uniform(420, 480)))
    p.atmosphere_modulate[:] = tuple(float(x) for x in rng.uniform(0.5, 1.0, 3))
    p.atmosphere_ambient_color[:] = tuple(float(x) for x in rng.uniform(0.0, 0.05, 3))
    p.sphere_depth_factor = float(rng.choice([0.0, 0.0, 0.3, 1.0]))
    b = float(rng.uniform(0.05, 0.5))
    p.cloud_bottom, p.cloud_top = b, b + float(rng.uniform(0.1, 0.45))
    p.cloud_density_scale = float(rng.uniform(0.5, 8.0)) * 8.0 / H
    p.cloud_blend = float(rng.uniform(0.0, 1.0))
    p.cloud_shape_invert = float(rng.integers(0, 2))
    p.cloud_coverage_bias = float(rng.uniform(-0.2, 0.3))
    p.cloud_shape_factor = float(rng.uniform(0.0, 1.0))
    p.cloud_shape_scale = float(rng.uniform(2.0, 12.0)) / R
    a = float(rng.uniform(0, 6.28))
    p.cloud_coverage_rotation[:] = (np.cos(a), np.sin(a), -np.sin(a), np.cos(a))
    # node transform: random rotation + translation; camera looks at the planet from 1.05 .. 4 radii
    q = rng.normal(size=4); q /= np.linalg.norm(q)
    w, x, y, z = q
    Rm = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                   [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                   [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    node = np.eye(4); node[:3, :3] = Rm; node[:3, 3] = rng.normal(size=3) * R * 3
    p.world_to_model[:] = sc.flat_colmajor(np.linalg.inv(node))
    dist = (R + H) * float(rng.uniform(1.02, 4.0)) if rng.random() < 0.7 else R + H * float(rng.uniform(0.05, 0.95))
    dirv = rng.normal(size=3); dirv /= np.linalg.norm(dirv)
    eye = node[:3, 3] + dirv * dist
    side = np.cross(dirv, rng.normal(size=3)); side /= np.linalg.norm(side)
    fwd = -dirv + side * float(rng.uniform(0.0, 0.6))
    up = np.cross(fwd, side)
    cam = sc.make_camera(eye, fwd, up=up, fovy_deg=float(rng.uniform(30, 90)), aspect=1.5, near=0.05 * H, far=50 * R, model=node)
    p.sun_position[:] = tuple(float(v) for v in (node[:3, 3] + rng.normal(size=3) * 40 * R))
    return p, cam
