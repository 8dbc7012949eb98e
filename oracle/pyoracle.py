"""ORACLE — TEST INFRASTRUCTURE ONLY. ctypes front end of oracle/liboracle.so (atmo_oracle.cpp).

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs only. The product package never imports this module. Pinned bit for bit to the reference's own shader sources compiled as C++ (oracle/pyref.py,
tests/test_reference_pin.py) — see atmo_oracle.hpp.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from godot_atmosphere_shader_b200.abi import LUT_SIZE, B200AtmoCamera, B200AtmoFrame, B200AtmoParams

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")


class OracleTextures(C.Structure):
    _fields_ = [
        ("lut", C.c_void_p),
        ("lut64", C.c_void_p),
        ("shape", C.c_void_p),
        ("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32),
        ("cube_padded", C.c_void_p),
        ("cube_res", C.c_int32),
        ("blue_noise", C.c_void_p),
        ("bn_w", C.c_int32), ("bn_h", C.c_int32),
    ]


class OracleVariant(C.Structure):
    _fields_ = [("scatter_model", C.c_int32), ("scatter_steps", C.c_int32), ("cloud_steps", C.c_int32),
                ("light_mode", C.c_int32)]


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (g++ -O2 -ffp-contract=off)."""
    src_newer = (not os.path.exists(_SO)) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_SO)
        for f in ("atmo_oracle.cpp", "atmo_oracle.hpp", "Makefile"))
    if force or src_newer:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.oracle_atmosphere_density_f32.restype = C.c_float
        _lib.oracle_sample_lut_f32.restype = C.c_float
        _lib.oracle_sample_lut_f32.argtypes = [C.c_void_p, C.c_float, C.c_float]
        _lib.oracle_sample_shape_f32.restype = C.c_float
        _lib.oracle_sample_cube_f32.restype = C.c_float
        _lib.oracle_cloud_density_f32.restype = C.c_float
        _lib.oracle_cloud_light_f32.restype = C.c_float
        _lib.oracle_decode_float.restype = C.c_float
        _lib.oracle_hardware_threads.restype = C.c_int
    return _lib


def hardware_threads() -> int:
    return int(lib().oracle_hardware_threads())


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Textures:
    """Owns numpy arrays and exposes them as an OracleTextures struct."""

    def __init__(self, lut=None, lut64=None, shape=None, cube_faces=None, blue_noise=None):
        self.lut = None if lut is None else np.ascontiguousarray(lut, dtype=np.float32)
        self.lut64 = None if lut64 is None else np.ascontiguousarray(lut64, dtype=np.float64)
        self.shape = None if shape is None else np.ascontiguousarray(shape, dtype=np.uint8)
        self.blue_noise = None if blue_noise is None else np.ascontiguousarray(blue_noise, dtype=np.uint8)
        self.cube_res = 0
        self.cube_padded = None
        if cube_faces is not None:
            faces = np.ascontiguousarray(cube_faces, dtype=np.uint8)
            assert faces.ndim == 3 and faces.shape[0] == 6 and faces.shape[1] == faces.shape[2]
            self.cube_res = int(faces.shape[1])
            self.cube_padded = cube_build_padded(faces)

    def struct(self) -> OracleTextures:
        t = OracleTextures()
        t.lut = _ptr(self.lut)
        t.lut64 = _ptr(self.lut64)
        t.shape = _ptr(self.shape)
        if self.shape is not None:
            t.nz, t.ny, t.nx = self.shape.shape
        t.cube_padded = _ptr(self.cube_padded)
        t.cube_res = self.cube_res
        t.blue_noise = _ptr(self.blue_noise)
        if self.blue_noise is not None:
            t.bn_h, t.bn_w = self.blue_noise.shape
        return t


def bake_lut(params: B200AtmoParams, dtype=np.float32, via_rgba8: bool = False) -> np.ndarray:
    out = np.empty((LUT_SIZE, LUT_SIZE), dtype=dtype)
    if dtype == np.float32:
        (lib().oracle_bake_lut_via_rgba8 if via_rgba8 else lib().oracle_bake_lut_f32)(C.byref(params), _ptr(out))
    else:
        lib().oracle_bake_lut_f64(C.byref(params), _ptr(out))
    return out


def cube_build_padded(faces: np.ndarray) -> np.ndarray:
    res = int(faces.shape[1])
    out = np.zeros((6, res + 2, res + 2), dtype=np.uint8)
    lib().oracle_cube_build_padded(_ptr(np.ascontiguousarray(faces)), C.c_int(res), _ptr(out))
    return out


def variant(scatter_steps=8, cloud_steps=0, light_mode=0, scatter_model=0) -> OracleVariant:
    return OracleVariant(scatter_model, scatter_steps, cloud_steps, light_mode)


def render_rays(params, var, frame: B200AtmoFrame, tex: Textures, origin_depth, dir_jitter, dtype=np.float32, threads=1):
    od = np.ascontiguousarray(origin_depth, dtype=np.float32)
    dj = np.ascontiguousarray(dir_jitter, dtype=np.float32)
    n = od.shape[0]
    rgba = np.empty((n, 4), dtype=dtype)
    disc = np.empty((n,), dtype=np.uint8)
    ts = tex.struct()
    fn = lib().oracle_render_rays_f32 if dtype == np.float32 else lib().oracle_render_rays_f64
    fn(C.byref(params), C.byref(var), C.byref(frame), C.byref(ts), _ptr(od), _ptr(dj), C.c_size_t(n), _ptr(rgba), _ptr(disc),
       C.c_int(threads))
    return rgba, disc


def render_frame(params, var, cam: B200AtmoCamera, tex: Textures, depth, w, h, row_begin=0, row_end=None, dtype=np.float32,
                 threads=1):
    row_end = h if row_end is None else row_end
    dep = np.ascontiguousarray(depth, dtype=np.float32)
    rgba = np.zeros((h, w, 4), dtype=dtype)
    disc = np.zeros((h, w), dtype=np.uint8)
    ts = tex.struct()
    fn = lib().oracle_render_frame_f32 if dtype == np.float32 else lib().oracle_render_frame_f64
    fn(C.byref(params), C.byref(var), C.byref(cam), C.byref(ts), _ptr(dep), C.c_int(w), C.c_int(h), C.c_int(row_begin),
       C.c_int(row_end), _ptr(rgba), _ptr(disc), C.c_int(threads))
    return rgba, disc


def make_rays(params, cam: B200AtmoCamera, tex: Textures, depth, w, h):
    dep = np.ascontiguousarray(depth, dtype=np.float32)
    od = np.empty((h * w, 4), dtype=np.float32)
    dj = np.empty((h * w, 4), dtype=np.float32)
    fr = B200AtmoFrame()
    ts = tex.struct()
    lib().oracle_make_rays_f32(C.byref(params), C.byref(cam), C.byref(ts), _ptr(dep), C.c_int(w), C.c_int(h), _ptr(od),
                               _ptr(dj), C.byref(fr))
    return od, dj, fr


# ---- per-function hooks ---------------------------------------------------------------------------
def ray_sphere(center, radius, origin, direction, dtype=np.float32):
    if dtype == np.float32:
        out = (C.c_float * 2)()
        lib().oracle_ray_sphere_f32(_f3(center), C.c_float(radius), _f3(origin), _f3(direction), out)
    else:
        out = (C.c_double * 2)()
        d3 = lambda v: (C.c_double * 3)(*[float(x) for x in v])
        lib().oracle_ray_sphere_f64(d3(center), C.c_double(radius), d3(origin), d3(direction), out)
    return float(out[0]), float(out[1])


def ray_box(ro, rd, box):
    out = (C.c_float * 2)()
    lib().oracle_ray_box_f32(_f3(ro), _f3(rd), _f3(box), out)
    return float(out[0]), float(out[1])


def atmosphere_density(params, height: float) -> float:
    return float(lib().oracle_atmosphere_density_f32(C.byref(params), C.c_float(height)))


def blend_colors(self_rgba, over_rgba):
    out = (C.c_float * 4)()
    lib().oracle_blend_colors_f32((C.c_float * 4)(*self_rgba), (C.c_float * 4)(*over_rgba), out)
    return tuple(float(x) for x in out)


def sample_lut(lut: np.ndarray, u: float, v: float) -> float:
    l = np.ascontiguousarray(lut, dtype=np.float32)
    return float(lib().oracle_sample_lut_f32(_ptr(l), C.c_float(u), C.c_float(v)))


def sample_shape(params, tex: Textures, pos) -> float:
    ts = tex.struct()
    return float(lib().oracle_sample_shape_f32(C.byref(params), C.byref(ts), _f3(pos)))


def sample_cube(params, tex: Textures, direction) -> float:
    ts = tex.struct()
    return float(lib().oracle_sample_cube_f32(C.byref(params), C.byref(ts), _f3(direction)))


def cloud_density(params, tex: Textures, pos) -> float:
    ts = tex.struct()
    return float(lib().oracle_cloud_density_f32(C.byref(params), C.byref(ts), _f3(pos)))


def cloud_light(params, tex: Textures, light_mode, pos, ray_dir, sun_dir, jitter=0.0, alpha=0.0) -> float:
    ts = tex.struct()
    return float(lib().oracle_cloud_light_f32(C.byref(params), C.byref(ts), C.c_int(light_mode), _f3(pos), _f3(ray_dir),
                                              _f3(sun_dir), C.c_float(jitter), C.c_float(alpha)))


def raymarch_cloud(params, tex: Textures, steps, light_mode, origin, direction, t_begin, t_end, jitter, sun_dir):
    ts = tex.struct()
    out = (C.c_float * 2)()
    lib().oracle_raymarch_cloud_f32(C.byref(params), C.byref(ts), C.c_int(steps), C.c_int(light_mode), _f3(origin),
                                    _f3(direction), C.c_float(t_begin), C.c_float(t_end), C.c_float(jitter), _f3(sun_dir), out)
    return float(out[0]), float(out[1])


def compute_atmosphere_v2(params, lut, steps, origin, direction, planet_center, t_begin, t_end, sun_dir, jitter):
    l = np.ascontiguousarray(lut, dtype=np.float32)
    out = (C.c_float * 4)()
    lib().oracle_compute_atmosphere_v2_f32(C.byref(params), _ptr(l), C.c_int(steps), _f3(origin), _f3(direction),
                                           _f3(planet_center), C.c_float(t_begin), C.c_float(t_end), _f3(sun_dir),
                                           C.c_float(jitter), out)
    return tuple(float(x) for x in out)


def noise_cubemap(noise, res: int, scale=(100.0, 100.0, 100.0)) -> np.ndarray:
    out = np.empty((6, res, res), dtype=np.uint8)
    lib().oracle_noise_cubemap(C.byref(noise), C.c_int(res), (C.c_float * 3)(*[float(v) for v in scale]), _ptr(out))
    return out


def noise3(x, y, z, seed=0) -> float:
    lib().oracle_noise3.restype = C.c_float
    return float(lib().oracle_noise3(C.c_float(x), C.c_float(y), C.c_float(z), C.c_uint32(seed)))


def cubemap_atlas(faces: np.ndarray) -> np.ndarray:
    res = int(faces.shape[1])
    out = np.empty((2 * res, 3 * res), dtype=np.uint8)
    lib().oracle_cubemap_atlas(_ptr(np.ascontiguousarray(faces)), C.c_int(res), _ptr(out))
    return out


def encode_float(h: float):
    out = (C.c_uint8 * 4)()
    lib().oracle_encode_float(C.c_float(h), out)
    return bytes(out)


def decode_float(b: bytes) -> float:
    return float(lib().oracle_decode_float((C.c_uint8 * 4)(*b)))
