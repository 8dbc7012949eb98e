"""TEST INFRASTRUCTURE ONLY — ctypes front end of oracle/_ref/libatmo_ref.so: the reference's own shader sources compiled
as C++ (oracle/ref/build_ref.py). Used to PIN the hand-written oracle (tests/test_reference_pin.py) and, where the
library has been built, as the CPU arm of bench.py (`cpu_baseline.kind = "reference"`). The product never imports this.

The library is built from /root/reference, which exists in the build container only; the built .so travels to the GPU
box. `available()` says whether it can be used here."""
import ctypes as C
import os

import numpy as np

from godot_atmosphere_shader_b200.abi import LUT_SIZE, B200AtmoCamera, B200AtmoParams

from .pyoracle import OracleVariant, Textures, _ptr

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libatmo_ref.so")
_SO64 = os.path.join(_HERE, "_ref", "libatmo_ref64.so")   # the same sources with `float` = double (rounding-error bound)
REFERENCE = os.environ.get("B200ATMO_REFERENCE", "/root/reference")

_lib = None
_lib64 = None


def reference_present() -> bool:
    return os.path.isdir(os.path.join(REFERENCE, "addons", "zylann.atmosphere", "shaders"))


def build(force: bool = False) -> str:
    """(Re)build from the reference tree when it is present; otherwise the prebuilt library must exist."""
    if reference_present():
        srcs = [os.path.join(_HERE, "ref", f) for f in os.listdir(os.path.join(_HERE, "ref"))] + [os.path.join(_HERE, "atmo_oracle.hpp")]
        stale = (not os.path.exists(_SO)) or (not os.path.exists(_SO64)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
        if force or stale:
            from .ref import build_ref
            build_ref.build(REFERENCE, verbose=False)
    if not os.path.exists(_SO):
        raise FileNotFoundError(f"{_SO} is missing and {REFERENCE} is not present to build it from")
    return _SO


def available() -> bool:
    return os.path.exists(_SO) or reference_present()


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.ref_entry_name.restype = C.c_char_p
        _lib.ref_entry_count.restype = C.c_int
    return _lib


def lib64():
    global _lib64
    if _lib64 is None:
        build()
        _lib64 = C.CDLL(_SO64)
        assert _lib64.ref_bake_real_size() == 8
    return _lib64


def entry_shaders() -> dict:
    """name -> (ATMOSPHERE_LITE, ATMOSPHERE_RAYMARCH_STEPS, CLOUDS_MAX_RAYMARCH_STEPS or 0, CLOUDS_RAYMARCHED_LIGHTING):
    the #defines of every shipped entry shader, as compiled."""
    out = {}
    for i in range(lib().ref_entry_count()):
        d = (C.c_int * 4)()
        lib().ref_entry_defines(C.c_int(i), d)
        out[lib().ref_entry_name(C.c_int(i)).decode()] = tuple(int(x) for x in d)
    return out


def bake_lut(params: B200AtmoParams, dtype=np.float32) -> np.ndarray:
    """optical_depth.gdshader run over the 256 x 256 canvas. float32: through the RGBA8 viewport and the FORMAT_RF
    reinterpretation, like OpticalDepthBaker; float64 (the twin): the value handed to encode_float_to_viewport."""
    out = np.empty((LUT_SIZE, LUT_SIZE), dtype=dtype)
    (lib() if dtype == np.float32 else lib64()).ref_bake_lut(C.byref(params), _ptr(out))
    return out


def render_frame(params, var: OracleVariant, cam: B200AtmoCamera, tex: Textures, depth, w, h, row_begin=0, row_end=None,
                 threads=1, shader: str = None, row_stride=1, dtype=np.float32):
    """One draw with the compiled entry shader `shader` (default: the shipped shader whose feature #defines match `var`).
    dtype=np.float64 runs the fp64 twin of the same sources."""
    assert not (cam.clip_box_size > 0.0), "the MODE_FAR proxy mesh is rasteriser behaviour, not shader code"
    row_end = h if row_end is None else row_end
    dep = np.ascontiguousarray(depth, dtype=np.float32)
    rgba = np.zeros((h, w, 4), dtype=dtype)
    disc = np.zeros((h, w), dtype=np.uint8)
    ts = tex.struct()
    L = lib() if dtype == np.float32 else lib64()
    rc = L.ref_render_frame(shader.encode() if shader else None, C.byref(params), C.byref(var), C.byref(cam), C.byref(ts),
                            _ptr(dep), C.c_int(w), C.c_int(h), C.c_int(row_begin), C.c_int(row_end), C.c_int(row_stride), _ptr(rgba),
                            _ptr(disc), C.c_int(threads))
    if rc != 0:
        raise ValueError(f"no compiled reference shader for variant {tuple(getattr(var, f) for f, _ in var._fields_)} / {shader}")
    return rgba, disc
