"""ctypes front end of libb200atmo.so (the C-ABI of include/b200atmo.h).

The library is the product; this file only marshals arguments. If the shared library is missing or
no CUDA device is present every entry point raises — there is no CPU or PyTorch fallback.
"""
import ctypes as C
import os

import numpy as np

from . import abi
from .abi import COLOR_RGBA32F, B200AtmoCamera, B200AtmoFrame, B200AtmoParams

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200ATMO_LIB", os.path.join(_HERE, "libb200atmo.so"))  # override: kernel-tuning builds only

_lib = None

# every symbol include/b200atmo.h declares
EXPORTS = [
    "b200atmo_version", "b200atmo_sizeof_params", "b200atmo_sizeof_frame", "b200atmo_sizeof_camera", "b200atmo_sizeof_peer_targets", "b200atmo_create",
    "b200atmo_destroy", "b200atmo_last_error", "b200atmo_default_params", "b200atmo_set_params", "b200atmo_get_params",
    "b200atmo_set_variant", "b200atmo_upload_blue_noise", "b200atmo_upload_shape3d", "b200atmo_upload_coverage_cube", "b200atmo_generate_noise_cubemap",
    "b200atmo_bake_optical_depth", "b200atmo_download_lut", "b200atmo_download_cube_padded", "b200atmo_render_rays",
    "b200atmo_render_rays_host", "b200atmo_render_frame", "b200atmo_render_frame_composite", "b200atmo_make_rays", "b200atmo_render_frame_host",
    "b200atmo_render_frame_host_submit", "b200atmo_frame_wait", "b200atmo_render_frame_composite_fmt", "b200atmo_composite_frame_host",
    "b200atmo_render_frame_peers", "b200atmo_render_rays_peers", "b200atmo_render_frame_peers_interleaved",
    "b200atmo_render_rays_2d", "b200atmo_render_frame_fmt", "b200atmo_render_frame_host_fmt", "b200atmo_render_frame_host_submit_fmt",
    "b200atmo_peers_wait", "b200atmo_peers_signal", "b200atmo_peers_wait_timeouts",
    "b200atmo_launch_count", "b200atmo_table_build_count",
]


class B200AtmoError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"b200atmo error {code}: {msg}")
        self.code = code


def lib():
    """Load libb200atmo.so (built by `__graft_entry__.build()` / csrc/build.sh). Raises if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with godot_atmosphere_shader_b200/csrc/build.sh "
                              "(python -c 'import __graft_entry__ as g; g.build()'). There is no fallback path.")
        L = C.CDLL(LIB_PATH)
        vp, i32, sz = C.c_void_p, C.c_int, C.c_size_t
        L.b200atmo_version.restype = i32
        for f in ("b200atmo_sizeof_params", "b200atmo_sizeof_frame", "b200atmo_sizeof_camera", "b200atmo_sizeof_peer_targets"):
            getattr(L, f).restype = sz
        L.b200atmo_create.argtypes = [i32, C.POINTER(vp)]
        L.b200atmo_destroy.argtypes = [vp]
        L.b200atmo_destroy.restype = None
        L.b200atmo_last_error.argtypes = [vp]
        L.b200atmo_last_error.restype = C.c_char_p
        L.b200atmo_default_params.argtypes = [C.POINTER(B200AtmoParams)]
        L.b200atmo_default_params.restype = None
        L.b200atmo_set_params.argtypes = [vp, C.POINTER(B200AtmoParams)]
        L.b200atmo_get_params.argtypes = [vp, C.POINTER(B200AtmoParams)]
        L.b200atmo_set_variant.argtypes = [vp, i32, i32, i32, i32]
        L.b200atmo_upload_blue_noise.argtypes = [vp, vp, i32, i32]
        L.b200atmo_upload_shape3d.argtypes = [vp, vp, i32, i32, i32]
        L.b200atmo_upload_coverage_cube.argtypes = [vp, vp, i32]
        L.b200atmo_generate_noise_cubemap.argtypes = [vp, C.POINTER(abi.B200AtmoNoise), i32, C.POINTER(C.c_float), vp, i32]
        L.b200atmo_bake_optical_depth.argtypes = [vp, vp]
        L.b200atmo_download_lut.argtypes = [vp, vp]
        L.b200atmo_download_cube_padded.argtypes = [vp, vp, sz, C.POINTER(i32)]
        L.b200atmo_render_rays.argtypes = [vp, C.POINTER(B200AtmoFrame), vp, vp, sz, vp, vp, vp]
        L.b200atmo_render_rays_host.argtypes = [vp, C.POINTER(B200AtmoFrame), vp, vp, sz, vp, vp]
        L.b200atmo_render_frame.argtypes = [vp, C.POINTER(B200AtmoCamera), vp, i32, i32, i32, i32, vp, vp, vp]
        L.b200atmo_render_frame_composite.argtypes = [vp, C.POINTER(B200AtmoCamera), vp, i32, i32, i32, i32, vp, vp]
        L.b200atmo_make_rays.argtypes = [vp, C.POINTER(B200AtmoCamera), vp, i32, i32, vp, vp, C.POINTER(B200AtmoFrame), vp]
        L.b200atmo_render_frame_host.argtypes = [vp, C.POINTER(B200AtmoCamera), vp, i32, i32, vp, vp]
        L.b200atmo_render_frame_host_submit.argtypes = [vp, C.POINTER(B200AtmoCamera), vp, i32, i32, vp, vp, i32]
        L.b200atmo_frame_wait.argtypes = [vp, i32]
        L.b200atmo_render_frame_peers.argtypes = [vp, C.POINTER(B200AtmoCamera), vp, i32, i32, i32, i32, C.POINTER(abi.B200AtmoPeerTargets), vp]
        L.b200atmo_render_frame_peers_interleaved.argtypes = [vp, C.POINTER(B200AtmoCamera), vp, i32, i32, i32, i32, C.POINTER(abi.B200AtmoPeerTargets), vp]
        L.b200atmo_render_rays_peers.argtypes = [vp, C.POINTER(B200AtmoFrame), vp, vp, C.c_size_t, C.POINTER(abi.B200AtmoPeerTargets), vp]
        L.b200atmo_render_frame_composite_fmt.argtypes = [vp, C.POINTER(B200AtmoCamera), vp, i32, i32, i32, i32, vp, i32, vp]
        L.b200atmo_composite_frame_host.argtypes = [vp, C.POINTER(B200AtmoCamera), vp, i32, i32, vp, i32]
        L.b200atmo_render_rays_2d.argtypes = [vp, C.POINTER(B200AtmoFrame), vp, vp, i32, i32, vp, vp, vp]
        L.b200atmo_render_frame_fmt.argtypes = [vp, C.POINTER(B200AtmoCamera), vp, i32, i32, i32, i32, vp, i32, vp, vp]
        L.b200atmo_render_frame_host_fmt.argtypes = [vp, C.POINTER(B200AtmoCamera), vp, i32, i32, vp, i32, vp]
        L.b200atmo_render_frame_host_submit_fmt.argtypes = [vp, C.POINTER(B200AtmoCamera), vp, i32, i32, vp, i32, vp, i32]
        L.b200atmo_peers_wait.argtypes = [vp, vp, i32, i32, C.c_uint32, vp]
        L.b200atmo_peers_signal.argtypes = [vp, C.POINTER(vp), i32, i32, C.c_uint32, vp]
        L.b200atmo_peers_wait_timeouts.argtypes = [vp]
        for f in ("b200atmo_launch_count", "b200atmo_table_build_count"):
            getattr(L, f).argtypes = [vp]
            getattr(L, f).restype = C.c_uint64
        for name in EXPORTS:
            getattr(L, name)  # AttributeError here = the library does not export what the header declares
        _lib = L
    return _lib


def _dptr(x):
    """Device/host pointer of a torch tensor, numpy array, int or None."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    raise TypeError(type(x))


class AtmosphereContext:
    """One context per device; owns the LUT and the textures (include/b200atmo.h conventions)."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        rc = lib().b200atmo_create(int(device), C.byref(self._h))
        if rc != abi.OK:
            raise B200AtmoError(rc, lib().b200atmo_last_error(None).decode())
        self.device = device
        self._cube_res = 1  # the context starts with a 1x1 white cube (unset sampler)

    def close(self):
        if getattr(self, "_h", None):
            lib().b200atmo_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != abi.OK:
            raise B200AtmoError(rc, lib().b200atmo_last_error(self._h).decode())

    # ---- uniforms ----
    def set_params(self, p: B200AtmoParams):
        self._check(lib().b200atmo_set_params(self._h, C.byref(p)))

    def get_params(self) -> B200AtmoParams:
        p = B200AtmoParams()
        self._check(lib().b200atmo_get_params(self._h, C.byref(p)))
        return p

    def set_variant(self, scatter_steps=8, cloud_steps=0, light_mode=abi.LIGHT_NONE, scatter_model=abi.SCATTER_V2):
        self._check(lib().b200atmo_set_variant(self._h, scatter_model, scatter_steps, cloud_steps, light_mode))

    # ---- textures ----
    def upload_blue_noise(self, tex: np.ndarray):
        t = np.ascontiguousarray(tex, dtype=np.uint8)
        self._check(lib().b200atmo_upload_blue_noise(self._h, t.ctypes.data, t.shape[1], t.shape[0]))

    def upload_shape3d(self, tex: np.ndarray):
        t = np.ascontiguousarray(tex, dtype=np.uint8)
        nz, ny, nx = t.shape
        self._check(lib().b200atmo_upload_shape3d(self._h, t.ctypes.data, nx, ny, nz))

    def upload_coverage_cube(self, faces: np.ndarray):
        t = np.ascontiguousarray(faces, dtype=np.uint8)
        assert t.ndim == 3 and t.shape[0] == 6 and t.shape[1] == t.shape[2]
        self._check(lib().b200atmo_upload_coverage_cube(self._h, t.ctypes.data, t.shape[1]))
        self._cube_res = int(t.shape[1])

    def generate_noise_cubemap(self, noise: "abi.B200AtmoNoise", res: int, scale=(100.0, 100.0, 100.0), download=True,
                               set_as_coverage=False):
        """NoiseCubemap._generate_images on the device; returns the 6 x res x res u8 faces if `download`."""
        out = np.empty((6, res, res), dtype=np.uint8) if download else None
        sc = (C.c_float * 3)(*[float(v) for v in scale])
        self._check(lib().b200atmo_generate_noise_cubemap(self._h, C.byref(noise), int(res), sc,
                                                          out.ctypes.data if download else None, 1 if set_as_coverage else 0))
        if set_as_coverage:
            self._cube_res = int(res)
        return out

    # ---- LUT ----
    def bake_optical_depth(self, stream=None):
        self._check(lib().b200atmo_bake_optical_depth(self._h, stream))

    def download_lut(self) -> np.ndarray:
        out = np.empty((abi.LUT_SIZE, abi.LUT_SIZE), dtype=np.float32)
        self._check(lib().b200atmo_download_lut(self._h, out.ctypes.data))
        return out

    def download_cube_padded(self) -> np.ndarray:
        r = self._cube_res
        buf = np.empty((6, r + 2, r + 2), dtype=np.uint8)
        res = C.c_int(0)
        self._check(lib().b200atmo_download_cube_padded(self._h, buf.ctypes.data, buf.size, C.byref(res)))
        assert res.value == r
        return buf

    # ---- rendering ----
    def render_rays(self, frame: B200AtmoFrame, origin_depth, dir_jitter, n, rgba, discard=None, stream=None, grid=None):
        """`grid=(width, height)`: the batch is a row-major pixel grid (n == width*height) and warps are mapped to 8x4 tiles
        (b200atmo_render_rays_2d); results are bit-identical to the linear mapping."""
        if grid is not None:
            w, h = int(grid[0]), int(grid[1])
            assert w * h == int(n), "grid does not match the ray count"
            self._check(lib().b200atmo_render_rays_2d(self._h, C.byref(frame), _dptr(origin_depth), _dptr(dir_jitter), w, h,
                                                      _dptr(rgba), _dptr(discard), stream))
            return
        self._check(lib().b200atmo_render_rays(self._h, C.byref(frame), _dptr(origin_depth), _dptr(dir_jitter), int(n),
                                               _dptr(rgba), _dptr(discard), stream))

    def render_rays_host(self, frame: B200AtmoFrame, origin_depth, dir_jitter, n, rgba, discard=None):
        self._check(lib().b200atmo_render_rays_host(self._h, C.byref(frame), _dptr(origin_depth), _dptr(dir_jitter), int(n),
                                                    _dptr(rgba), _dptr(discard)))

    def render_frame(self, cam: B200AtmoCamera, depth, w, h, rgba, discard=None, row_begin=0, row_end=None, stream=None,
                     rgba_format=COLOR_RGBA32F):
        """`rgba_format=COLOR_RGBA16F`: rgba is half4 per pixel (the fp32 result rounded to nearest-even)."""
        self._check(lib().b200atmo_render_frame_fmt(self._h, C.byref(cam), _dptr(depth), int(w), int(h), int(row_begin),
                                                    int(h if row_end is None else row_end), _dptr(rgba), int(rgba_format),
                                                    _dptr(discard), stream))

    def render_frame_composite(self, cam: B200AtmoCamera, depth, w, h, color_inout, row_begin=0, row_end=None, stream=None,
                               color_format=COLOR_RGBA32F):
        """Render and alpha-blend into the frame's colour buffer (float4 or half4 per pixel) like the ROP's blend_mix."""
        self._check(lib().b200atmo_render_frame_composite_fmt(self._h, C.byref(cam), _dptr(depth), int(w), int(h), int(row_begin),
                                                              int(h if row_end is None else row_end), _dptr(color_inout),
                                                              int(color_format), stream))

    def composite_frame_host(self, cam: B200AtmoCamera, depth, w, h, color_inout, color_format=COLOR_RGBA32F):
        """Host buffers: depth + colour up, render + blend, colour down (in place)."""
        self._check(lib().b200atmo_composite_frame_host(self._h, C.byref(cam), _dptr(depth), int(w), int(h), _dptr(color_inout),
                                                        int(color_format)))

    def make_rays(self, cam: B200AtmoCamera, depth, w, h, origin_depth, dir_jitter, stream=None) -> B200AtmoFrame:
        fr = B200AtmoFrame()
        self._check(lib().b200atmo_make_rays(self._h, C.byref(cam), _dptr(depth), int(w), int(h), _dptr(origin_depth),
                                             _dptr(dir_jitter), C.byref(fr), stream))
        return fr

    def render_frame_host(self, cam: B200AtmoCamera, depth, w, h, rgba, discard=None, rgba_format=COLOR_RGBA32F):
        self._check(lib().b200atmo_render_frame_host_fmt(self._h, C.byref(cam), _dptr(depth), int(w), int(h), _dptr(rgba),
                                                         int(rgba_format), _dptr(discard)))

    def render_frame_peers(self, cam: B200AtmoCamera, depth, w, h, targets, row_begin=0, row_end=None, stream=None):
        """Fused render + all-gather: rows [row_begin, row_end) go straight into every rank's symmetric buffer."""
        self._check(lib().b200atmo_render_frame_peers(self._h, C.byref(cam), _dptr(depth), int(w), int(h), int(row_begin),
                                                      int(h if row_end is None else row_end), C.byref(targets), stream))

    def render_frame_peers_interleaved(self, cam: B200AtmoCamera, depth, w, h, targets, first_tile, tile_pitch, stream=None):
        """Fused render + delivery of the 8-row tiles first_tile, first_tile + tile_pitch, ... (rank g of G: (g, G))."""
        self._check(lib().b200atmo_render_frame_peers_interleaved(self._h, C.byref(cam), _dptr(depth), int(w), int(h), int(first_tile),
                                                                  int(tile_pitch), C.byref(targets), stream))

    def render_rays_peers(self, frame: B200AtmoFrame, origin_depth, dir_jitter, n, targets, stream=None):
        self._check(lib().b200atmo_render_rays_peers(self._h, C.byref(frame), _dptr(origin_depth), _dptr(dir_jitter), int(n),
                                                     C.byref(targets), stream))

    def peers_wait(self, flags, first_slot, n_slots, epoch, stream=None):
        """Queue a wait until flags[first_slot : first_slot + n_slots] have all reached `epoch` (b200atmo_peers_wait)."""
        self._check(lib().b200atmo_peers_wait(self._h, _dptr(flags), int(first_slot), int(n_slots), int(epoch) & 0xFFFFFFFF, stream))

    def peers_signal(self, flag_ptrs, slot, epoch, stream=None):
        """Queue the publication of `epoch` into element `slot` of every listed flag array (b200atmo_peers_signal)."""
        arr = (C.c_void_p * len(flag_ptrs))(*[int(q) for q in flag_ptrs])
        self._check(lib().b200atmo_peers_signal(self._h, arr, len(flag_ptrs), int(slot), int(epoch) & 0xFFFFFFFF, stream))

    def peers_wait_timeouts(self) -> int:
        rc = lib().b200atmo_peers_wait_timeouts(self._h)
        if rc < 0:
            self._check(rc)
        return rc

    def render_frame_host_submit(self, cam: B200AtmoCamera, depth, w, h, rgba, discard=None, slot=0, rgba_format=COLOR_RGBA32F):
        """Pipelined host-buffer frame: returns after enqueueing; `frame_wait(slot)` completes it."""
        self._check(lib().b200atmo_render_frame_host_submit_fmt(self._h, C.byref(cam), _dptr(depth), int(w), int(h), _dptr(rgba),
                                                                int(rgba_format), _dptr(discard), int(slot)))

    def frame_wait(self, slot=0):
        self._check(lib().b200atmo_frame_wait(self._h, int(slot)))

    @property
    def launch_count(self) -> int:
        return int(lib().b200atmo_launch_count(self._h))

    @property
    def table_build_count(self) -> int:
        return int(lib().b200atmo_table_build_count(self._h))
