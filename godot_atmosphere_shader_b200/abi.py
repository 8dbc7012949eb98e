"""ctypes mirrors of the PODs declared in include/b200atmo.h (the C-ABI of the hot path).

Field order and types must match the header exactly; tests/test_abi.py checks sizes against the
compiled library (b200atmo_sizeof_*).
"""
import ctypes as C

LUT_SIZE = 256

OK = 0
E_INVALID = -1
E_CUDA = -2
E_NOMEM = -3
E_STATE = -4

SCATTER_V2 = 0
SCATTER_V1 = 1

LIGHT_NONE = 0
LIGHT_CHEAP = 1
LIGHT_RAYMARCHED = 2

COLOR_RGBA32F = 0
COLOR_RGBA16F = 1
PIPELINE_SLOTS = 4   # B200ATMO_PIPELINE_SLOTS


class B200AtmoParams(C.Structure):
    """Shader uniform set, SURVEY.md §8(b2). Names are the reference's uniform names minus `u_`."""

    _fields_ = [
        ("planet_radius", C.c_float),
        ("atmosphere_height", C.c_float),
        ("sun_position", C.c_float * 3),
        ("density", C.c_float),
        ("scattering_strength", C.c_float),
        ("scattering_wavelengths", C.c_float * 3),
        ("atmosphere_modulate", C.c_float * 3),
        ("atmosphere_ambient_color", C.c_float * 3),
        ("clip_mode", C.c_float),
        ("sphere_depth_factor", C.c_float),
        ("cloud_density_scale", C.c_float),
        ("cloud_bottom", C.c_float),
        ("cloud_top", C.c_float),
        ("cloud_blend", C.c_float),
        ("cloud_shape_invert", C.c_float),
        ("cloud_coverage_bias", C.c_float),
        ("cloud_shape_factor", C.c_float),
        ("cloud_shape_scale", C.c_float),
        ("cloud_coverage_rotation", C.c_float * 4),
        ("world_to_model", C.c_float * 16),
        ("day_color0", C.c_float * 4),
        ("day_color1", C.c_float * 4),
        ("night_color0", C.c_float * 4),
        ("night_color1", C.c_float * 4),
        ("day_night_transition_scale", C.c_float),
    ]

    def copy(self):
        out = B200AtmoParams()
        C.memmove(C.byref(out), C.byref(self), C.sizeof(B200AtmoParams))
        return out


class B200AtmoFrame(C.Structure):
    _fields_ = [
        ("planet_center_view", C.c_float * 3),
        ("sun_center_view", C.c_float * 3),
        ("inv_view", C.c_float * 16),
    ]


class B200AtmoCamera(C.Structure):
    _fields_ = [
        ("inv_projection", C.c_float * 16),
        ("inv_view", C.c_float * 16),
        ("view", C.c_float * 16),
        ("model", C.c_float * 16),
        ("double_precision", C.c_int32),
        ("clip_box_size", C.c_float),
    ]


MAX_PEERS = 8


class B200AtmoPeerSync(C.Structure):
    """Hand-shake carried out by the render kernel itself (include/b200atmo.h)."""

    _fields_ = [("d_done_flags", C.c_void_p * MAX_PEERS), ("n_done_flags", C.c_int32), ("done_slot", C.c_int32), ("epoch", C.c_uint32),
                ("credit_epoch", C.c_uint32), ("d_credit_flags", C.c_void_p), ("credit_first_slot", C.c_int32), ("n_credit", C.c_int32),
                ("d_consumed_flags", C.c_void_p * MAX_PEERS), ("n_consumed_flags", C.c_int32), ("consumed_slot", C.c_int32),
                ("consumed_epoch", C.c_uint32), ("n_wait", C.c_int32), ("d_wait_flags", C.c_void_p), ("wait_first_slot", C.c_int32),
                ("reserved", C.c_int32)]


class B200AtmoPeerTargets(C.Structure):
    """Where the fused render + all-gather kernels store: the same symmetric buffer on every rank (include/b200atmo.h)."""

    _fields_ = [("d_rgba_peers", C.c_void_p * MAX_PEERS), ("n_peers", C.c_int32), ("d_rgba_multicast", C.c_void_p),
                ("elem_offset", C.c_uint64), ("first_peer", C.c_int32), ("use_tma", C.c_int32), ("rgba_format", C.c_int32),
                ("reserved", C.c_int32), ("sync", B200AtmoPeerSync)]


class B200AtmoNoise(C.Structure):
    """Subset of FastNoiseLite's properties used by the generator (include/b200atmo.h)."""

    _fields_ = [("seed", C.c_int32), ("frequency", C.c_float), ("octaves", C.c_int32), ("lacunarity", C.c_float),
                ("gain", C.c_float)]


IDENTITY16 = (1.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0, 1.0)


def default_params() -> B200AtmoParams:
    """Shader-source defaults (same values b200atmo_default_params() writes)."""
    p = B200AtmoParams()
    p.planet_radius = 1.0
    p.atmosphere_height = 0.1
    p.sun_position[:] = (0.0, 0.0, 0.0)
    p.density = 0.2
    p.scattering_strength = 20.0
    p.scattering_wavelengths[:] = (700.0, 530.0, 440.0)
    p.atmosphere_modulate[:] = (1.0, 1.0, 1.0)
    p.atmosphere_ambient_color[:] = (0.0, 0.0, 0.002)
    p.clip_mode = 0.0
    p.sphere_depth_factor = 0.0
    p.cloud_density_scale = 50.0
    p.cloud_bottom = 0.2
    p.cloud_top = 0.5
    p.cloud_blend = 0.5
    p.cloud_shape_invert = 0.0
    p.cloud_coverage_bias = 0.0
    p.cloud_shape_factor = 0.8
    p.cloud_shape_scale = 1.0
    p.cloud_coverage_rotation[:] = (1.0, 0.0, 0.0, 1.0)
    p.world_to_model[:] = IDENTITY16
    p.day_color0[:] = (0.5, 0.8, 1.0, 1.0)
    p.day_color1[:] = (0.5, 0.8, 1.0, 1.0)
    p.night_color0[:] = (0.2, 0.4, 0.8, 1.0)
    p.night_color1[:] = (0.2, 0.4, 0.8, 1.0)
    p.day_night_transition_scale = 2.0
    return p
