"""Host-side mirror of the reference's PlanetAtmosphere node (addons/zylann.atmosphere/planet_atmosphere.gd)
and OpticalDepthBaker (addons/zylann.atmosphere/optical_depth_baker.gd) above the C-ABI.

Same member names, argument meaning, defaults, deprecation warnings and bake triggers as the GDScript, so code
written against the node reads the same here; what the engine did implicitly (scene tree, camera, material) is
passed in explicitly. The GDExtension shim in INTEGRATION.md binds the same C-ABI calls from C++.

This module contains no arithmetic of the hot path: it only fills B200AtmoParams / B200AtmoCamera and calls
libb200atmo.so. Matrices are 4x4 numpy arrays in math convention (M[row, col]); they are flattened
column-major for the ABI.
"""
import math
import time
import warnings

import numpy as np

from . import abi
from .context import AtmosphereContext
from .scenes import flat_colmajor

# planet_atmosphere.gd:9-11
MODE_NEAR = 0
MODE_FAR = 1
SWITCH_MARGIN_RATIO = 1.1

# custom_shader (planet_atmosphere.gd:44-49): the entry shaders are nothing but #defines
# (shaders/planet_atmosphere_*.gdshader:4-7); name -> (scatter model, ATMOSPHERE_RAYMARCH_STEPS,
# CLOUDS_MAX_RAYMARCH_STEPS, light mode)
SHADER_VARIANTS = {
    "planet_atmosphere_no_clouds": (abi.SCATTER_V2, 8, 0, abi.LIGHT_NONE),
    "planet_atmosphere_clouds": (abi.SCATTER_V2, 8, 32, abi.LIGHT_CHEAP),
    "planet_atmosphere_clouds_high": (abi.SCATTER_V2, 8, 64, abi.LIGHT_CHEAP),
    "planet_atmosphere_clouds_high_rm": (abi.SCATTER_V2, 8, 64, abi.LIGHT_RAYMARCHED),
    "planet_atmosphere_v1_no_clouds": (abi.SCATTER_V1, 16, 0, abi.LIGHT_NONE),
    "planet_atmosphere_v1_clouds": (abi.SCATTER_V1, 16, 32, abi.LIGHT_CHEAP),
    "planet_atmosphere_v1_clouds_high": (abi.SCATTER_V1, 16, 64, abi.LIGHT_CHEAP),
}
DEFAULT_SHADER = "planet_atmosphere_no_clouds"  # planet_atmosphere.gd:13-14

# uniform name -> (B200AtmoParams field, kind); kinds: f = float, v3 = vec3, c3/c4 = source_color (sRGB in, linear stored)
_V2_UNIFORMS = {
    "u_density": ("density", "f"),
    "u_scattering_strength": ("scattering_strength", "f"),
    "u_scattering_wavelengths": ("scattering_wavelengths", "v3"),
    "u_atmosphere_modulate": ("atmosphere_modulate", "c3"),
    "u_atmosphere_ambient_color": ("atmosphere_ambient_color", "c3"),
    "u_sphere_depth_factor": ("sphere_depth_factor", "f"),
}
_V1_UNIFORMS = {
    "u_density": ("density", "f"),
    "u_day_color0": ("day_color0", "c4"),
    "u_day_color1": ("day_color1", "c4"),
    "u_night_color0": ("night_color0", "c4"),
    "u_night_color1": ("night_color1", "c4"),
    "u_day_night_transition_scale": ("day_night_transition_scale", "f"),
    "u_sphere_depth_factor": ("sphere_depth_factor", "f"),
}
_CLOUD_UNIFORMS = {
    "u_cloud_density_scale": ("cloud_density_scale", "f"),
    "u_cloud_bottom": ("cloud_bottom", "f"),
    "u_cloud_top": ("cloud_top", "f"),
    "u_cloud_blend": ("cloud_blend", "f"),
    "u_cloud_shape_invert": ("cloud_shape_invert", "f"),
    "u_cloud_coverage_bias": ("cloud_coverage_bias", "f"),
    "u_cloud_shape_factor": ("cloud_shape_factor", "f"),
    "u_cloud_shape_scale": ("cloud_shape_scale", "f"),
    "u_cloud_shape_texture": (None, "tex3d"),
    "u_cloud_coverage_cubemap": (None, "cube"),
}
# planet_atmosphere.gd:68-77 — assigned internally, hidden from the shader_params list
_API_SHADER_PARAMS = {"u_planet_radius", "u_atmosphere_height", "u_clip_mode", "u_sun_position", "u_world_to_model_matrix",
                      "u_blue_noise_texture", "u_cloud_coverage_rotation", "u_optical_depth_texture"}
# planet_atmosphere.gd:79-81
_SHADER_PARAMS_AFFECTING_OPTICAL_DEPTH = {"u_density"}


def srgb_to_linear(c):
    """Godot converts `source_color` uniforms before upload (engine behaviour, Color.srgb_to_linear)."""
    c = np.asarray(c, dtype=np.float64)
    return np.where(c < 0.04045, c / 12.92, ((c + 0.055) / 1.055) ** 2.4)


class OpticalDepthBaker:
    """optical_depth_baker.gd:3-85 — the 3-state machine, minus the SubViewport: `_process` #1 launches the bake
    kernel, `_process` #2 emits `baked`."""

    STATE_IDLE = 0
    STATE_REQUEST_BAKE = 1
    STATE_PENDING_RENDER = 2

    def __init__(self, ctx: AtmosphereContext):
        self._ctx = ctx
        self._state = self.STATE_IDLE
        self._baked_callbacks = []
        self.processing = False  # set_process()

    def connect_baked(self, fn):
        self._baked_callbacks.append(fn)

    def request_bake(self, atmosphere_params):
        self._state = self.STATE_REQUEST_BAKE
        self._params = atmosphere_params
        self.processing = True

    def _process(self, _delta=0.0):
        if self._state == self.STATE_REQUEST_BAKE:
            # _setup_bake: uniforms are copied by name from the atmosphere material (baker.gd:55-59)
            self._ctx.set_params(self._params)
            self._ctx.bake_optical_depth()
            self._state = self.STATE_PENDING_RENDER
        elif self._state == self.STATE_PENDING_RENDER:
            for fn in self._baked_callbacks:
                fn(self._ctx)  # the "texture" stays on the device, owned by the context
            self._state = self.STATE_IDLE
            self.processing = False


class PlanetAtmosphere:
    """Drop-in surface of the `PlanetAtmosphere` node.

    Engine objects become plain arguments: `global_transform` is a 4x4 matrix, `sun_path` any object with a
    `global_transform` 4x4 attribute or a 3-vector, the camera is passed to `_process`/`render`.
    """

    MODE_NEAR = MODE_NEAR
    MODE_FAR = MODE_FAR
    SWITCH_MARGIN_RATIO = SWITCH_MARGIN_RATIO

    def __init__(self, device: int = 0, blue_noise=None, ctx=None):
        # `ctx`: an already-created AtmosphereContext (or, in host-logic tests, a recording stand-in)
        self._ctx = ctx if ctx is not None else AtmosphereContext(device)
        self._p = abi.default_params()
        self._planet_radius = 1.0          # planet_atmosphere.gd:20
        self._atmosphere_height = 0.1      # :28
        self._sun_path = None
        self._custom_shader = None
        self.clouds_rotation_speed = 1.0   # :52, degrees per second
        self.force_fullscreen = False      # :54
        self.global_transform = np.eye(4)
        self._mode = MODE_FAR              # :58
        self._prev_atmo_clip_distance = 0.0
        self._far_mesh_size = 1.0          # :99
        self._uses_baked_optical_depth = False
        self._optical_depth_baker = None
        self._optical_depth_ready = False
        self._shader = DEFAULT_SHADER
        self._raw_params = {}              # values as the user set them (get_shader_parameter returns these)
        self._t0 = time.monotonic()
        self.extra_cull_margin = 0.0
        self._update_cull_margin()
        # _init defaults (:105-108)
        self._p.sun_position[:] = (5000.0, 0.0, 0.0)
        self._p.clip_mode = 0.0
        if blue_noise is None:
            from .scenes import blue_noise_tile
            blue_noise = blue_noise_tile()
        self._ctx.upload_blue_noise(blue_noise)
        self._apply_variant()

    # ---- exported properties (:20-54) ----
    @property
    def planet_radius(self):
        return self._planet_radius

    @planet_radius.setter
    def planet_radius(self, v):
        self.set_planet_radius(v)

    @property
    def atmosphere_height(self):
        return self._atmosphere_height

    @atmosphere_height.setter
    def atmosphere_height(self, v):
        self.set_atmosphere_height(v)

    @property
    def sun_path(self):
        return self._sun_path

    @sun_path.setter
    def sun_path(self, v):
        self.set_sun_path(v)

    @property
    def custom_shader(self):
        return self._custom_shader

    @custom_shader.setter
    def custom_shader(self, v):
        self.set_custom_shader(v)

    # ---- :111-115 ----
    def _ready(self):
        self._p.planet_radius = self._planet_radius
        self._p.atmosphere_height = self._atmosphere_height

    # ---- :118-141 ----
    def set_custom_shader(self, shader):
        """`shader`: a key of SHADER_VARIANTS (with or without `.gdshader`), a (model, steps, cloud_steps, light)
        tuple for custom step counts, or None for the default shader."""
        self._custom_shader = shader
        if shader is None:
            self._shader = DEFAULT_SHADER
        elif isinstance(shader, str):
            name = shader.rsplit("/", 1)[-1].replace(".gdshader", "")
            if name == "planet_atmosphere_clouds_high_m":  # README.md:35 spelling of the _rm file
                name = "planet_atmosphere_clouds_high_rm"
            if name not in SHADER_VARIANTS:
                raise ValueError(f"unknown atmosphere shader {shader!r}; known: {sorted(SHADER_VARIANTS)}")
            self._shader = name
        else:
            self._shader = tuple(int(x) for x in shader)
        self._apply_variant()
        # the LUT is baked when the shader has a `u_optical_depth_texture` uniform = every v2 variant (:132-139)
        if self._variant()[0] == abi.SCATTER_V2:
            self._uses_baked_optical_depth = True
        if self._uses_baked_optical_depth:
            self._request_bake_optical_depth()

    def _variant(self):
        return SHADER_VARIANTS[self._shader] if isinstance(self._shader, str) else self._shader

    def _apply_variant(self):
        m, ns, nc, lm = self._variant()
        self._ctx.set_variant(ns, nc, lm, m)

    # ---- :144-156 ----
    def _request_bake_optical_depth(self):
        if self._optical_depth_baker is None:
            self._optical_depth_baker = OpticalDepthBaker(self._ctx)
            self._optical_depth_baker.connect_baked(self._on_optical_depth_baked)
        self._optical_depth_ready = False
        self._optical_depth_baker.request_bake(self._p)

    def _on_optical_depth_baked(self, _tex):
        self._optical_depth_ready = True

    # ---- :164-180 ----
    def set_shader_param(self, param_name, value):
        warnings.warn("set_shader_param is deprecated, use set_shader_parameter", DeprecationWarning, stacklevel=2)
        self.set_shader_parameter(param_name, value)

    def get_shader_param(self, param_name):
        warnings.warn("get_shader_param is deprecated, use get_shader_parameter", DeprecationWarning, stacklevel=2)
        return self.get_shader_parameter(param_name)

    def _uniform_table(self):
        m, _, _, lm = self._variant()
        t = dict(_V1_UNIFORMS if m == abi.SCATTER_V1 else _V2_UNIFORMS)
        if lm != abi.LIGHT_NONE:
            t.update(_CLOUD_UNIFORMS)
        return t

    def set_shader_parameter(self, param_name, value):
        name = str(param_name)
        self._raw_params[name] = value
        direct = {"u_planet_radius": "planet_radius", "u_atmosphere_height": "atmosphere_height", "u_clip_mode": "clip_mode"}
        if name in direct:
            setattr(self._p, direct[name], float(value))
            return
        if name == "u_sun_position":
            self._p.sun_position[:] = tuple(float(x) for x in value)
            return
        if name == "u_world_to_model_matrix":
            self._p.world_to_model[:] = flat_colmajor(value)
            return
        if name == "u_cloud_coverage_rotation":
            m = np.asarray(value, dtype=np.float64)  # 2x2, math convention
            self._p.cloud_coverage_rotation[:] = (m[0, 0], m[1, 0], m[0, 1], m[1, 1])
            return
        if name == "u_blue_noise_texture":
            self._ctx.upload_blue_noise(value)
            return
        if name == "u_optical_depth_texture":
            return  # owned by the context
        table = {**_V1_UNIFORMS, **_V2_UNIFORMS, **_CLOUD_UNIFORMS}
        if name not in table:
            return  # ShaderMaterial silently keeps unknown parameters
        field, kind = table[name]
        if kind == "f":
            setattr(self._p, field, float(value))
        elif kind == "v3":
            getattr(self._p, field)[:] = tuple(float(x) for x in value)
        elif kind == "c3":
            getattr(self._p, field)[:] = tuple(srgb_to_linear(list(value)[:3]).tolist())
        elif kind == "c4":
            v = list(value) + [1.0] * (4 - len(value))
            getattr(self._p, field)[:] = tuple(srgb_to_linear(v[:3]).tolist()) + (float(v[3]),)
        elif kind == "tex3d":
            self._ctx.upload_shape3d(value)
        elif kind == "cube":
            self._ctx.upload_coverage_cube(value)

    def get_shader_parameter(self, param_name):
        return self._raw_params.get(str(param_name))

    # ---- :185-218 — dynamic `shader_params/*` properties ----
    def _get_property_list(self):
        return [{"name": f"shader_params/{u}"} for u in self._uniform_table() if u not in _API_SHADER_PARAMS]

    def get(self, key):
        key = str(key)
        if key.startswith("shader_params/"):
            name = key[len("shader_params/"):]
            value = self.get_shader_parameter(name)
            if value is None:  # fall back to the shader default (:206-207)
                value = self._shader_default(name)
            return value
        return getattr(self, key)

    def set(self, key, value):
        key = str(key)
        if key.startswith("shader_params/"):
            name = key[len("shader_params/"):]
            self.set_shader_parameter(name, value)
            if self._uses_baked_optical_depth and name in _SHADER_PARAMS_AFFECTING_OPTICAL_DEPTH:
                self._request_bake_optical_depth()
            return
        setattr(self, key, value)

    @staticmethod
    def _shader_default(name):
        d = abi.default_params()
        table = {**_V1_UNIFORMS, **_V2_UNIFORMS, **_CLOUD_UNIFORMS}
        field, kind = table.get(name, (None, None))
        if field is None:
            return None
        v = getattr(d, field)
        return float(v) if kind == "f" else tuple(v)

    # ---- :221-227 ----
    def _get_configuration_warnings(self):
        if self._sun_path is None:
            return ["The path to the sun is not assigned."]
        if not (hasattr(self._sun_path, "global_transform") or np.shape(self._sun_path) == (3,)):
            return ["The assigned sun node is not a Node3D."]
        return []

    # ---- :230-258 ----
    def set_planet_radius(self, new_radius):
        if self._planet_radius == new_radius:
            return
        self._planet_radius = max(float(new_radius), 0.0)
        self._p.planet_radius = self._planet_radius
        self._update_cull_margin()
        if self._uses_baked_optical_depth:
            self._request_bake_optical_depth()

    def _update_cull_margin(self):
        self.extra_cull_margin = self._planet_radius + self._atmosphere_height

    def set_atmosphere_height(self, new_height):
        if self._atmosphere_height == new_height:
            return
        self._atmosphere_height = max(float(new_height), 0.0)
        self._p.atmosphere_height = self._atmosphere_height
        self._update_cull_margin()
        if self._uses_baked_optical_depth:
            self._request_bake_optical_depth()

    def set_sun_path(self, new_sun_path):
        self._sun_path = new_sun_path

    # ---- :261-282 ----
    def _set_mode(self, mode):
        if mode == self._mode:
            return
        self._mode = mode
        self._p.clip_mode = 1.0 if mode == MODE_NEAR else 0.0

    @property
    def mode(self):
        return self._mode

    # ---- :285-341 ----
    def _process(self, _delta=0.0, camera_position=None, camera_near=0.1, now=None):
        cam_pos = np.zeros(3) if camera_position is None else np.asarray(camera_position, dtype=np.float64)
        origin = np.asarray(self.global_transform, dtype=np.float64)[:3, 3]
        # 1.75 ~ sqrt(3): the far mesh is a cube (:300-303)
        atmo_clip_distance = 1.75 * (self._planet_radius + self._atmosphere_height + camera_near) * SWITCH_MARGIN_RATIO
        d = float(np.linalg.norm(origin - cam_pos))
        is_near = d < atmo_clip_distance
        self._set_mode(MODE_NEAR if (is_near or self.force_fullscreen) else MODE_FAR)
        if self._mode == MODE_FAR and self._prev_atmo_clip_distance != atmo_clip_distance:
            self._prev_atmo_clip_distance = atmo_clip_distance
            self._far_mesh_size = atmo_clip_distance
        if self._sun_path is not None:  # :328-331
            s = self._sun_path
            pos = np.asarray(s.global_transform, dtype=np.float64)[:3, 3] if hasattr(s, "global_transform") else np.asarray(s)
            self._p.sun_position[:] = tuple(float(x) for x in pos)
        # :335-336 — Transform3D.inverse(): transposed basis and -B^T * origin (Godot assumes an orthonormal basis there;
        # the node is never scaled, :315). Same arithmetic as the C++ core (csrc/node/planet_atmosphere_node.cpp).
        g = np.asarray(self.global_transform, dtype=np.float32)
        w2m = np.eye(4, dtype=np.float32)
        w2m[:3, :3] = g[:3, :3].T
        o = g[:3, 3].astype(np.float64)
        for r in range(3):
            b = w2m[r, :3].astype(np.float64)
            w2m[r, 3] = np.float32(-(b[0] * o[0] + b[1] * o[1] + b[2] * o[2]))
        self._p.world_to_model[:] = flat_colmajor(w2m)
        # :339-341 — Transform2D().rotated(a): columns (cos, sin), (-sin, cos)
        t = (time.monotonic() - self._t0) if now is None else float(now)
        a = t * math.radians(self.clouds_rotation_speed)
        self._p.cloud_coverage_rotation[:] = (math.cos(a), math.sin(a), -math.sin(a), math.cos(a))
        if self._optical_depth_baker is not None and self._optical_depth_baker.processing:
            self._optical_depth_baker._process(_delta)

    # ---- the draw call: fragment stage over the target (what the engine does after _process) ----
    def make_camera(self, inv_projection, inv_view, view=None, double_precision=False) -> abi.B200AtmoCamera:
        cam = abi.B200AtmoCamera()
        cam.inv_projection[:] = flat_colmajor(inv_projection)
        cam.inv_view[:] = flat_colmajor(inv_view)
        cam.view[:] = flat_colmajor(np.linalg.inv(np.asarray(inv_view, dtype=np.float64)) if view is None else view)
        cam.model[:] = flat_colmajor(self.global_transform)
        cam.double_precision = 1 if double_precision else 0
        # MODE_FAR draws the resized BoxMesh (:314-321); MODE_NEAR the fullscreen quad (u_clip_mode, :268-275)
        cam.clip_box_size = float(self._far_mesh_size) if self._mode == MODE_FAR else 0.0
        return cam

    def render(self, camera: abi.B200AtmoCamera, depth, width, height, rgba, discard=None, stream=None):
        """Device buffers (torch tensors / pointers): depth [h*w] f32 -> rgba [h*w*4] f32 (+ discard [h*w] u8)."""
        self._ctx.set_params(self._p)
        self._ctx.render_frame(camera, depth, width, height, rgba, discard, stream=stream)

    def render_host(self, camera: abi.B200AtmoCamera, depth, width, height, rgba, discard=None):
        self._ctx.set_params(self._p)
        self._ctx.render_frame_host(camera, depth, width, height, rgba, discard)

    @property
    def context(self) -> AtmosphereContext:
        return self._ctx

    @property
    def params(self) -> abi.B200AtmoParams:
        return self._p

    def free(self):
        self._ctx.close()
