"""Host-side mirror of the reference's `NoiseCubemap` resource (addons/zylann.atmosphere/noise_cubemap.gd):
a procedural cubemap of 3D noise, generated on the GPU through `b200atmo_generate_noise_cubemap` instead of the
GDScript triple loop the reference calls "really slow" (noise_cubemap.gd:100).

Same properties and clamps: `noise`, `resolution` (1..4096, default 256), `scale` (default (100,100,100)); updates
are coalesced like `_request_update` / `call_deferred` (call `flush()` = the deferred `_update`); the image data
is never serialised (noise_cubemap.gd:84-90). `generate_importable_image()` builds the 3x2 atlas of
noise_cubemap.gd:143-155. The noise CONTENT is this repo's "b200 gradient fBm v1" (FastNoiseLite is engine
code outside the reference tree), see include/b200atmo.h.
"""
import numpy as np

from . import abi


class FastNoiseLiteParams:
    """The FastNoiseLite properties the generator reads; defaults are Godot's resource defaults."""

    def __init__(self, seed=0, frequency=0.01, fractal_octaves=5, fractal_lacunarity=2.0, fractal_gain=0.5):
        self.seed, self.frequency = int(seed), float(frequency)
        self.fractal_octaves, self.fractal_lacunarity, self.fractal_gain = int(fractal_octaves), float(fractal_lacunarity), float(fractal_gain)
        self._listeners = []

    def struct(self) -> abi.B200AtmoNoise:
        return abi.B200AtmoNoise(self.seed, self.frequency, self.fractal_octaves, self.fractal_lacunarity, self.fractal_gain)

    def emit_changed(self):  # Resource.changed
        for fn in list(self._listeners):
            fn()


class NoiseCubemap:
    def __init__(self, ctx, noise=None):
        self._ctx = ctx
        self._noise = None
        self._resolution = 256                     # noise_cubemap.gd:25
        self._scale = (100.0, 100.0, 100.0)        # :38
        self._update_scheduled = False
        self._images = None
        self.changed_callbacks = []
        self.noise = noise if noise is not None else FastNoiseLiteParams()   # _init: FastNoiseLite.new(), :51-54

    # ---- properties (:9-45) ----
    @property
    def noise(self):
        return self._noise

    @noise.setter
    def noise(self, value):
        if self._noise is not None and self._on_noise_changed in self._noise._listeners:
            self._noise._listeners.remove(self._on_noise_changed)
        self._noise = value
        if self._noise is not None:
            self._noise._listeners.append(self._on_noise_changed)
            self._request_update()

    @property
    def resolution(self):
        return self._resolution

    @resolution.setter
    def resolution(self, value):
        r = min(max(int(value), 1), 4096)          # clampi(value, 1, 4096), :30
        if r != self._resolution:
            self._resolution = r
            self._request_update()

    @property
    def scale(self):
        return self._scale

    @scale.setter
    def scale(self, value):
        v = tuple(float(x) for x in value)
        if v != self._scale:
            self._scale = v
            self._request_update()

    # ---- update machinery (:57-81) ----
    def _on_noise_changed(self):
        self._request_update()

    def _request_update(self):
        self._update_scheduled = True              # _update.call_deferred()

    def flush(self):
        """Runs the deferred `_update` if one is scheduled (the engine does this at the end of the frame)."""
        if self._update_scheduled:
            self._update()

    def _update(self):
        if self._noise is None:
            self._update_scheduled = False
            return
        self._images = self._ctx.generate_noise_cubemap(self._noise.struct(), self._resolution, self._scale, download=True,
                                                        set_as_coverage=False)
        self._update_scheduled = False
        for fn in self.changed_callbacks:          # emit_changed()
            fn()

    # ---- data access ----
    def get_layer_data(self, side: int) -> np.ndarray:
        self.flush()
        return self._images[side]

    def get_faces(self) -> np.ndarray:
        self.flush()
        return self._images

    def bind_as_coverage(self):
        """shader_params/u_cloud_coverage_cubemap = this resource: regenerates on the device and installs it."""
        self._ctx.generate_noise_cubemap(self._noise.struct(), self._resolution, self._scale, download=False, set_as_coverage=True)

    def generate_importable_image(self) -> np.ndarray:
        """3 x 2 atlas, side index = x + 3*y (noise_cubemap.gd:143-155)."""
        faces = self.get_faces()
        res = self._resolution
        im = np.empty((2 * res, 3 * res), dtype=np.uint8)
        for y in range(2):
            for x in range(3):
                im[y * res:(y + 1) * res, x * res:(x + 1) * res] = faces[x + 3 * y]
        return im
