"""B200-native batched atmosphere raymarcher (hot path of Zylann/godot_atmosphere_shader).

Only what the path needs lives here: `csrc/` (sm_100a CUDA kernels + the C-ABI of include/b200atmo.h),
`context.py` (ctypes front end of that C-ABI), `planet_atmosphere.py` (host-side mirror of the
reference's PlanetAtmosphere node surface) and `scenes.py` (synthetic inputs).
"""
from . import abi  # noqa: F401

__all__ = ["abi"]
