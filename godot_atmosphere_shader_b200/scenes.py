"""Deterministic synthetic inputs for the hot path: scenes, cameras, depth buffers, textures.

Host-side input preparation only (numpy); nothing here is on the timed path. The reference's
parameter sets come from addons/zylann.atmosphere/planet_atmosphere.tscn:8-15 ("template") and
addons/zylann.atmosphere/demo/planet_atmosphere_test.tscn:96-114 ("demo").
"""
import math

import numpy as np

from .abi import IDENTITY16, B200AtmoCamera, B200AtmoFrame, B200AtmoParams, default_params


# ------------------------------------------------------------------------------------------------
# parameter sets
# ------------------------------------------------------------------------------------------------
def template_params() -> B200AtmoParams:
    """planet_atmosphere.tscn:8-15 (R=1, H=0.2, u_density=10, strength=0.5)."""
    p = default_params()
    p.planet_radius = 1.0
    p.atmosphere_height = 0.2
    p.density = 10.0
    p.scattering_strength = 0.5
    p.scattering_wavelengths[:] = (700.0, 530.0, 440.0)
    p.atmosphere_modulate[:] = (1.0, 1.0, 1.0)
    p.sphere_depth_factor = 0.0
    p.sun_position[:] = (5000.0, 0.0, 0.0)  # planet_atmosphere.gd:106
    return p


def demo_params() -> B200AtmoParams:
    """demo/planet_atmosphere_test.tscn:96-114; colour values are taken as already-linear."""
    p = default_params()
    p.planet_radius = 100.0
    p.atmosphere_height = 8.0
    p.density = 0.5
    p.scattering_strength = 1.0
    p.atmosphere_modulate[:] = (1.0, 0.980392, 0.964706)
    p.atmosphere_ambient_color[:] = (0.0196078, 0.0196078, 0.0431373)
    p.cloud_density_scale = 2.0
    p.cloud_bottom = 0.2
    p.cloud_top = 0.6
    p.cloud_blend = 0.5
    p.cloud_shape_invert = 1.0
    p.cloud_coverage_bias = 0.0
    p.cloud_shape_factor = 0.5
    p.cloud_shape_scale = 0.1
    p.sun_position[:] = (0.0, 0.0, 478.677)  # Sun (0,0,598.677) + DirectionalLight (0,0,-120)
    return p


# ------------------------------------------------------------------------------------------------
# matrices (math convention M[row, col]; flattened column-major for the ABI)
# ------------------------------------------------------------------------------------------------
def flat_colmajor(m) -> tuple:
    return tuple(np.asarray(m, dtype=np.float32).T.reshape(-1).tolist())


def perspective_reverse_z(fovy_deg: float, aspect: float, near: float, far: float) -> np.ndarray:
    """Godot >=4.3 Vulkan projection as seen by shaders: y flipped, depth 0..1 reversed (near -> 1, far -> 0)."""
    cot = 1.0 / math.tan(math.radians(fovy_deg) * 0.5)
    P = np.zeros((4, 4), dtype=np.float64)
    P[0, 0] = cot / aspect
    P[1, 1] = -cot
    P[2, 2] = near / (far - near)
    P[2, 3] = far * near / (far - near)
    P[3, 2] = -1.0
    return P


def godot_camera_projection(fovy_deg: float, aspect: float, near: float, far: float) -> np.ndarray:
    """Godot's Projection::set_perspective (core/math/projection.cpp): the GL-convention camera projection the engine keeps
    (y up, clip z in [-w, w]) — what RenderSceneData.get_cam_projection() returns to a CompositorEffect."""
    cot = 1.0 / math.tan(math.radians(fovy_deg) * 0.5)
    dz = far - near
    P = np.zeros((4, 4), dtype=np.float64)
    P[0, 0] = cot / aspect
    P[1, 1] = cot
    P[2, 2] = -(far + near) / dz
    P[3, 2] = -1.0
    P[2, 3] = -2.0 * near * far / dz
    return P


def godot_depth_correction(flip_y: bool = True, reverse_z: bool = True, remap_z: bool = True) -> np.ndarray:
    """Godot 4.3 Projection::set_depth_correction: the matrix RenderSceneDataRD::update_ubo multiplies in front of the camera
    projection before it uploads PROJECTION_MATRIX / INV_PROJECTION_MATRIX (y flip for Vulkan, reverse-Z, z remapped 0..1)."""
    M = np.eye(4)
    M[1, 1] = -1.0 if flip_y else 1.0
    M[2, 2] = (-0.5 if reverse_z else 0.5) if remap_z else (-1.0 if reverse_z else 1.0)
    M[2, 3] = 0.5 if remap_z else 0.0
    return M


def camera_transform(eye, forward, up=(0.0, 1.0, 0.0)) -> np.ndarray:
    """Camera-to-world transform (= INV_VIEW_MATRIX); the camera looks down its -Z."""
    f = np.asarray(forward, dtype=np.float64)
    f = f / np.linalg.norm(f)
    u = np.asarray(up, dtype=np.float64)
    r = np.cross(f, u)
    r = r / np.linalg.norm(r)
    u2 = np.cross(r, f)
    M = np.eye(4)
    M[:3, 0] = r
    M[:3, 1] = u2
    M[:3, 2] = -f
    M[:3, 3] = np.asarray(eye, dtype=np.float64)
    return M


def make_camera(eye, forward, up=(0, 1, 0), fovy_deg=75.0, aspect=16 / 9, near=0.1, far=800.0, model=None,
                double_precision=False) -> B200AtmoCamera:
    cam = B200AtmoCamera()
    P = perspective_reverse_z(fovy_deg, aspect, near, far)
    inv_view = camera_transform(eye, forward, up)
    view = np.linalg.inv(inv_view)
    cam.inv_projection[:] = flat_colmajor(np.linalg.inv(P))
    cam.inv_view[:] = flat_colmajor(inv_view)
    cam.view[:] = flat_colmajor(view)
    cam.model[:] = IDENTITY16 if model is None else flat_colmajor(model)
    cam.double_precision = 1 if double_precision else 0
    cam._meta = dict(P=P, inv_view=inv_view, near=near, far=far, fovy_deg=fovy_deg, aspect=aspect)
    return cam


def camera_a(w: int, h: int, orbit_deg: float = 0.0) -> B200AtmoCamera:
    """Camera A (orbit, realistic): the demo avatar pose, looking at the planet; optional orbit about +Y."""
    eye = np.array([0.357289, 0.105603, 157.92054])
    a = math.radians(orbit_deg)
    rot = np.array([[math.cos(a), 0, math.sin(a)], [0, 1, 0], [-math.sin(a), 0, math.cos(a)]])
    eye = rot @ eye
    fwd = rot @ np.array([0.0, 0.0, -1.0])
    return make_camera(eye, fwd, aspect=w / h)


def camera_b(w: int, h: int, params: B200AtmoParams = None) -> B200AtmoCamera:
    """Camera B (all-hit): 0.25*H above the surface on +Y, looking along +X; every ray starts inside the atmosphere."""
    p = params if params is not None else demo_params()
    eye = np.array([0.0, p.planet_radius + 0.25 * p.atmosphere_height, 0.0])
    return make_camera(eye, (1.0, 0.0, 0.0), up=(0, 1, 0), aspect=w / h)


def camera_c(w: int, h: int, params: B200AtmoParams = None, pitch_deg: float = 45.0) -> B200AtmoCamera:
    """Camera C (cloud deck): 0.9*H above the surface on the sunlit +Z axis, above the cloud shell, looking 45 degrees down
    into it — every ray hits the atmosphere and ~85 % of the pixels of the demo scene see cloud (camera B sees almost none)."""
    p = params if params is not None else demo_params()
    a = math.radians(pitch_deg)
    eye = np.array([0.0, 0.0, p.planet_radius + 0.9 * p.atmosphere_height])
    return make_camera(eye, (math.cos(a), 0.0, -math.sin(a)), up=(0, 0, 1), aspect=w / h)


def synth_depth(cam: B200AtmoCamera, params: B200AtmoParams, w: int, h: int, planet_center=(0.0, 0.0, 0.0)) -> np.ndarray:
    """Depth buffer a Godot opaque pass would leave: the ground sphere where hit, else the clear value 0 (far)."""
    m = cam._meta
    P, inv_view = m["P"], m["inv_view"]
    xs = (np.arange(w, dtype=np.float64) + 0.5) / w * 2.0 - 1.0
    ys = (np.arange(h, dtype=np.float64) + 0.5) / h * 2.0 - 1.0
    X, Y = np.meshgrid(xs, ys)
    # view-space direction through the pixel (z = -1 plane)
    dv = np.stack([X / P[0, 0], Y / P[1, 1], -np.ones_like(X)], axis=-1)
    dv /= np.linalg.norm(dv, axis=-1, keepdims=True)
    R3 = inv_view[:3, :3]
    eye = inv_view[:3, 3]
    dw = dv @ R3.T
    oc = eye - np.asarray(planet_center, dtype=np.float64)
    b = dw @ oc
    c = oc @ oc - float(params.planet_radius) ** 2
    disc = b * b - c
    hit = disc > 0
    t = np.where(hit, -b - np.sqrt(np.where(hit, disc, 0.0)), -1.0)
    hit &= t > m["near"]
    zv = t * dv[..., 2]  # negative
    A, B = P[2, 2], P[2, 3]
    zn = np.where(hit, (A * zv + B) / np.where(hit, -zv, 1.0), 0.0)
    return np.clip(zn, 0.0, 1.0).astype(np.float32)


def frame_constants(cam: B200AtmoCamera, params: B200AtmoParams) -> B200AtmoFrame:
    """float64 host computation of the atmosphere_vertex varyings — for tests that only need plausible constants."""
    m = cam._meta
    view = np.linalg.inv(m["inv_view"])
    model = np.array(cam.model[:], dtype=np.float64).reshape(4, 4).T
    fr = B200AtmoFrame()
    pc = view @ (model @ np.array([0, 0, 0, 1.0]))
    sc = view @ np.array([params.sun_position[0], params.sun_position[1], params.sun_position[2], 1.0])
    fr.planet_center_view[:] = tuple(np.float32(pc[:3]).tolist())
    fr.sun_center_view[:] = tuple(np.float32(sc[:3]).tolist())
    fr.inv_view[:] = cam.inv_view[:]
    return fr


# ------------------------------------------------------------------------------------------------
# textures (contents are NOT pinned by the reference: engine-side FastNoiseLite / imported PNG)
# ------------------------------------------------------------------------------------------------
def blue_noise_tile(size: int = 256, seed: int = 12345) -> np.ndarray:
    """size x size u8 jitter tile with the reference tile's histogram (every value 0..255 equally often).

    The reference ships blue_noise.png (256x256, each grey level exactly 256 times); it is an input
    texture, not code, and is regenerated here instead of copied (a seeded permutation: same value
    distribution, white instead of blue spectrum — irrelevant to the arithmetic under test).
    """
    rng = np.random.default_rng(seed)
    vals = np.repeat(np.arange(256, dtype=np.uint8), (size * size) // 256)
    rng.shuffle(vals)
    return vals.reshape(size, size)


def _value_noise_3d(n: int, period: int, rng) -> np.ndarray:
    """Tileable trilinear value noise on an n^3 grid with `period` lattice cells per axis."""
    lat = rng.random((period, period, period))
    t = (np.arange(n) + 0.5) / n * period
    i0 = np.floor(t).astype(int) % period
    i1 = (i0 + 1) % period
    f = t - np.floor(t)
    f = f * f * (3 - 2 * f)

    def ax(a, axis):
        sl0 = np.take(a, i0, axis=axis)
        sl1 = np.take(a, i1, axis=axis)
        shape = [1, 1, 1]
        shape[axis] = n
        ff = f.reshape(shape)
        return sl0 * (1 - ff) + sl1 * ff

    return ax(ax(ax(lat, 0), 1), 2)


def shape_texture(n: int = 64, seed: int = 1, octaves: int = 4) -> np.ndarray:
    """n^3 u8 tileable fBm (stand-in for Godot's seamless NoiseTexture3D, demo .tscn:55-57). Index [z][y][x]."""
    rng = np.random.default_rng(seed)
    acc = np.zeros((n, n, n))
    amp, tot, period = 1.0, 0.0, 4
    for _ in range(octaves):
        acc += amp * _value_noise_3d(n, min(period, n), rng)
        tot += amp
        amp *= 0.6
        period *= 2
    acc /= tot
    acc = (acc - acc.min()) / (acc.max() - acc.min())
    return np.round(acc * 255.0).astype(np.uint8)


def cube_texel_directions(res: int) -> np.ndarray:
    """Unit direction of every texel centre, [6][res][res][3], per noise_cubemap.gd:110-128."""
    x = np.arange(res) + 0.5
    px = x / (res * 0.5) - 1.0
    py = (res - np.arange(res) - 1 + 0.5) / (res * 0.5) - 1.0
    PX, PY = np.meshgrid(px, py)  # [y][x]
    one = np.ones_like(PX)
    base = np.stack([one, PY, -PX], axis=-1)
    base /= np.linalg.norm(base, axis=-1, keepdims=True)
    bx, by, bz = base[..., 0], base[..., 1], base[..., 2]
    faces = [
        np.stack([bx, by, bz], -1),     # +X
        np.stack([-bx, by, -bz], -1),   # -X
        np.stack([-bz, bx, -by], -1),   # +Y
        np.stack([-bz, -bx, by], -1),   # -Y
        np.stack([-bz, by, bx], -1),    # +Z
        np.stack([bz, by, -bx], -1),    # -Z
    ]
    return np.stack(faces, 0)


def coverage_cubemap(res: int = 256, seed: int = 1) -> np.ndarray:
    """6 x res x res u8 coverage faces: 3D fBm sampled on the unit sphere (stand-in for NoiseCubemap + FastNoiseLite)."""
    rng = np.random.default_rng(seed + 1000)
    dirs = cube_texel_directions(res)
    acc = np.zeros(dirs.shape[:-1])
    amp, tot, freq = 1.0, 0.0, 2.0
    for _ in range(5):
        period = 64
        lat = rng.random((period, period, period))
        p = (dirs * freq * 0.5 + 0.5 * freq + 7.0) % period
        i0 = np.floor(p).astype(int) % period
        i1 = (i0 + 1) % period
        f = p - np.floor(p)
        f = f * f * (3 - 2 * f)
        c = 0.0
        for dz in (0, 1):
            for dy in (0, 1):
                for dx in (0, 1):
                    ix = i1[..., 0] if dx else i0[..., 0]
                    iy = i1[..., 1] if dy else i0[..., 1]
                    iz = i1[..., 2] if dz else i0[..., 2]
                    wgt = (f[..., 0] if dx else 1 - f[..., 0]) * (f[..., 1] if dy else 1 - f[..., 1]) * (
                        f[..., 2] if dz else 1 - f[..., 2])
                    c = c + wgt * lat[iz, iy, ix]
        acc += amp * c
        tot += amp
        amp *= 0.55
        freq *= 2.0
    acc /= tot
    acc = (acc - acc.min()) / (acc.max() - acc.min())
    return np.round(acc * 255.0).astype(np.uint8)
