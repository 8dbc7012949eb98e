"""Engine-side conventions the GDExtension wrapper (gdextension/planet_atmosphere_b200.cpp, source only) must reproduce when it
fills B200AtmoCamera: INV_PROJECTION_MATRIX is the inverse of (depth correction x camera projection), not of the camera
projection Godot hands to a CompositorEffect. Checked here on the host against the oracle's ray generator."""
import numpy as np

from godot_atmosphere_shader_b200 import scenes
from oracle import pyoracle as O


def test_shader_projection_is_correction_times_camera_projection():
    for fov, aspect, near, far in ((75.0, 16 / 9, 0.1, 800.0), (40.0, 1.0, 0.05, 4000.0)):
        shader_p = scenes.godot_depth_correction(True) @ scenes.godot_camera_projection(fov, aspect, near, far)
        assert np.allclose(shader_p, scenes.perspective_reverse_z(fov, aspect, near, far), rtol=1e-12, atol=1e-15)


def test_uncorrected_projection_mirrors_the_rays_and_breaks_the_depth():
    """A known view-space point -> its pixel and reverse-Z depth under the engine's shader projection; the oracle's ray
    through that pixel must come back to the point with the corrected matrix and must NOT with cam_projection.inverse()."""
    w, h = 64, 36
    fov, aspect, near, far = 75.0, w / h, 0.1, 800.0
    p = scenes.demo_params()
    cam = scenes.make_camera((0.0, 0.0, 157.9), (0.0, 0.0, -1.0), fovy_deg=fov, aspect=aspect, near=near, far=far)
    shader_p = scenes.godot_depth_correction(True) @ scenes.godot_camera_projection(fov, aspect, near, far)
    assert np.allclose(np.array(cam.inv_projection[:]).reshape(4, 4).T, np.linalg.inv(shader_p), rtol=1e-5, atol=1e-7)
    point_view = np.array([3.0, 5.0, -40.0, 1.0])          # above and right of the optical axis, 40 units ahead
    clip = shader_p @ point_view
    ndc = clip[:3] / clip[3]
    assert 0.0 < ndc[2] < 1.0 and ndc[1] < 0.0             # reverse-Z depth in (0,1); +y in view space is the UPPER half = negative ndc.y
    px = int((ndc[0] * 0.5 + 0.5) * w)
    py = int((ndc[1] * 0.5 + 0.5) * h)
    assert py < h // 2                                     # row 0 = top
    depth = np.zeros((h, w), np.float32)
    depth[py, px] = ndc[2]
    tex = O.Textures(lut=O.bake_lut(p), blue_noise=scenes.blue_noise_tile())
    od, dj, _ = O.make_rays(p, cam, tex, depth, w, h)
    i = py * w + px
    d, lin = dj[i, :3].astype(np.float64), float(od[i, 3])
    want_dir = point_view[:3] / np.linalg.norm(point_view[:3])
    assert np.dot(d, want_dir) > 0.999                     # within the pixel's footprint
    assert abs(lin - np.linalg.norm(point_view[:3])) < 1e-2 * np.linalg.norm(point_view[:3])
    # the bug the advisor found: cam_projection.inverse() without the correction
    bad = scenes.make_camera((0.0, 0.0, 157.9), (0.0, 0.0, -1.0), fovy_deg=fov, aspect=aspect, near=near, far=far)
    bad.inv_projection[:] = scenes.flat_colmajor(np.linalg.inv(scenes.godot_camera_projection(fov, aspect, near, far)))
    od2, dj2, _ = O.make_rays(p, bad, tex, depth, w, h)
    d2 = dj2[i, :3].astype(np.float64)
    assert d2[1] * want_dir[1] < 0.0                       # mirrored vertically
    assert abs(float(od2[i, 3]) - np.linalg.norm(point_view[:3])) > 1.0   # and the linear depth is wrong
