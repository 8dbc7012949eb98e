"""PCIe ceiling vs b200atmo_render_frame_host: raw pinned H2D/D2H of the e2e payload, and e2e time vs band count."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from godot_atmosphere_shader_b200 import context, scenes
w, h = 1920, 1080
p = scenes.demo_params(); cam = scenes.camera_b(w, h, p); depth = scenes.synth_depth(cam, p, w, h)
hd = torch.from_numpy(depth).pin_memory(); hr = torch.empty((h * w, 4), dtype=torch.float32).pin_memory()
dd = torch.empty_like(hd, device="cuda"); dr = torch.empty((h * w, 4), dtype=torch.float32, device="cuda")
def t(fn, n=30):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("raw H2D 8.3MB  %.3f ms" % t(lambda: dd.copy_(hd, non_blocking=True)))
print("raw D2H 33.2MB %.3f ms" % t(lambda: hr.copy_(dr, non_blocking=True)))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): dd.copy_(hd, non_blocking=True)
    with torch.cuda.stream(s2): hr.copy_(dr, non_blocking=True)
print("H2D || D2H     %.3f ms" % t(both))
for bands in (0, 1, 2, 3, 4, 5, 6, 8, 12):
    os.environ.pop("B200ATMO_E2E_BANDS", None)
    if bands: os.environ["B200ATMO_E2E_BANDS"] = str(bands)  # 0 = library default
    ctx = context.AtmosphereContext(0); ctx.set_params(p); ctx.set_variant(32, 0, 0); ctx.upload_blue_noise(scenes.blue_noise_tile())
    print("e2e bands=%2d    %.3f ms" % (bands, t(lambda: ctx.render_frame_host(cam, hd, w, h, hr, None), 20)))
    ctx.close()
