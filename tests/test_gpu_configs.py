"""GPU parity at the sizes BASELINE.json configures (the full frame is rendered on the GPU, a row-strided sample of >= 20 000
rays is checked against the oracle AND, where it was built, against the reference's own shader sources compiled as C++),
plus the round-2 C-ABI additions: RGBA16F results, tile-mapped ray batches, the per-stream ray-table cache.

Tolerance as everywhere: |got - want| <= 1e-4*|want| + 2e-6 per channel, discard masks bit-exact; RGBA16F results must be
the round-to-nearest-even half of the fp32 result, bit for bit."""
import numpy as np
import pytest

from godot_atmosphere_shader_b200 import abi, scenes, sharding
from oracle import pyoracle as O
from oracle import pyref as R
from tests import helpers as Hh

pytestmark = pytest.mark.gpu


def _torch():
    import torch
    return torch


def _camera(name, w, h, p):
    return {"A": lambda: scenes.camera_a(w, h), "B": lambda: scenes.camera_b(w, h, p), "C": lambda: scenes.camera_c(w, h, p)}[name]()


_TEX = {}


def _bench_textures():
    """The textures bench.py uses: 64^3 shape, 6 x 256^2 coverage, 256^2 jitter tile."""
    if not _TEX:
        _TEX["t"] = (scenes.shape_texture(64, seed=1), scenes.coverage_cubemap(256, seed=1), scenes.blue_noise_tile())
    return _TEX["t"]


def _setup(ctx, p, ns, nc, lm):
    shape, cube, bn = _bench_textures()
    ctx.set_params(p)
    ctx.set_variant(ns, nc, lm)
    ctx.upload_blue_noise(bn)
    ctx.upload_shape3d(shape)
    ctx.upload_coverage_cube(cube)
    return O.Textures(lut=O.bake_lut(p), shape=shape, cube_faces=cube, blue_noise=bn)


# (name, width, height, scatter steps, cloud steps, light, camera, row stride)  — BASELINE.json configs[1..3] as bench.py runs them
CONFIGS = [
    ("cfg2_camB", 1920, 1080, 32, 0, abi.LIGHT_NONE, "B", 97),
    ("cfg2_camA", 1920, 1080, 32, 0, abi.LIGHT_NONE, "A", 97),
    ("cfg3_camA", 1920, 1080, 8, 64, abi.LIGHT_CHEAP, "A", 97),
    ("cfg3_camC", 1920, 1080, 8, 64, abi.LIGHT_CHEAP, "C", 97),
    ("cfg3_n32_camC", 1920, 1080, 32, 64, abi.LIGHT_CHEAP, "C", 97),
    ("cfg4_camA", 3840, 2160, 8, 128, abi.LIGHT_RAYMARCHED, "A", 359),
    ("cfg4_camC", 3840, 2160, 8, 128, abi.LIGHT_RAYMARCHED, "C", 359),
]


@pytest.mark.parametrize("cfg", CONFIGS, ids=[c[0] for c in CONFIGS])
def test_configured_size_sample_parity(cuda_ctx_factory, cfg):
    torch = _torch()
    name, w, h, ns, nc, lm, cam_name, stride = cfg
    ctx = cuda_ctx_factory()
    p = scenes.demo_params()
    tex = _setup(ctx, p, ns, nc, lm)
    cam = _camera(cam_name, w, h, p)
    depth = scenes.synth_depth(cam, p, w, h)
    d_depth = torch.from_numpy(depth).cuda()
    d_rgba = torch.full((h, w, 4), -7.0, dtype=torch.float32, device="cuda")
    d_disc = torch.full((h, w), 9, dtype=torch.uint8, device="cuda")
    ctx.render_frame(cam, d_depth, w, h, d_rgba, d_disc)
    # the ray-batch API (what bench.py times) on the same frame: linear and tile-mapped, bit-identical to the frame API
    d_od = torch.empty((h * w, 4), dtype=torch.float32, device="cuda")
    d_dj = torch.empty((h * w, 4), dtype=torch.float32, device="cuda")
    fr = ctx.make_rays(cam, d_depth, w, h, d_od, d_dj)
    lin = torch.empty((h, w, 4), dtype=torch.float32, device="cuda")
    til = torch.empty((h, w, 4), dtype=torch.float32, device="cuda")
    ctx.render_rays(fr, d_od, d_dj, h * w, lin, None)
    ctx.render_rays(fr, d_od, d_dj, h * w, til, None, grid=(w, h))
    torch.cuda.synchronize()
    assert torch.equal(lin, d_rgba) and torch.equal(til, d_rgba)
    rows = np.arange(stride // 2, h, stride)
    assert len(rows) * w >= 20000
    got = d_rgba.cpu().numpy()[rows]
    gdisc = d_disc.cpu().numpy()[rows]
    var = O.variant(ns, nc, lm)
    od, dj, ofr = O.make_rays(p, cam, tex, depth, w, h)
    sel = (rows[:, None] * w + np.arange(w)[None, :]).reshape(-1)
    ref, rdisc = O.render_rays(p, var, ofr, tex, od[sel], dj[sel], threads=0)
    assert np.array_equal(gdisc.reshape(-1), rdisc), f"{name}: discard mask differs from the oracle"
    Hh.assert_rgba_close(got.reshape(-1, 4), ref, what=f"{name} vs oracle")
    if cam_name == "C" and nc:   # the camera exists to exercise the cloud path: most sampled pixels must see cloud
        plain, _ = O.render_rays(p, O.variant(ns, 0, 0), ofr, tex, od[sel][::7], dj[sel][::7], threads=0)
        assert (np.abs(plain - ref[::7]).max(axis=1) > 1e-6).mean() > 0.6
    if R.available():
        # the reference's own GDShader sources compiled as C++, every `stride`-th row starting at rows[0]
        k0 = int(rows[0])
        ref2, rdisc2 = R.render_frame(p, var, cam, tex, depth, w, h, row_begin=k0, threads=0, row_stride=stride)
        assert np.array_equal(gdisc, rdisc2[rows]), f"{name}: discard mask differs from the compiled reference shaders"
        Hh.assert_rgba_close(got, ref2[rows], what=f"{name} vs compiled reference")


def test_rgba16f_results_are_the_rounded_fp32_results(cuda_ctx_factory):
    """Every RGBA16F output path (device frame, host frame, pipelined host frame, peer tiles by store / TMA) holds exactly
    float16(round-to-nearest-even(fp32 result)); the fp32 path is the parity path."""
    torch = _torch()
    ctx = cuda_ctx_factory()
    w, h = 200, 121
    p = scenes.demo_params()
    _setup(ctx, p, 8, 32, abi.LIGHT_CHEAP)
    cam = scenes.camera_c(w, h, p)
    depth = scenes.synth_depth(cam, p, w, h)
    d_depth = torch.from_numpy(depth).cuda()
    f32 = torch.empty((h, w, 4), dtype=torch.float32, device="cuda")
    disc = torch.empty((h, w), dtype=torch.uint8, device="cuda")
    ctx.render_frame(cam, d_depth, w, h, f32, disc)
    f16 = torch.full((h, w, 4), -7.0, dtype=torch.float16, device="cuda")
    disc16 = torch.empty_like(disc)
    ctx.render_frame(cam, d_depth, w, h, f16, disc16, rgba_format=abi.COLOR_RGBA16F)
    torch.cuda.synchronize()
    want = f32.cpu().numpy().astype(np.float16)      # numpy rounds to nearest-even
    assert float(np.abs(f32.cpu().numpy()).max()) > 0.5
    assert np.array_equal(f16.cpu().numpy().view(np.uint16), want.view(np.uint16))
    assert torch.equal(disc, disc16)
    # row range: only the rows asked for are written
    part = torch.full((h, w, 4), -7.0, dtype=torch.float16, device="cuda")
    ctx.render_frame(cam, d_depth, w, h, part, None, row_begin=40, row_end=77, rgba_format=abi.COLOR_RGBA16F)
    torch.cuda.synchronize()
    pn = part.cpu().numpy()
    assert np.array_equal(pn[40:77].view(np.uint16), want[40:77].view(np.uint16)) and np.all(pn[:40] == -7.0) and np.all(pn[77:] == -7.0)
    # host paths
    h16 = np.full((h, w, 4), -7.0, np.float16)
    hd = np.empty((h, w), np.uint8)
    ctx.render_frame_host(cam, depth, w, h, h16, hd, rgba_format=abi.COLOR_RGBA16F)
    assert np.array_equal(h16.view(np.uint16), want.view(np.uint16)) and np.array_equal(hd, disc.cpu().numpy())
    p16 = [torch.full((h, w, 4), -7.0, dtype=torch.float16).pin_memory() for _ in range(2)]
    hdep = torch.from_numpy(depth).pin_memory()
    for k in range(4):
        ctx.frame_wait(k & 1)
        ctx.render_frame_host_submit(cam, hdep, w, h, p16[k & 1], None, slot=k & 1, rgba_format=abi.COLOR_RGBA16F)
    ctx.frame_wait(0)
    ctx.frame_wait(1)
    for b in p16:
        assert np.array_equal(b.numpy().view(np.uint16), want.view(np.uint16))
    # peer tiles in half format: P2P stores (frame + ray kernels) and the TMA bulk-store flavour with a ragged tail
    bufs = [torch.full((2, h * w, 4), -7.0, dtype=torch.float16, device="cuda") for _ in range(2)]
    t = sharding.peer_targets([b.data_ptr() for b in bufs], elem_offset=h * w, rgba_format=abi.COLOR_RGBA16F)
    ctx.render_frame_peers(cam, d_depth, w, h, t)
    torch.cuda.synchronize()
    for b in bufs:
        bn_ = b.cpu().numpy()
        assert np.array_equal(bn_[1].view(np.uint16), want.reshape(-1, 4).view(np.uint16)) and np.all(bn_[0] == -7.0)
    d_od = torch.empty((h * w, 4), dtype=torch.float32, device="cuda")
    d_dj = torch.empty((h * w, 4), dtype=torch.float32, device="cuda")
    fr = ctx.make_rays(cam, d_depth, w, h, d_od, d_dj)
    n_tail = 128 * 40 + 37
    for use_tma in (False, True):
        for b in bufs:
            b.fill_(-7.0)
        t2 = sharding.peer_targets([b.data_ptr() for b in bufs], elem_offset=0, first_peer=1, use_tma=use_tma, rgba_format=abi.COLOR_RGBA16F)
        ctx.render_rays_peers(fr, d_od, d_dj, n_tail, t2)
        torch.cuda.synchronize()
        for b in bufs:
            bn_ = b.cpu().numpy()
            assert np.array_equal(bn_[0][:n_tail].view(np.uint16), want.reshape(-1, 4)[:n_tail].view(np.uint16)), f"tma={use_tma}"
            assert np.all(bn_[0][n_tail:] == -7.0) and np.all(bn_[1] == -7.0)
    # deliver-to-root = a peer table that names only the consuming rank's buffer
    for b in bufs:
        b.fill_(-7.0)
    t3 = sharding.peer_targets([bufs[1].data_ptr()], elem_offset=0, rgba_format=abi.COLOR_RGBA16F)
    ctx.render_frame_peers(cam, d_depth, w, h, t3, row_begin=0, row_end=h)
    torch.cuda.synchronize()
    assert np.array_equal(bufs[1].cpu().numpy()[0].view(np.uint16), want.reshape(-1, 4).view(np.uint16)) and bool((bufs[0] == -7.0).all())
    from godot_atmosphere_shader_b200.context import B200AtmoError
    with pytest.raises(B200AtmoError):
        ctx.render_frame(cam, d_depth, w, h, f16, None, rgba_format=7)


@pytest.mark.parametrize("size", [(203, 77), (16, 8), (1, 1), (640, 360)])
def test_tile_mapped_ray_batch_is_bit_identical(cuda_ctx_factory, size):
    """b200atmo_render_rays_2d (warps cover 8x4 pixel tiles) == b200atmo_render_rays, also for sizes that are not
    multiples of the 16x8 block tile; the discard mask too."""
    torch = _torch()
    ctx = cuda_ctx_factory()
    w, h = size
    p = scenes.demo_params()
    _setup(ctx, p, 8, 64, abi.LIGHT_RAYMARCHED)
    cam = scenes.camera_a(w, h)
    depth = scenes.synth_depth(cam, p, w, h)
    d_depth = torch.from_numpy(depth).cuda()
    d_od = torch.empty((h * w, 4), dtype=torch.float32, device="cuda")
    d_dj = torch.empty((h * w, 4), dtype=torch.float32, device="cuda")
    fr = ctx.make_rays(cam, d_depth, w, h, d_od, d_dj)
    a = torch.full((h * w, 4), -7.0, dtype=torch.float32, device="cuda")
    b = torch.full((h * w, 4), -8.0, dtype=torch.float32, device="cuda")
    da = torch.full((h * w,), 9, dtype=torch.uint8, device="cuda")
    db = torch.full((h * w,), 8, dtype=torch.uint8, device="cuda")
    ctx.render_rays(fr, d_od, d_dj, h * w, a, da)
    ctx.render_rays(fr, d_od, d_dj, h * w, b, db, grid=(w, h))
    torch.cuda.synchronize()
    assert torch.equal(a, b) and torch.equal(da, db)


def test_ray_table_cache_alternating_viewports(cuda_ctx_factory):
    """Two viewports (1920x1080 and 3840x2160, different projections) alternating over the two pipeline slots and over two
    caller streams: after the first frame of each (stream, viewport) pair no ray table is ever rebuilt, and the frames are
    bit-identical to a fresh context's."""
    torch = _torch()
    ctx = cuda_ctx_factory()
    p = scenes.demo_params()
    _setup(ctx, p, 8, 0, abi.LIGHT_NONE)
    views = []
    for (w, h, orbit) in ((1920, 1080, 0.0), (3840, 2160, 30.0)):
        cam = scenes.camera_a(w, h, orbit)
        depth = torch.from_numpy(scenes.synth_depth(cam, p, w, h)).pin_memory()
        views.append((cam, depth, w, h, torch.empty((h, w, 4), dtype=torch.float32).pin_memory()))
    want = []
    fresh = cuda_ctx_factory()
    _setup(fresh, p, 8, 0, abi.LIGHT_NONE)
    for cam, depth, w, h, _ in views:
        out = np.empty((h, w, 4), np.float32)
        fresh.render_frame_host(cam, depth, w, h, out, None)
        want.append(out)
    # pipelined host frames: viewport k on slot k (each slot owns a stream)
    for rnd in range(4):
        for k, (cam, depth, w, h, out) in enumerate(views):
            ctx.frame_wait(k)
            ctx.render_frame_host_submit(cam, depth, w, h, out, None, slot=k)
        if rnd == 0:
            builds_after_first_pair = ctx.table_build_count
            launches_after_first_pair = ctx.launch_count
    ctx.frame_wait(0)
    ctx.frame_wait(1)
    assert builds_after_first_pair == 2
    assert ctx.table_build_count == builds_after_first_pair                      # zero rebuilds
    assert ctx.launch_count == launches_after_first_pair + 3 * 2                  # exactly one render kernel per later frame
    for (cam, depth, w, h, out), ref in zip(views, want):
        assert np.array_equal(out.numpy().view(np.uint32), ref.view(np.uint32))
    # both viewports alternating on ONE caller stream: the per-stream cache holds both
    s = torch.cuda.Stream()
    d_views = [(cam, depth.cuda(), w, h, torch.empty((h, w, 4), dtype=torch.float32, device="cuda")) for cam, depth, w, h, _ in views]
    b0 = ctx.table_build_count
    for rnd in range(3):
        for cam, d_depth, w, h, d_out in d_views:
            ctx.render_frame(cam, d_depth, w, h, d_out, None, stream=s.cuda_stream)
    s.synchronize()
    assert ctx.table_build_count == b0 + 2
    for (cam, d_depth, w, h, d_out), ref in zip(d_views, want):
        assert np.array_equal(d_out.cpu().numpy().view(np.uint32), ref.view(np.uint32))
    # a projection that changes every frame (TAA jitter): one small build per frame, stream-ordered, still correct
    cam, d_depth, w, h, d_out = d_views[0]
    b1 = ctx.table_build_count
    for j in range(6):
        camj = scenes.camera_a(w, h, 0.0)
        camj.inv_projection[12] += 1e-4 * (j + 1)     # sub-pixel shift of the projection
        ctx.render_frame(camj, d_depth, w, h, d_out, None, stream=s.cuda_stream)
    ctx.render_frame(cam, d_depth, w, h, d_out, None, stream=s.cuda_stream)
    s.synchronize()
    assert ctx.table_build_count >= b1 + 6
    assert np.array_equal(d_out.cpu().numpy().view(np.uint32), want[0].view(np.uint32))


def test_blue_noise_window_is_256_whatever_the_texture_size(cuda_ctx_factory):
    """planet_atmosphere_main.gdshaderinc:168-169 masks the pixel with 0xff: a 512x512 jitter texture contributes only its
    top-left 256x256 texels; textures smaller than 256 are rejected."""
    torch = _torch()
    from godot_atmosphere_shader_b200.context import B200AtmoError
    ctx = cuda_ctx_factory()
    w, h = 600, 300
    p = scenes.demo_params()
    rng = np.random.default_rng(3)
    big = rng.integers(0, 256, size=(512, 512), dtype=np.uint8)
    ctx.set_params(p)
    ctx.set_variant(8, 0, abi.LIGHT_NONE)
    ctx.upload_blue_noise(big)
    cam = scenes.camera_b(w, h, p)
    depth = scenes.synth_depth(cam, p, w, h)
    tex = O.Textures(lut=O.bake_lut(p), blue_noise=big)
    od, dj, _ = O.make_rays(p, cam, tex, depth, w, h)
    d_depth = torch.from_numpy(depth).cuda()
    d_od = torch.empty((h * w, 4), dtype=torch.float32, device="cuda")
    d_dj = torch.empty((h * w, 4), dtype=torch.float32, device="cuda")
    ctx.make_rays(cam, d_depth, w, h, d_od, d_dj)
    torch.cuda.synchronize()
    assert np.array_equal(d_dj.cpu().numpy().view(np.uint32), dj.view(np.uint32))
    jit = dj[:, 3].reshape(h, w)
    assert np.array_equal(jit[:, 256:512], jit[:, 0:256]) and np.array_equal(jit[256:300], jit[0:44])   # period 256, not 512
    assert np.array_equal(jit[:256, :256], (big[:256, :256].astype(np.float32) / np.float32(255.0))[:256, :256])
    for bad in ((128, 128), (256, 100)):
        with pytest.raises(B200AtmoError):
            ctx.upload_blue_noise(np.zeros(bad, np.uint8))


@pytest.mark.parametrize("clouds", [True, False], ids=["clouds", "scatter_only"])
@pytest.mark.parametrize("world", [1, 3, 8])
def test_interleaved_row_tiles_reassemble_the_frame(cuda_ctx_factory, world, clouds):
    """b200atmo_render_frame_peers_interleaved: rank g of G renders the 8-row tiles g, g+G, ...; the G launches together
    write every pixel exactly once with the bits of the unsharded frame (height not a multiple of 8, more ranks than
    some frames have tiles). The scatter-only peers kernel maps warps to 16x2 pixels, the cloud ones to 8x4: both are checked,
    in both tile formats."""
    torch = _torch()
    ctx = cuda_ctx_factory()
    p = scenes.demo_params()
    if clouds:
        _setup(ctx, p, 8, 32, abi.LIGHT_CHEAP)
    else:
        _setup(ctx, p, 8, 0, abi.LIGHT_NONE)
    for (w, h) in ((200, 121), (64, 20)):
        cam = scenes.camera_a(w, h)
        d_depth = torch.from_numpy(scenes.synth_depth(cam, p, w, h)).cuda()
        want = torch.empty((h, w, 4), dtype=torch.float32, device="cuda")
        ctx.render_frame(cam, d_depth, w, h, want, None)
        buf = torch.full((h, w, 4), -7.0, dtype=torch.float32, device="cuda")
        t = sharding.peer_targets([buf.data_ptr()])
        seen = np.zeros(h, np.int32)
        for g in range(world):
            before = buf.clone()
            ctx.render_frame_peers_interleaved(cam, d_depth, w, h, t, g, world)
            torch.cuda.synchronize()
            changed = (buf != before).any(dim=2).any(dim=1).cpu().numpy()
            rows = sharding.interleaved_rows(h, g, world)
            assert not changed[np.setdiff1d(np.arange(h), rows)].any()      # only this rank's rows are touched
            seen[rows] += 1
        assert (seen == 1).all() and torch.equal(buf, want)
        half = torch.full((h, w, 4), -7.0, dtype=torch.float16, device="cuda")
        t16 = sharding.peer_targets([half.data_ptr()], rgba_format=abi.COLOR_RGBA16F)
        for g in range(world):
            ctx.render_frame_peers_interleaved(cam, d_depth, w, h, t16, g, world)
        ctx.render_frame_peers(cam, d_depth, w, h, sharding.peer_targets([buf.data_ptr()]), row_begin=3, row_end=h - 2)   # a band that is not tile aligned
        torch.cuda.synchronize()
        assert torch.equal(half, want.to(torch.float16)) and torch.equal(buf, want)
    from godot_atmosphere_shader_b200.context import B200AtmoError
    with pytest.raises(B200AtmoError):
        ctx.render_frame_peers_interleaved(cam, d_depth, w, h, t, 3, 3)


def test_fused_handshake_and_flag_calls_on_one_gpu(cuda_ctx_factory):
    """B200AtmoPeerSync on a single GPU (this rank is its own producer and consumer): the render kernel publishes the
    "consumed" epoch at its start, the completion epoch from its last block, and only ends once the awaited flags are there;
    b200atmo_peers_signal / b200atmo_peers_wait do the same as stand-alone calls. Pixels are those of the plain call."""
    torch = _torch()
    ctx = cuda_ctx_factory()
    w, h = 200, 121
    p = scenes.demo_params()
    _setup(ctx, p, 8, 32, abi.LIGHT_CHEAP)
    cam = scenes.camera_a(w, h)
    d_depth = torch.from_numpy(scenes.synth_depth(cam, p, w, h)).cuda()
    want = torch.empty((h, w, 4), dtype=torch.float32, device="cuda")
    ctx.render_frame(cam, d_depth, w, h, want, None)
    flags = torch.zeros(16, dtype=torch.int32, device="cuda")
    buf = torch.full((h, w, 4), -7.0, dtype=torch.float32, device="cuda")
    d_od = torch.empty((h * w, 4), dtype=torch.float32, device="cuda")
    d_dj = torch.empty((h * w, 4), dtype=torch.float32, device="cuda")
    fr = ctx.make_rays(cam, d_depth, w, h, d_od, d_dj)
    for epoch, api in ((1, "frame"), (2, "rays"), (3, "interleaved")):
        buf.fill_(-7.0)
        t = sharding.peer_targets([buf.data_ptr()])
        y = t.sync
        y.d_done_flags[0] = flags.data_ptr()
        y.n_done_flags, y.done_slot, y.epoch = 1, 0, epoch
        y.d_wait_flags, y.wait_first_slot, y.n_wait = flags.data_ptr(), 0, 1           # wait for my own completion flag
        if epoch > 1:
            y.d_consumed_flags[0] = flags.data_ptr()
            y.n_consumed_flags, y.consumed_slot, y.consumed_epoch = 1, 8, epoch - 1
            y.d_credit_flags, y.credit_first_slot, y.n_credit, y.credit_epoch = flags.data_ptr(), 8, 1, epoch - 1   # published by block 0 of this very kernel
        if api == "frame":
            ctx.render_frame_peers(cam, d_depth, w, h, t)
        elif api == "rays":
            ctx.render_rays_peers(fr, d_od, d_dj, h * w, t)
        else:
            ctx.render_frame_peers_interleaved(cam, d_depth, w, h, t, 0, 1)
        torch.cuda.synchronize()
        f = flags.cpu().numpy()
        assert f[0] == epoch and f[8] == epoch - 1, (api, f)
        assert torch.equal(buf, want), api
    # stand-alone calls. Signal first, wait afterwards: a wait kernel spins on the GPU, and a flag that only a LATER launch on
    # this same GPU publishes may never come (streams can share a hardware queue) — the protocol always waits for flags of
    # OTHER GPUs or of work queued earlier (include/b200atmo.h); real cross-GPU blocking is covered by tests/test_multigpu_fused.py
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    ctx.peers_wait(flags, 0, 1, 3)                                               # already reached
    ctx.peers_signal([flags.data_ptr(), flags.data_ptr()], 3, 7, stream=s1.cuda_stream)
    ctx.peers_signal([flags.data_ptr()], 4, 9, stream=s1.cuda_stream)
    s1.synchronize()
    ctx.peers_wait(flags, 3, 2, 7, stream=s2.cuda_stream)                      # 7 >= 7 and 9 >= 7: epochs only need to be reached
    ctx.peers_wait(flags, 4, 1, 0xFFFFFFF0, stream=s2.cuda_stream)             # wrap-around: 9 is "after" 0xFFFFFFF0
    s2.synchronize()
    f = flags.cpu().numpy()
    assert f[3] == 7 and f[4] == 9 and f[0] == 3 and f[8] == 2
    assert ctx.peers_wait_timeouts() == 0
    from godot_atmosphere_shader_b200.context import B200AtmoError
    bad = sharding.peer_targets([buf.data_ptr()], use_tma=True)
    bad.sync.d_done_flags[0] = flags.data_ptr()
    bad.sync.n_done_flags = 1
    with pytest.raises(B200AtmoError):
        ctx.render_rays_peers(fr, d_od, d_dj, h * w, bad)


def test_heaviest_first_block_dispatch_keeps_every_pixel(cuda_ctx_factory):
    """Raymarched-cloud launches learn a heaviest-first dispatch order of their blocks from the previous launch with the same
    geometry on the same stream (block_order_kernel). The order is a permutation of the blocks, so the first (unordered) and
    every later (ordered) launch must write the same bits — frame API (full frame, row bands), tile-mapped ray batch, peer
    frames (interleaved), RGBA16F, several geometries alternating on one stream, two streams."""
    torch = _torch()
    ctx = cuda_ctx_factory()
    p = scenes.demo_params()
    _setup(ctx, p, 8, 48, abi.LIGHT_RAYMARCHED)
    ref_ctx = cuda_ctx_factory()
    _setup(ref_ctx, p, 8, 48, abi.LIGHT_RAYMARCHED)
    s2 = torch.cuda.Stream()
    for (w, h, camname) in ((333, 187, "A"), (256, 144, "C")):
        cam = _camera(camname, w, h, p)
        d_depth = torch.from_numpy(scenes.synth_depth(cam, p, w, h)).cuda()
        want = torch.empty((h, w, 4), dtype=torch.float32, device="cuda")
        wdisc = torch.empty((h, w), dtype=torch.uint8, device="cuda")
        ref_ctx.render_frame(cam, d_depth, w, h, want, wdisc)            # a fresh context's first launch: blockIdx order
        d_od = torch.empty((h * w, 4), dtype=torch.float32, device="cuda")
        d_dj = torch.empty((h * w, 4), dtype=torch.float32, device="cuda")
        fr = ctx.make_rays(cam, d_depth, w, h, d_od, d_dj)
        torch.cuda.synchronize()
        for rep in range(4):
            n0 = ctx.launch_count
            a = torch.full((h, w, 4), -7.0, dtype=torch.float32, device="cuda")
            ad = torch.full((h, w), 9, dtype=torch.uint8, device="cuda")
            ctx.render_frame(cam, d_depth, w, h, a, ad)
            if rep:
                assert ctx.launch_count == n0 + 2                         # the render kernel + the sort of its block costs (rep 0 also bakes the LUT)
            b = torch.full((h, w, 4), -7.0, dtype=torch.float32, device="cuda")
            ctx.render_rays(fr, d_od, d_dj, h * w, b, None, grid=(w, h))
            c16 = torch.full((h, w, 4), -7.0, dtype=torch.float16, device="cuda")
            ctx.render_frame(cam, d_depth, w, h, c16, None, rgba_format=abi.COLOR_RGBA16F)
            banded = torch.full((h, w, 4), -7.0, dtype=torch.float32, device="cuda")
            for (r0, r1) in ((0, 50), (50, h)):
                ctx.render_frame(cam, d_depth, w, h, banded, None, row_begin=r0, row_end=r1)
            peers = torch.full((h, w, 4), -7.0, dtype=torch.float32, device="cuda")
            t = sharding.peer_targets([peers.data_ptr()])
            for g in range(3):
                ctx.render_frame_peers_interleaved(cam, d_depth, w, h, t, g, 3)
            other = torch.full((h, w, 4), -7.0, dtype=torch.float32, device="cuda")
            s2.wait_stream(torch.cuda.current_stream())
            ctx.render_frame(cam, d_depth, w, h, other, None, stream=s2.cuda_stream)
            torch.cuda.synchronize()
            for name, got in (("frame", a), ("rays2d", b), ("bands", banded), ("peers", peers), ("stream2", other)):
                assert torch.equal(got, want), f"{name}, launch {rep}, {w}x{h}"
            assert torch.equal(ad, wdisc) and torch.equal(c16, want.to(torch.float16))
    # host paths (their own streams and staging)
    w, h = 333, 187
    cam = _camera("A", w, h, p)
    depth = scenes.synth_depth(cam, p, w, h)
    want_h = np.empty((h, w, 4), np.float32)
    ref_ctx.render_frame_host(cam, depth, w, h, want_h, None)
    for rep in range(3):
        got_h = np.full((h, w, 4), -7.0, np.float32)
        ctx.render_frame_host(cam, depth, w, h, got_h, None)
        assert np.array_equal(got_h.view(np.uint32), want_h.view(np.uint32))
    pinned = [torch.full((h, w, 4), -7.0, dtype=torch.float32).pin_memory() for _ in range(3)]
    hdep = torch.from_numpy(depth).pin_memory()
    for k in range(9):
        ctx.frame_wait(k % 3)
        ctx.render_frame_host_submit(cam, hdep, w, h, pinned[k % 3], None, slot=k % 3)
    for sl in range(3):
        ctx.frame_wait(sl)
    for b in pinned:
        assert np.array_equal(b.numpy().view(np.uint32), want_h.view(np.uint32))
