"""torchrun worker of tests/test_multigpu_fused.py (world >= 2, one process per GPU): the fused render + delivery
(b200atmo_render_*_peers over symmetric memory: NVLS multicast, P2P stores, TMA bulk stores; float4 and half4 tiles;
all-gather and deliver-to-root; contiguous and interleaved row shards) must leave on the consuming GPUs exactly the
bytes of render + ncclAllGather, and those bytes must agree with the ORACLE on a sample."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from godot_atmosphere_shader_b200 import abi, context, scenes, sharding  # noqa: E402
from oracle import pyoracle as O  # noqa: E402  (the checker)


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    w, h = 320, 184
    n = w * h
    p = scenes.demo_params()
    shape, cube, bn = scenes.shape_texture(16, seed=1), scenes.coverage_cubemap(32, seed=1), scenes.blue_noise_tile()
    ctx = context.AtmosphereContext(lr)
    ctx.set_params(p)
    ctx.set_variant(8, 32, abi.LIGHT_CHEAP)
    ctx.upload_blue_noise(bn)
    ctx.upload_shape3d(shape)
    ctx.upload_coverage_cube(cube)
    otex = O.Textures(lut=O.bake_lut(p), shape=shape, cube_faces=cube, blue_noise=bn)
    ovar = O.variant(8, 32, abi.LIGHT_CHEAP)
    side = torch.cuda.Stream()
    report = {}

    def oracle_check(got_tile, cam, what):
        """A strided sample of one fp32 tile against the oracle (tolerance 1e-4 rel + 2e-6)."""
        depth = scenes.synth_depth(cam, p, w, h)
        od, dj, fr = O.make_rays(p, cam, otex, depth, w, h)
        sel = np.arange(rank, n, 23)
        ref, _ = O.render_rays(p, ovar, fr, otex, od[sel], dj[sel])
        got = got_tile.reshape(-1, 4)[sel].astype(np.float64)
        gate = np.abs(got - ref) / (1e-4 * np.abs(ref) + 2e-6)
        assert gate.max() <= 1.0, f"rank {rank}: {what}: tile differs from the oracle (gate {gate.max():.2f})"

    # ---- (1) weak scaling: every rank renders its own tile (orbiting camera) ------------------------------------------
    cam = scenes.camera_a(w, h, orbit_deg=30.0 * rank)
    d_depth = torch.from_numpy(scenes.synth_depth(cam, p, w, h)).to(dev)
    d_od = torch.empty((n, 4), dtype=torch.float32, device=dev)
    d_dj = torch.empty((n, 4), dtype=torch.float32, device=dev)
    fr = ctx.make_rays(cam, d_depth, w, h, d_od, d_dj)
    mine = torch.empty((n, 4), dtype=torch.float32, device=dev)
    ctx.render_rays(fr, d_od, d_dj, n, mine, None)
    ref = torch.empty((world * n, 4), dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(ref, mine)
    torch.cuda.synchronize()
    oracle_check(ref.view(world, n, 4)[rank].cpu().numpy(), cam, "own tile")
    ref16 = ref.to(torch.float16)    # torch rounds to nearest-even
    for fmt, want in ((abi.COLOR_RGBA32F, ref), (abi.COLOR_RGBA16F, ref16)):
        for use_mc, use_tma, sync in ((True, False, "flags"), (False, False, "flags"), (False, False, "barrier"), (False, True, "barrier")):
            for stream in (None, side):
                tiles = sharding.SymmetricTiles(world, n, dev, use_multicast=use_mc, use_tma=use_tma, rgba_format=fmt, sync=sync)
                for t in tiles.tensors:
                    t.fill_(-1.0)
                torch.cuda.synchronize()
                dist.barrier()
                for _ in range(5):   # several frames through the double buffer (flow control kicks in from the third)
                    out = sharding.render_rays_and_gather_fused(ctx, fr, d_od, d_dj, n, tiles, stream=stream)
                torch.cuda.synchronize()
                assert torch.equal(out.view(world * n, 4), want), \
                    f"rank {rank}: fused tiles differ from render + all-gather (fmt={fmt}, multicast={use_mc}, tma={use_tma}, sync={sync})"
                report[f"gather_f{fmt}_mc{int(use_mc)}_tma{int(use_tma)}_{sync}"] = bool(tiles.multicast_ptr) if use_mc else True
                dist.barrier()
                del tiles
        # frame API + deliver-to-root: only the root's buffer is written
        root = world - 1
        tiles = sharding.SymmetricTiles(world, n, dev, rgba_format=fmt, root=root, sync="flags")
        for t in tiles.tensors:
            t.fill_(-1.0)
        torch.cuda.synchronize()
        dist.barrier()
        for _ in range(4):
            out = sharding.render_frame_tile_fused(ctx, cam, d_depth, w, h, tiles)
        torch.cuda.synchronize()
        dist.barrier()      # non-root ranks do not wait for anyone: let the root finish receiving before checking
        torch.cuda.synchronize()
        if rank == root:
            assert torch.equal(out.view(world * n, 4), want), f"root delivery differs (fmt={fmt})"
        else:
            assert bool((out == -1.0).all()), "deliver-to-root wrote into a non-root buffer"
        dist.barrier()
        del tiles

    # ---- (2) strong scaling: ONE frame sharded over the ranks, contiguous bands and interleaved 8-row tiles ------------
    cam1 = scenes.camera_a(w, h)
    d_depth1 = torch.from_numpy(scenes.synth_depth(cam1, p, w, h)).to(dev)
    full = torch.empty((h, w, 4), dtype=torch.float32, device=dev)
    ctx.render_frame(cam1, d_depth1, w, h, full, None)
    torch.cuda.synchronize()
    oracle_check(full.cpu().numpy(), cam1, "single-GPU frame")
    rows = sharding.interleaved_rows(h, rank, world)
    assert len(np.unique(np.concatenate([sharding.interleaved_rows(h, r, world) for r in range(world)]))) == h
    for fmt, want in ((abi.COLOR_RGBA32F, full), (abi.COLOR_RGBA16F, full.to(torch.float16))):
        for interleave in (False, True):
            for use_mc in (True, False):
                for root in (None, 0):
                    if use_mc and root is not None:
                        continue
                    ft = sharding.SymmetricTiles(1, n, dev, use_multicast=use_mc, rgba_format=fmt, root=root,
                                                 sync="flags" if interleave else "barrier")
                    for t in ft.tensors:
                        t.fill_(-1.0)
                    torch.cuda.synchronize()
                    dist.barrier()
                    for _ in range(4):
                        got = sharding.render_frame_sharded_fused(ctx, cam1, d_depth1, w, h, ft, interleave=interleave)
                    torch.cuda.synchronize()
                    dist.barrier()
                    torch.cuda.synchronize()
                    if root is None or rank == root:
                        assert torch.equal(got, want), \
                            f"rank {rank}: sharded fused frame differs from the single-GPU frame (fmt={fmt}, interleave={interleave}, mc={use_mc}, root={root})"
                    dist.barrier()
                    del ft
    assert len(rows) > 0
    assert ctx.peers_wait_timeouts() == 0, "a completion-flag wait timed out"
    ctx.close()
    if rank == 0:
        print("FUSED_GATHER_OK", world, report, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
