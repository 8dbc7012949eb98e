"""torchrun worker of tests/test_multigpu_fused.py (world >= 2, one process per GPU): the fused render + all-gather
(b200atmo_render_*_peers over symmetric memory, NVLS multicast and plain P2P stores) must leave on every GPU exactly the
bytes of render + ncclAllGather."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from godot_atmosphere_shader_b200 import abi, context, scenes, sharding  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    w, h = 320, 184
    p = scenes.demo_params()
    ctx = context.AtmosphereContext(lr)
    ctx.set_params(p)
    ctx.set_variant(8, 32, abi.LIGHT_CHEAP)
    ctx.upload_blue_noise(scenes.blue_noise_tile())
    ctx.upload_shape3d(scenes.shape_texture(16, seed=1))
    ctx.upload_coverage_cube(scenes.coverage_cubemap(32, seed=1))
    stream = torch.cuda.current_stream().cuda_stream
    report = {}
    for use_mc, use_tma in ((True, False), (False, False), (False, True)):
        # (1) weak scaling: every rank renders its own tile (orbiting camera), all ranks receive all tiles
        cam = scenes.camera_a(w, h, orbit_deg=30.0 * rank)
        d_depth = torch.from_numpy(scenes.synth_depth(cam, p, w, h)).to(dev)
        n = w * h
        d_od = torch.empty((n, 4), dtype=torch.float32, device=dev)
        d_dj = torch.empty((n, 4), dtype=torch.float32, device=dev)
        fr = ctx.make_rays(cam, d_depth, w, h, d_od, d_dj, stream=stream)
        mine = torch.empty((n, 4), dtype=torch.float32, device=dev)
        ctx.render_rays(fr, d_od, d_dj, n, mine, None, stream=stream)
        ref = torch.empty((world * n, 4), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(ref, mine)
        tiles = sharding.SymmetricTiles(world, n, dev, use_multicast=use_mc, use_tma=use_tma)
        tiles.tensor.fill_(-1.0)
        tiles.barrier()
        out = sharding.render_rays_and_gather_fused(ctx, fr, d_od, d_dj, n, tiles, stream=stream)
        torch.cuda.synchronize()
        assert torch.equal(out.view(world * n, 4), ref), f"rank {rank}: fused tiles differ from render + all-gather (multicast={use_mc}, tma={use_tma})"
        report[f"tiles_mc{int(use_mc)}_tma{int(use_tma)}"] = bool(tiles.multicast_ptr)
        # (2) strong scaling: ONE frame, rank g renders its row band into every rank's full frame
        cam1 = scenes.camera_a(w, h)
        d_depth1 = torch.from_numpy(scenes.synth_depth(cam1, p, w, h)).to(dev)
        full = torch.empty((h, w, 4), dtype=torch.float32, device=dev)
        ctx.render_frame(cam1, d_depth1, w, h, full, None, stream=stream)
        frame_tiles = sharding.SymmetricTiles(1, n, dev, use_multicast=use_mc)
        frame_tiles.tensor.fill_(-1.0)
        frame_tiles.barrier()
        got = sharding.render_frame_sharded_fused(ctx, cam1, d_depth1, w, h, frame_tiles, stream=stream)
        torch.cuda.synchronize()
        assert torch.equal(got, full), f"rank {rank}: band-sharded fused frame differs from the single-GPU frame (multicast={use_mc}, tma={use_tma})"
        dist.barrier()
    ctx.close()
    if rank == 0:
        print("FUSED_GATHER_OK", world, report, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
