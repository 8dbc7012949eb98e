"""2-GPU probe of the fused render + peer-store kernel under an 8-GPU-like load: every rank stores each result once
locally and SEVEN times into the other rank's buffer (same addresses, harmless), i.e. 232 MB outbound per 1080p tile over
one NVLink direction, as at N=8. Times the kernel (+ barrier) for several step counts to tell apart
  (a) remote writes overlapped with compute at a sustained bandwidth B:   t = max(t_compute, bytes / B)
  (b) compute and transfer in lock-step:                                   t = t_compute + bytes / 900 GB/s
run: torchrun --nproc-per-node 2 profiles/microbench/fused_gather_probe.py"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from godot_atmosphere_shader_b200 import context, scenes, sharding  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    assert world == 2
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    w, h = 1920, 1080
    n = w * h
    p = scenes.demo_params()
    cam = scenes.camera_b(w, h, p)
    ctx = context.AtmosphereContext(lr)
    ctx.set_params(p)
    ctx.upload_blue_noise(scenes.blue_noise_tile())
    stream = torch.cuda.current_stream().cuda_stream
    d_depth = torch.from_numpy(scenes.synth_depth(cam, p, w, h)).to(dev)
    d_od = torch.empty((n, 4), dtype=torch.float32, device=dev)
    d_dj = torch.empty((n, 4), dtype=torch.float32, device=dev)
    fr = ctx.make_rays(cam, d_depth, w, h, d_od, d_dj, stream=stream)
    local = torch.empty((n, 4), dtype=torch.float32, device=dev)
    tiles = sharding.SymmetricTiles(world, n, dev)
    other = tiles.buffer_ptrs[1 - rank]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timed(fn, steps=40):
        for _ in range(3):
            fn()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        dist.barrier(); torch.cuda.synchronize()
        for s, e in ev:
            flush.zero_()
            s.record(); fn(); e.record()
        torch.cuda.synchronize()
        t = torch.tensor([sum(s.elapsed_time(e) for s, e in ev) / steps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for steps in (32, 8, 2):
        ctx.set_variant(steps, 0, 0)
        t_local = timed(lambda: ctx.render_rays(fr, d_od, d_dj, n, local, None, stream=stream))
        row = {"steps": steps, "compute_ms": round(t_local, 4)}
        for copies in (1, 7):
            for tma in (False, True):
                tg = sharding.peer_targets([tiles.buffer_ptrs[rank]] + [other] * copies, elem_offset=rank * n, first_peer=1, use_tma=tma)

                def fused():
                    ctx.render_rays_peers(fr, d_od, d_dj, n, tg, stream=stream)
                    tiles.barrier()
                row[f"remote_x{copies}{'_tma' if tma else ''}_ms"] = round(timed(fused), 4)
        if rank == 0:
            mb = 7 * n * 16 / 1e6
            row["remote_x7_MB"] = round(mb, 1)
            row["model_a_GBps_if_overlapped"] = round(mb / 1e3 / (row["remote_x7_ms"] * 1e-3), 1)
            row["model_b_ms_compute_plus_900GBps"] = round(row["compute_ms"] + mb / 900e3 * 1e3, 4)
            print(row, flush=True)
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
