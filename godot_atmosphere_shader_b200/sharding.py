"""Screen-tile sharding of one frame across the GPUs of a box.

Rays are independent, so the frame is cut into contiguous row bands, one per rank; every rank owns a full
context (LUT + textures are < 15 MB: replicated, the LUT is baked redundantly) and renders only its band through
`b200atmo_render_frame(row_begin, row_end)`. There is no data-path collective. The only exchange is the optional
delivery of the finished RGBA bands (`gather_bands`, one all-gather over NVLink/NVSwitch via torch.distributed —
plumbing, not product); bands of unequal height are padded to the tallest band for the collective.
"""


def band(height: int, rank: int, world: int):
    """Rows [begin, end) of rank `rank`: GPU g gets rows [g*H/G, (g+1)*H/G)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return (height * rank) // world, (height * (rank + 1)) // world


def bands(height: int, world: int):
    return [band(height, r, world) for r in range(world)]


def gather_bands(local_band, height: int, width: int, rank: int, world: int, channels: int = 4):
    """All-gather the per-rank bands into the full [height, width, channels] image on every rank.

    `local_band`: torch tensor [rows_of_this_rank, width, channels] on the rank's device (CUDA for NCCL, CPU for gloo).
    """
    import torch
    import torch.distributed as dist

    bs = bands(height, world)
    rows_max = max(e - b for b, e in bs)
    b, e = bs[rank]
    assert tuple(local_band.shape) == (e - b, width, channels), (tuple(local_band.shape), (e - b, width, channels))
    if world == 1:
        return local_band
    send = local_band
    if e - b != rows_max:
        send = torch.zeros((rows_max, width, channels), dtype=local_band.dtype, device=local_band.device)
        send[: e - b] = local_band
    # concatenated (not stacked) output layout: accepted by both the NCCL and the gloo backends
    recv = torch.empty((world * rows_max, width, channels), dtype=local_band.dtype, device=local_band.device)
    dist.all_gather_into_tensor(recv, send.contiguous())
    recv = recv.view(world, rows_max, width, channels)
    if all(e2 - b2 == rows_max for b2, e2 in bs):
        return recv.reshape(height, width, channels)
    return torch.cat([recv[r, : bs[r][1] - bs[r][0]] for r in range(world)], dim=0)


def render_frame_sharded(ctx, cam, d_depth, width: int, height: int, rank: int, world: int, d_rgba_full, stream=None,
                         gather: bool = True):
    """Render this rank's band into `d_rgba_full` ([height, width, 4] on this rank's GPU) and optionally gather."""
    b, e = band(height, rank, world)
    ctx.render_frame(cam, d_depth, width, height, d_rgba_full, None, row_begin=b, row_end=e, stream=stream)
    if gather and world > 1:
        full = gather_bands(d_rgba_full[b:e], height, width, rank, world)
        d_rgba_full.copy_(full.reshape(d_rgba_full.shape))
    return d_rgba_full


def render_tile_and_gather_overlapped(ctx, cam, d_depth, width: int, height: int, d_mine, d_chunks, rank: int, world: int,
                                      chunks: int = 4, stream=None):
    """Weak-scaling delivery (BASELINE config[4] shape): every rank renders its OWN width x height tile (`d_mine`,
    [height, width, 4]) and all ranks receive all tiles.

    The tile is rendered in `chunks` equal row chunks on the current stream; right after a chunk is launched its
    all-gather is issued asynchronously (torch.distributed runs NCCL on its own stream and makes it wait for the work
    already queued on the current stream), so the NVLink transfer of chunk k overlaps the rendering of chunk k+1.
    `d_chunks` is the receive buffer in CHUNK-MAJOR layout [chunks, world, height/chunks, width, 4]: each collective
    writes one contiguous slab (no staging copies); tile r, rows of chunk k = d_chunks[k, r]. height % chunks == 0.
    """
    import torch.distributed as dist

    assert height % chunks == 0, "chunked gather needs equal row chunks"
    rows = height // chunks
    works = []
    for k in range(chunks):
        r0, r1 = k * rows, (k + 1) * rows
        ctx.render_frame(cam, d_depth, width, height, d_mine, None, row_begin=r0, row_end=r1, stream=stream)
        if world > 1:
            works.append(dist.all_gather_into_tensor(d_chunks[k].view(world * rows, width, 4), d_mine[r0:r1], async_op=True))
    for wk in works:
        wk.wait()
    return d_chunks
