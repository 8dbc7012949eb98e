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


# ------------------------------------------------------------------------------------------------
# fused render + all-gather over NVLink / NVSwitch peer memory (b200atmo_render_*_peers)
# ------------------------------------------------------------------------------------------------
def peer_targets(buffer_ptrs, multicast_ptr=None, elem_offset: int = 0, first_peer: int = 0, use_tma: bool = False,
                 rgba_format: int = 0):
    """B200AtmoPeerTargets from the device addresses of one symmetric buffer as mapped in this process."""
    from .abi import MAX_PEERS, B200AtmoPeerTargets

    ptrs = [int(p) for p in buffer_ptrs]
    if not (1 <= len(ptrs) <= MAX_PEERS):
        raise ValueError(f"1..{MAX_PEERS} peers supported, got {len(ptrs)}")
    if any(p == 0 for p in ptrs):
        raise ValueError("NULL peer buffer")
    t = B200AtmoPeerTargets()
    for r, p in enumerate(ptrs):
        t.d_rgba_peers[r] = p
    t.n_peers = len(ptrs)
    t.d_rgba_multicast = int(multicast_ptr) if multicast_ptr else None
    t.elem_offset = int(elem_offset)
    t.first_peer = int(first_peer) % len(ptrs)   # (rank + 1) % world staggers the ranks' destinations
    t.use_tma = 1 if use_tma else 0
    t.rgba_format = int(rgba_format)
    return t


class SymmetricTiles:
    """One [slots, rays_per_slot, 4] fp32 buffer per rank in symmetric memory (torch.distributed._symmetric_memory: CUDA
    VMM allocations exchanged between the ranks' processes, plus the NVLS multicast mapping when the fabric has one).
    Rank r's render kernels store slot r of EVERY rank's buffer directly (`targets(rank)`), so after `barrier()` each GPU
    holds all tiles: the render is the all-gather. PyTorch is the plumbing (allocation, rendezvous, barrier) only."""

    def __init__(self, slots: int, rays_per_slot: int, device, group=None, use_multicast: bool = False, stagger: bool = True,
                 use_tma: bool = False):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        self.slots, self.rays_per_slot = int(slots), int(rays_per_slot)
        self.stagger = bool(stagger)
        self.use_tma = bool(use_tma)
        group = group if group is not None else dist.group.WORLD
        self.tensor = symm_mem.empty((self.slots, self.rays_per_slot, 4), dtype=torch.float32, device=device)
        self.handle = symm_mem.rendezvous(self.tensor, group.group_name)
        self.world = self.handle.world_size
        self.rank = self.handle.rank
        mc = getattr(self.handle, "multicast_ptr", 0) if use_multicast else 0
        self.multicast_ptr = int(mc) if mc else None
        self.buffer_ptrs = [int(p) for p in self.handle.buffer_ptrs]

    def targets(self, slot: int):
        return peer_targets(self.buffer_ptrs, self.multicast_ptr, elem_offset=int(slot) * self.rays_per_slot,
                            first_peer=(self.rank + 1) % self.world if self.stagger else 0, use_tma=self.use_tma)

    def barrier(self):
        """Stream-ordered inter-rank barrier on the current CUDA stream: after it, every rank's stores have landed here."""
        self.handle.barrier()


def render_rays_and_gather_fused(ctx, frame, d_origin_depth, d_dir_jitter, n_rays: int, tiles: "SymmetricTiles", stream=None):
    """Weak-scaling delivery without a collective pass: this rank's rays are rendered straight into slot `rank` of every
    rank's tile buffer; returns after the inter-rank barrier is queued (results are complete in stream order)."""
    ctx.render_rays_peers(frame, d_origin_depth, d_dir_jitter, n_rays, tiles.targets(tiles.rank), stream=stream)
    tiles.barrier()
    return tiles.tensor


def render_frame_sharded_fused(ctx, cam, d_depth, width: int, height: int, tiles: "SymmetricTiles", stream=None):
    """Screen-tile shard of ONE frame: rank g renders rows [g*H/G, (g+1)*H/G) into every rank's full-frame buffer
    (`tiles` built with slots=1, rays_per_slot=width*height); returns the [height, width, 4] view after the barrier."""
    b, e = band(height, tiles.rank, tiles.world)
    ctx.render_frame_peers(cam, d_depth, width, height, tiles.targets(0), row_begin=b, row_end=e, stream=stream)
    tiles.barrier()
    return tiles.tensor.view(height, width, 4)
