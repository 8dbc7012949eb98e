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


def interleaved_rows(height: int, rank: int, world: int):
    """Rows of rank `rank` under the interleaved shard: the 8-row tiles t with t % world == rank (numpy index array)."""
    import numpy as np
    tiles = np.arange(rank, (height + 7) // 8, world)
    rows = (tiles[:, None] * 8 + np.arange(8)[None, :]).reshape(-1)
    return rows[rows < height]


class SymmetricTiles:
    """`depth` x [slots, rays_per_slot, 4] tile buffers per rank in symmetric memory (torch.distributed._symmetric_memory:
    CUDA VMM allocations exchanged between the ranks' processes, plus the NVLS multicast mapping when the fabric has one).
    Rank r's render kernels store slot r straight into the buffers of the ranks that consume it (`targets(rank)`), so
    once the frame is complete those GPUs hold the tiles: the render is the delivery. PyTorch is the plumbing (allocation,
    rendezvous) only.

    rgba_format : abi.COLOR_RGBA32F (float4 tiles) or abi.COLOR_RGBA16F (half4 tiles — Godot's own colour-target format,
                  half the NVLink bytes; each channel is the fp32 result rounded to nearest-even).
    root        : None = all-gather (every rank receives every tile); r = deliver-to-root (only rank r's buffer is
                  written: 1/world of the all-gather's fabric traffic).
    depth       : number of buffers used round-robin, one per frame.
    sync        : "barrier" (default) — one symmetric-memory barrier per frame behind the render kernel (torch's kernel, ~14 us
                  at 2 GPUs); with depth >= 2 it also rules out overwriting a buffer a slower rank still reads.
                  "flags" — the hand-shake is carried out by the render kernel itself (B200AtmoPeerSync): its last block
                  publishes a completion flag into the consumers' flag arrays and, on a consumer, waits for the other
                  producers' flags, so the kernel ends when the frame is complete; its first block publishes "consumed" for
                  the previous frame and every block waits for that credit before it overwrites a buffer. No barrier, no
                  extra launch — but every block needs a fence behind its peer stores, which holds its CTA slot for one
                  NVLink round trip (~2 us): measured 4-24 us SLOWER per frame than the barrier, whose
                  kernel-boundary drain is free. Kept for loosely coupled producers / consumers that cannot meet in a barrier.
    All calls of one frame (render, hand-shake, the consumer's reads) must be queued on the same CUDA stream."""

    DONE_BASE = 0           # flags[DONE_BASE + r]     = last epoch whose pixels producer r has delivered here
    CONSUMED_BASE = 8       # flags[CONSUMED_BASE + r] = last epoch consumer r has finished reading (published to the producers)

    def __init__(self, slots: int, rays_per_slot: int, device, group=None, use_multicast: bool = False, stagger: bool = True,
                 use_tma: bool = False, rgba_format: int = 0, root=None, depth: int = 2, sync: str = "barrier"):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        from .abi import COLOR_RGBA16F

        self.slots, self.rays_per_slot = int(slots), int(rays_per_slot)
        self.stagger = bool(stagger)
        self.use_tma = bool(use_tma)
        self.rgba_format = int(rgba_format)
        self.root = None if root is None else int(root)
        self.depth = max(1, int(depth))
        self.sync = "barrier" if self.use_tma else sync      # the TMA flavour has no fused completion signal
        assert self.sync in ("flags", "barrier")
        group = group if group is not None else dist.group.WORLD
        dtype = torch.float16 if self.rgba_format == COLOR_RGBA16F else torch.float32
        self.tensors, self.handles = [], []
        for _ in range(self.depth):
            t = symm_mem.empty((self.slots, self.rays_per_slot, 4), dtype=dtype, device=device)
            self.tensors.append(t)
            self.handles.append(symm_mem.rendezvous(t, group.group_name))
        self.world = self.handles[0].world_size
        self.rank = self.handles[0].rank
        self.cur = 0
        self.epoch = 0
        self._mc = [int(getattr(hd, "multicast_ptr", 0) or 0) if (use_multicast and self.root is None) else 0 for hd in self.handles]
        self._ptrs = [[int(q) for q in hd.buffer_ptrs] for hd in self.handles]
        self.flags = symm_mem.empty((16,), dtype=torch.int32, device=device)
        self.flags.zero_()
        self._flag_handle = symm_mem.rendezvous(self.flags, group.group_name)
        self._flag_ptrs = [int(q) for q in self._flag_handle.buffer_ptrs]
        torch.cuda.synchronize(device)
        self._flag_handle.barrier()          # every rank's flags are zero before anyone signals
        torch.cuda.synchronize(device)

    # the buffer of the current frame
    @property
    def tensor(self):
        return self.tensors[self.cur]

    @property
    def handle(self):
        return self.handles[self.cur]

    @property
    def multicast_ptr(self):
        return self._mc[self.cur] or None

    @property
    def buffer_ptrs(self):
        return self._ptrs[self.cur]

    @property
    def consumers(self):
        return list(range(self.world)) if self.root is None else [self.root]

    @property
    def is_consumer(self):
        return self.root is None or self.rank == self.root

    def advance(self):
        """Switch to the next buffer (once per frame, before rendering it)."""
        self.cur = (self.cur + 1) % self.depth

    def targets(self, slot: int):
        ptrs = self.buffer_ptrs if self.root is None else [self.buffer_ptrs[self.root]]
        first = (self.rank + 1) % self.world if (self.stagger and self.root is None) else 0
        t = peer_targets(ptrs, self.multicast_ptr, elem_offset=int(slot) * self.rays_per_slot, first_peer=first,
                         use_tma=self.use_tma, rgba_format=self.rgba_format)
        if self.sync == "flags":
            y, e = t.sync, self.epoch
            cons = self.consumers
            for k, r in enumerate(cons):                      # producer: tell every consumer when my pixels have landed
                y.d_done_flags[k] = self._flag_ptrs[r]
            y.n_done_flags = len(cons)
            y.done_slot = self.DONE_BASE + self.rank
            y.epoch = e & 0xFFFFFFFF
            if e > self.depth:                                # producer: the buffer I overwrite held frame e - depth
                y.d_credit_flags = self._flag_ptrs[self.rank]
                y.credit_first_slot = self.CONSUMED_BASE + cons[0]
                y.n_credit = len(cons)
                y.credit_epoch = (e - self.depth) & 0xFFFFFFFF
            if self.is_consumer:
                if e > 1:                                     # consumer: everything queued before this kernel has read frame e - 1
                    for k in range(self.world):
                        y.d_consumed_flags[k] = self._flag_ptrs[k]
                    y.n_consumed_flags = self.world
                    y.consumed_slot = self.CONSUMED_BASE + self.rank
                    y.consumed_epoch = (e - 1) & 0xFFFFFFFF
                y.d_wait_flags = self._flag_ptrs[self.rank]   # consumer: the kernel ends when every producer has delivered
                y.wait_first_slot = self.DONE_BASE
                y.n_wait = self.world
        return t

    def bytes_sent_per_frame(self, pixels_rendered: int) -> int:
        """NVLink egress of this rank for `pixels_rendered` pixels (stores into its own buffer do not touch the fabric)."""
        from .abi import COLOR_RGBA16F
        px = 8 if self.rgba_format == COLOR_RGBA16F else 16
        if self.root is None:
            return (self.world - 1) * pixels_rendered * px
        return 0 if self.rank == self.root else pixels_rendered * px

    def begin_frame(self, ctx=None, stream=None):
        """Next buffer, next epoch (once per frame, before `targets`)."""
        self.advance()
        self.epoch += 1

    def end_frame(self, ctx=None, stream=None):
        """sync="flags": nothing to queue — the render kernel carried the whole hand-shake (B200AtmoPeerSync) and, on a
        consumer, only completes when every producer's pixels are here. sync="barrier": one inter-rank barrier."""
        if self.sync != "flags":
            self.barrier(stream)

    def barrier(self, stream=None):
        """Stream-ordered inter-rank barrier: after it, every rank's stores have landed here. `stream` = the
        torch.cuda.Stream the render was queued on (None = the current stream); the barrier kernel is queued there."""
        import torch
        if stream is None:
            self.handle.barrier()
        else:
            with torch.cuda.stream(stream):
                self.handle.barrier()


def _raw_stream(stream):
    """torch.cuda.Stream | None -> the cudaStream_t the C-ABI takes (None = torch's current stream)."""
    import torch
    return (torch.cuda.current_stream() if stream is None else stream).cuda_stream


def render_rays_and_gather_fused(ctx, frame, d_origin_depth, d_dir_jitter, n_rays: int, tiles: "SymmetricTiles", stream=None,
                                 grid=None):
    """Weak-scaling delivery without a collective pass: this rank's rays are rendered straight into slot `rank` of the
    consuming ranks' tile buffers; returns after the inter-rank barrier is queued (results are complete in stream order).
    `stream`: torch.cuda.Stream or None (current); `grid=(w, h)` selects the tile-mapped ray kernel."""
    if grid is not None:
        raise NotImplementedError("peer stores are implemented for the linear ray mapping and the frame API")
    tiles.begin_frame(ctx, stream)
    ctx.render_rays_peers(frame, d_origin_depth, d_dir_jitter, n_rays, tiles.targets(tiles.rank), stream=_raw_stream(stream))
    tiles.end_frame(ctx, stream)
    return tiles.tensor


def render_frame_tile_fused(ctx, cam, d_depth, width: int, height: int, tiles: "SymmetricTiles", stream=None):
    """Weak scaling through the FRAME API: this rank's whole width x height tile (its own camera / depth buffer) goes to
    slot `rank` of the consuming ranks' buffers."""
    tiles.begin_frame(ctx, stream)
    ctx.render_frame_peers(cam, d_depth, width, height, tiles.targets(tiles.rank), stream=_raw_stream(stream))
    tiles.end_frame(ctx, stream)
    return tiles.tensor


def render_frame_sharded_fused(ctx, cam, d_depth, width: int, height: int, tiles: "SymmetricTiles", stream=None,
                               interleave: bool = False):
    """Screen-tile shard of ONE frame (strong scaling) into the consuming ranks' full-frame buffers (`tiles` built with
    slots=1, rays_per_slot=width*height): rank g renders rows [g*H/G, (g+1)*H/G), or — `interleave` — the 8-row tiles
    g, g+G, g+2G, ... (balanced when the work is not uniform over the frame). Returns the [height, width, 4] view after
    the barrier is queued."""
    tiles.begin_frame(ctx, stream)
    if interleave:
        ctx.render_frame_peers_interleaved(cam, d_depth, width, height, tiles.targets(0), tiles.rank, tiles.world,
                                           stream=_raw_stream(stream))
    else:
        b, e = band(height, tiles.rank, tiles.world)
        ctx.render_frame_peers(cam, d_depth, width, height, tiles.targets(0), row_begin=b, row_end=e, stream=_raw_stream(stream))
    tiles.end_frame(ctx, stream)
    return tiles.tensor.view(height, width, 4)
