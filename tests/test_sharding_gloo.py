"""N>1 path on CPU: world_size-2 (and 3, uneven bands) `gloo` runs of the band partition + gather plumbing.
The per-band renderer stand-in is the oracle (the CUDA kernels need a GPU); the GPU variant of the same check is
tests/test_gpu_parity.py::test_full_size_properties (row-band invariance of the real kernels)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_band_partition():
    from godot_atmosphere_shader_b200.sharding import band, bands
    for h in (1, 7, 54, 1080, 2160):
        for world in (1, 2, 3, 4, 8):
            bs = bands(h, world)
            assert bs[0][0] == 0 and bs[-1][1] == h
            assert all(bs[i][1] == bs[i + 1][0] for i in range(world - 1))
            assert max(e - b for b, e in bs) - min(e - b for b, e in bs) <= 1
    assert band(1080, 3, 8) == (405, 540)
    with pytest.raises(ValueError):
        band(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, h, w, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from godot_atmosphere_shader_b200 import scenes
    from godot_atmosphere_shader_b200.sharding import band, gather_bands
    from oracle import pyoracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p = scenes.demo_params()
        cam = scenes.camera_a(w, h)
        depth = scenes.synth_depth(cam, p, w, h)
        tex = O.Textures(lut=O.bake_lut(p), blue_noise=scenes.blue_noise_tile())
        b, e = band(h, rank, world)
        rgba, _ = O.render_frame(p, O.variant(8), cam, tex, depth, w, h, row_begin=b, row_end=e)
        full = gather_bands(torch.from_numpy(rgba[b:e].copy()), h, w, rank, world)
        np.save(os.path.join(out_dir, f"full_{rank}.npy"), full.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,h", [(2, 36), (3, 37)])
def test_gloo_band_gather(tmp_path, world, h):
    import torch.multiprocessing as mp

    from godot_atmosphere_shader_b200 import scenes
    from oracle import pyoracle as O
    w = 64
    port = _free_port()
    mp.spawn(_worker, args=(world, port, h, w, str(tmp_path)), nprocs=world, join=True)
    p = scenes.demo_params()
    cam = scenes.camera_a(w, h)
    depth = scenes.synth_depth(cam, p, w, h)
    tex = O.Textures(lut=O.bake_lut(p), blue_noise=scenes.blue_noise_tile())
    want, _ = O.render_frame(p, O.variant(8), cam, tex, depth, w, h)
    for r in range(world):
        got = np.load(tmp_path / f"full_{r}.npy")
        assert np.array_equal(got, want), f"rank {r}: gathered frame differs from the single-process frame"


def test_peer_targets_validation():
    """Host side of the fused render + all-gather: the B200AtmoPeerTargets block built from symmetric-memory addresses."""
    from godot_atmosphere_shader_b200 import abi, sharding
    t = sharding.peer_targets([0x1000, 0x2000, 0x3000], multicast_ptr=0x9000, elem_offset=7)
    assert t.n_peers == 3 and [t.d_rgba_peers[r] for r in range(3)] == [0x1000, 0x2000, 0x3000]
    assert t.d_rgba_peers[3] is None and t.d_rgba_multicast == 0x9000 and t.elem_offset == 7
    assert sharding.peer_targets([0x1000]).d_rgba_multicast is None
    import pytest
    with pytest.raises(ValueError):
        sharding.peer_targets([])
    with pytest.raises(ValueError):
        sharding.peer_targets([0x1000] * (abi.MAX_PEERS + 1))
    with pytest.raises(ValueError):
        sharding.peer_targets([0x1000, 0])


def test_interleaved_row_tiles_partition():
    """The interleaved strong-scaling shard (b200atmo_render_frame_peers_interleaved): rank g owns the 8-row tiles g, g+G, ...;
    together the ranks own every row exactly once, also when the height is not a multiple of 8 or there are more ranks than tiles."""
    import numpy as np

    from godot_atmosphere_shader_b200.sharding import interleaved_rows
    for h in (1, 7, 8, 9, 54, 121, 1080, 2160):
        for world in (1, 2, 3, 8):
            owned = np.concatenate([interleaved_rows(h, r, world) for r in range(world)])
            assert sorted(owned.tolist()) == list(range(h)), (h, world)
            sizes = [len(interleaved_rows(h, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 8
    rows = interleaved_rows(2160, 3, 8)
    assert rows[0] == 24 and rows[8] == 24 + 64 and len(rows) == 8 * len(range(3, 270, 8))


def test_peer_targets_formats_and_hand_shake_block():
    """B200AtmoPeerTargets as sharding.peer_targets fills it: tile format, no hand-shake unless asked for."""
    from godot_atmosphere_shader_b200 import abi, sharding
    t = sharding.peer_targets([0x1000, 0x2000], rgba_format=abi.COLOR_RGBA16F, elem_offset=10, first_peer=5, use_tma=True)
    assert t.rgba_format == abi.COLOR_RGBA16F and t.first_peer == 1 and t.use_tma == 1 and t.elem_offset == 10
    y = t.sync
    assert (y.n_done_flags, y.n_consumed_flags, y.n_credit, y.n_wait) == (0, 0, 0, 0)
    assert sharding.peer_targets([0x1000]).rgba_format == abi.COLOR_RGBA32F
