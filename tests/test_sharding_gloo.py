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
