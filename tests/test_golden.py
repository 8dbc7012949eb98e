"""Golden vectors (tests/golden/atmo_golden_v1.npz, generator: tests/golden/make_golden.py).

CPU: the oracle must reproduce them bit for bit (fp32) — guards the checker itself.
GPU: the CUDA path (through the C-ABI) must match them within the parity tolerance."""
import ctypes as C
import hashlib
import os

import numpy as np
import pytest

from godot_atmosphere_shader_b200 import abi
from oracle import pyoracle as O
from tests import helpers as Hh
from tests.golden.make_golden import VARIANTS, params_for

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "atmo_golden_v1.npz"))


def _structs():
    p = abi.B200AtmoParams.from_buffer_copy(G["params"].tobytes())
    fr = abi.B200AtmoFrame.from_buffer_copy(G["frame"].tobytes())
    return p, fr


def test_golden_inputs_are_what_the_generator_makes():
    from tests.golden.make_golden import inputs
    p, shape, cube, od, dj, fr = inputs()
    assert bytes(p) == G["params"].tobytes() and bytes(fr) == G["frame"].tobytes()
    assert np.array_equal(od, G["origin_depth"]) and np.array_equal(dj, G["dir_jitter"])
    assert np.array_equal(shape, G["shape"]) and np.array_equal(cube, G["cube"])


def test_oracle_lut_and_cube_layout_match_golden():
    p, _ = _structs()
    lut = O.bake_lut(p)
    assert hashlib.sha256(lut.tobytes()).digest() == G["lut_sha256"].tobytes()
    assert np.array_equal(lut[::8, ::8], G["lut_sub"])
    assert hashlib.sha256(O.cube_build_padded(G["cube"]).tobytes()).digest() == G["cube_padded_sha256"].tobytes()


@pytest.mark.parametrize("name", list(VARIANTS))
def test_oracle_reproduces_golden(name):
    p, fr = _structs()
    v = VARIANTS[name]
    q = params_for(p, v)
    tex = O.Textures(lut=O.bake_lut(q), shape=G["shape"], cube_faces=G["cube"])
    rgba, disc = O.render_rays(q, O.variant(v[1], v[2], v[3], v[0]), fr, tex, G["origin_depth"], G["dir_jitter"])
    assert np.array_equal(disc, G[f"discard_{name}"])
    assert np.array_equal(rgba.view(np.uint32), G[f"rgba_{name}"].view(np.uint32))
    # the fp64 twin bounds the fp32 oracle's own rounding error (documented, not a pass/fail on accuracy)
    assert np.isfinite(G[f"rgba64_{name}"]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(VARIANTS))
def test_cuda_matches_golden(cuda_ctx_factory, name):
    import torch
    p, fr = _structs()
    v = VARIANTS[name]
    q = params_for(p, v)
    ctx = cuda_ctx_factory()
    ctx.set_params(q)
    ctx.set_variant(v[1], v[2], v[3], v[0])
    ctx.upload_shape3d(G["shape"])
    ctx.upload_coverage_cube(G["cube"])
    n = G["origin_depth"].shape[0]
    d_od = torch.from_numpy(G["origin_depth"]).cuda()
    d_dj = torch.from_numpy(G["dir_jitter"]).cuda()
    d_rgba = torch.empty((n, 4), dtype=torch.float32, device="cuda")
    d_disc = torch.empty((n,), dtype=torch.uint8, device="cuda")
    ctx.render_rays(fr, d_od, d_dj, n, d_rgba, d_disc)
    torch.cuda.synchronize()
    assert np.array_equal(d_disc.cpu().numpy(), G[f"discard_{name}"])
    Hh.assert_rgba_close(d_rgba.cpu().numpy(), G[f"rgba_{name}"], what=f"golden/{name}")
    if name == "no_clouds":
        lut = ctx.download_lut()
        assert hashlib.sha256(lut.tobytes()).digest() == G["lut_sha256"].tobytes()
