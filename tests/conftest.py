import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # a fresh checkout has no built artefacts (they are git-ignored): build them once (nvcc cross-compiles without a GPU)
    pkg = os.path.join(ROOT, "godot_atmosphere_shader_b200")
    if not (os.path.exists(os.path.join(pkg, "libb200atmo.so")) and os.path.exists(os.path.join(pkg, "libb200atmo_node.so"))):
        import subprocess
        subprocess.check_call(["bash", os.path.join(pkg, "csrc", "build.sh")], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


@pytest.fixture(scope="session")
def cuda_ctx_factory():
    """Factory of AtmosphereContext objects on cuda:0; fails loudly when the native library is missing."""
    import torch

    from godot_atmosphere_shader_b200 import context

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    context.lib()
    made = []

    def make():
        c = context.AtmosphereContext(0)
        made.append(c)
        return c

    yield make
    for c in made:
        c.close()
