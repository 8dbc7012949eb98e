"""Fused render + all-gather over NVLink peer memory on >= 2 GPUs (skipped on single-GPU boxes): launches
tests/multigpu_fused_worker.py under torchrun, one process per GPU."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.gpu
def test_fused_render_allgather_matches_nccl():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    world = 2 if n < 4 else 4
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(HERE, "multigpu_fused_worker.py")],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "FUSED_GATHER_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
