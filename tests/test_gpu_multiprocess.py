"""The real multi-GPU path: one process per GPU, NCCL send/recv of the halo messages (skipped with fewer than 2 GPUs;
the single-GPU emulation of the same library path is tests/test_gpu_parity.py::test_slabs_bitwise_equal_to_whole_box)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _ngpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_slabs_over_nccl_bitwise_equal_to_whole_box(world):
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29600 + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(HERE, "mp_slab_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MP_SLAB_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
