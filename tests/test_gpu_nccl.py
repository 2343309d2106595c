"""The NCCL path (pnp_comm.cu) against the GLOBAL oracle: spawns one process per GPU with torch.distributed.run when at
least two GPUs are visible (tests/nccl_worker.py does the checking), skips otherwise.  Run on a multi-GPU box with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_nccl.py -m gpu`; its log is kept under profiles/."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world,levels", [(2, 2), (4, 1)])
def test_nccl_ranks_match_global_oracle(world, levels):
    if _ngpus() < world:
        pytest.skip("needs %d GPUs" % world)
    port = 29500 + (os.getpid() % 2000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "nccl_worker.py"), str(levels)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=420, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("NCCL_WORKER_OK") == world, out.stdout[-3000:]
