"""GPU, needs >= 2 devices (skipped on a 1-GPU box): points sharded over ranks, NCCL all-reduce of the sufficient
statistics inside libhgmm -- results must equal the single-GPU fit of the whole cloud."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_fits_match_single_gpu(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + world), os.path.join(HERE, "multigpu_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    sys.stdout.write(out.stdout[-3000:])
    sys.stderr.write(out.stderr[-3000:])
    assert out.returncode == 0
    assert "MULTIGPU %d" % world in out.stdout
