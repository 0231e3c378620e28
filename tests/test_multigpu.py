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


def test_two_contexts_on_two_devices_in_one_process(bun000):
    """function attributes (the opt-in shared-memory limit of the sweeps, the cooperative level kernel) are PER DEVICE: a second
    context on another device of the same process must launch the same kernels (process-static caches broke this)."""
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import hgmm_b200
    from hgmm_b200 import hgmm as H
    J = 800
    mu0 = bun000[np.random.default_rng(1).choice(len(bun000), J, replace=False)]
    cov0 = np.tile(np.eye(3, dtype=np.float32) * 1e-4, (J, 1, 1))
    w0 = np.full(J, 1 / J, np.float32)
    init = bun000[H.reference_init_indices(3)]
    out = []
    for dev in (0, 1):
        eng = hgmm_b200.Engine(dev)
        eng.set_points(bun000)
        f = eng.fit_flat(mu0, cov0, w0, cov_type="full", max_iter=5)           # em_flat7: 232 kB of opt-in shared memory
        t = eng.fit_tree(init, 3, ls=20.0, ld=1e-4, sig2=4e-4, ll_mode="estep", want_current=False)      # cooperative level kernel
        out.append((f, t))
        eng.close()
    assert np.array_equal(out[0][0]["means"], out[1][0]["means"]) and np.array_equal(out[0][0]["covs"], out[1][0]["covs"])
    # the tree build uses fp64 atomics (order varies run to run): the root level agrees to rounding, below it a massless node can
    # flip blank / alive on the last bit, so the deeper levels are compared through the mixing weights
    assert np.allclose(out[0][1]["mu"][:8], out[1][1]["mu"][:8], rtol=0, atol=1e-6)
    assert np.abs(out[0][1]["pi"] - out[1][1]["pi"]).sum() < 1e-3

