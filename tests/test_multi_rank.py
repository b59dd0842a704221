"""N>1 coverage: world_size-2 gloo processes on CPU for the partition logic; torchrun on 2 GPUs for the real thing."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "mg_worker.py")


def _torchrun(nproc, args, port, timeout):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), WORKER] + args
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout)


@pytest.mark.parametrize("world,dim", [(2, 2), (3, 2), (2, 3)])
def test_partition_host_gloo(lib, world, dim):
    r = _torchrun(world, ["host", str(dim)], 29610 + world + 10 * dim, 600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "host partition OK" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("dim,levels", [(2, ()), (3, ()), (2, ("4", "10")), (3, ("3", "6"))])
def test_two_gpus_match_oracle(gpu, dim, levels):
    """the advection loop on 2 GPUs (fused cross-GPU wavefront, halo values stored into the peer by the producing threads) against
    the oracle: meshes identical on both ranks at every step, leaf values gathered from their owners bit-identical.  The larger level
    ranges have phases with many chunks per rank: they caught a missing barrier between the zero fill of the transfer buffers and
    the peers' halo stores that the small meshes never exposed."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = _torchrun(2, ["gpu", str(dim)] + list(levels), 29650 + dim + (10 if levels else 0), 900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-4000:]
    assert "multi-GPU parity OK" in r.stdout
