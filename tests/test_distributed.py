"""Multi-process tests: world_size-2 gloo on CPU for the host-side sharding logic, NCCL on >= 2 GPUs for parity of the
sharded path (skipped on a single-GPU box)."""
import os
import subprocess
import sys

import pytest

from common import ROOT

WORKER = os.path.join(ROOT, "tests", "dist_worker.py")


def launch(nproc, args, port, env=None):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
           "--master-port", str(port), WORKER] + args
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=dict(os.environ, **(env or {})))


@pytest.mark.parametrize("world", [2, 3])
def test_shards_tile_the_grid_gloo(world):
    r = launch(world, ["shards"], 29517 + world)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "SHARDS_OK" in r.stdout, r.stdout + r.stderr[-2000:]


@pytest.mark.gpu
@pytest.mark.parametrize("peer", [1, 0], ids=["peer_memory", "nccl_only"])
@pytest.mark.parametrize("name", ["benzene_p631_fine", "h2o_sto3g", "ethane_p631_fine"])
def test_sharded_iteration_matches_golden_nccl(name, peer):
    """One process per GPU.  peer=1: the [J | XC] / F sum must really take the peer-memory kernels when the GPUs have a
    P2P route (the worker fails otherwise); peer=0: the ncclAllReduce variant."""
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    r = launch(2, ["parity", name], 29531, env={"DFTGRID_TEST_PEER": str(peer)})
    assert r.returncode == 0, r.stderr[-3000:]
    assert "PARITY_OK" in r.stdout, r.stdout + r.stderr[-3000:]
    assert ("peer_path=True" in r.stdout) == (peer == 1 and "peer_expected=True" in r.stdout), r.stdout
