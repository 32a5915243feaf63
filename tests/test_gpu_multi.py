"""Multi-GPU parity (-m gpu, needs >= 2 GPUs of one NVSwitch box; skipped otherwise): the density maps assembled by the
kernel's own multimem.st stores into a symmetric buffer (`dist.MulticastMaps.fill`, fc_grid_log_prob_multicast) must be
bit-identical on every rank to the NCCL all-gather of the same shards (`dist.sharded_grid_log_prob`), for the
agents-sharded, the points-sharded and the config-2-sized case.  One process per GPU under torch.distributed.run."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multicast_gather_is_bit_identical_to_nccl(world):
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29600 + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "check_multicast.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:])
    print(r.stderr[-2000:])
    assert r.returncode == 0 and "CHECK OK" in r.stdout, "multicast-assembled maps differ from the NCCL all-gather"
    assert r.stdout.count("multicast == nccl all-gather: True") == 3
