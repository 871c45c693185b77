"""multi-GPU parity (C-side NCCL exchange, sharded arc tally) against the oracle on the whole read set; every world size
the box offers among 2, 3, 4, 8. Skipped on a single-GPU box -- bench.py runs the same check before timing whenever it is
launched with WORLD_SIZE > 1, so the driver's scaling record carries the result too."""
import os
import subprocess
import sys
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_nccl_exchange_matches_oracle(world):
    import torch
    n = torch.cuda.device_count()
    if n < world:
        pytest.skip("needs >= %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29517 + world), os.path.join(ROOT, "tests", "run_multigpu_parity.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and r.stdout.count("MULTIGPU_PARITY OK") == 2, r.stdout[-3000:] + r.stderr[-3000:]
