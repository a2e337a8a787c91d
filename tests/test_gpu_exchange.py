"""Multi-GPU test of the fused gradient exchange (C1 fused into K4): needs >= 2 GPUs on the box, otherwise skipped
(the single-GPU driver tier).  Spawns tools/exchange_check.py under torchrun: the multimem flush must equal the
NCCL all-reduce of the same per-rank gradients (1e-5), and an eager data-parallel run must follow the same
trajectory with either exchange."""
import os
import re
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_multicast_exchange_matches_nccl():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "exchange_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    out = res.stdout + res.stderr
    assert res.returncode == 0, out[-3000:]
    if "exchange: UNAVAILABLE" in out:
        pytest.skip("no NVLink multicast on this box; the NCCL path was exercised")
    errs = [float(x) for x in re.findall(r"max \|fused - nccl\| / max\|nccl\| = ([0-9.e+-]+)", out)]
    assert len(errs) == 3 and max(errs) < 1e-5, out[-3000:]
    m = re.search(r"eager, same RNG: params after 50 steps \|multicast - nccl\| max = ([0-9.e+-]+) \(of max ([0-9.e+-]+)\)", out)
    assert m and float(m.group(1)) < 0.05 * float(m.group(2)), out[-3000:]
