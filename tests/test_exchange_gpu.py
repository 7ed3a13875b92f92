"""The gradient exchange fused with the optimizer over NVLink multicast (csrc/exchange.cu) needs two GPUs of one
NVSwitch node: the checks live in tools/check_exchange.py / tools/check_overlap.py (torchrun programs) and this test
launches them when the box has at least two devices (a single-GPU box skips; `gpurun --gpus 2` runs them)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(script, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", script)]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs of one node")
@pytest.mark.parametrize("script,port", [("check_exchange.py", 29541), ("check_overlap.py", 29542)])
def test_two_rank_exchange_program(script, port):
    p = _torchrun(script, port)
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert p.returncode == 0 and lines and '"ok": true' in lines[-1], (p.stdout[-2000:], p.stderr[-2000:])
