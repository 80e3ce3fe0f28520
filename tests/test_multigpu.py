"""2-GPU NCCL parity of the sharded PPO path (skipped unless >= 2 GPUs are visible; run with `gpurun --gpus 2`)."""
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_sharded_ppo_matches_single_gpu():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", str(ROOT / "tools" / "multigpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert "MULTIGPU_CHECK PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MULTIGPU_GRAPH_CHECK PASS" in r.stdout, r.stdout[-3000:]
    assert "MULTIGPU_PEER_CHECK PASS" in r.stdout, r.stdout[-3000:]   # one-shot NVLink peer reduction == ncclAllReduce path
    assert "MULTIGPU_DETERMINISM_CHECK PASS" in r.stdout, r.stdout[-3000:]
