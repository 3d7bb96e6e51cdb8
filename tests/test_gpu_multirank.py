"""B200 (`-m gpu`), >= 2 GPUs: the NCCL side of the path — `tools/check_multigpu.py` under torchrun with 2 ranks.
DDP (co_training_marco_train.py:107-114) + all-gathered [CLS] embeddings + the fused global InfoNCE: the loss equals the
single-process loss on the concatenated batch, world x the DDP-averaged gradient equals the single-process gradient,
ranks hold identical parameters after fused clip + AdamW steps, the sharded exact search equals the single index, and
the dropout + checkpointing configuration stays in lock-step.  Skipped on a single-GPU box (the driver's 1-GPU test run);
a kept log of a 2-GPU run is under profiles/."""
import os
import socket
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_nccl_parity():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run under `gpurun --gpus 2`)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), str(ROOT / "tools" / "check_multigpu.py")]
    env = dict(os.environ, NCCL_DEBUG="WARN")
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env, cwd=str(ROOT))
    log = proc.stdout + proc.stderr
    out_dir = ROOT / "gpurun_out"
    out_dir.mkdir(exist_ok=True)
    (out_dir / "multigpu_2rank.log").write_text(log)
    assert proc.returncode == 0, log[-4000:]
    assert log.count(": OK") >= 6, log[-4000:]
