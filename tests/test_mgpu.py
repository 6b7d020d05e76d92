"""Multi-GPU parity under the driver's `pytest -m gpu`: when the box has at least two GPUs this test torchruns
tests/mgpu_check.py (one process per GPU, NCCL for the rendezvous, NVLink peer memory for the exchange) and requires it
to pass; on a single-GPU box it is skipped.  The world_size-2 host logic is covered on CPU by tests/test_sharded.py."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.gpu
def test_sharded_buffer_multi_gpu_parity():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip(f"needs >= 2 GPUs, this box has {n}")
    world = 8 if n >= 8 else (4 if n >= 4 else 2)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", RLB_MGPU_LOG="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), str(ROOT / "tests" / "mgpu_check.py")]
    res = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    out = ROOT / "gpurun_out"
    if out.is_dir():
        (out / f"mgpu_check_w{world}.log").write_text(res.stdout + "\n--- stderr ---\n" + res.stderr[-20000:])
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-6000:]
    assert "mgpu_check ok" in res.stdout
