"""Tensor-parallel parity on real GPUs (needs >= 2 devices; the driver's 1-GPU run skips it).
Launches tools/tp_check.py under torchrun: one process per GPU, NCCL all-reduce after Wo and w2."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("n", [2, 4, 8])
def test_tensor_parallel_matches_oracle(n):
    if _ngpu() < n:
        pytest.skip(f"needs {n} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + n), os.path.join(ROOT, "tools", "tp_check.py"), "tiny"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert out.returncode == 0 and lines, out.stdout[-2000:] + out.stderr[-2000:]
    res = json.loads(lines[-1])
    assert res["ok"], res


@pytest.mark.parametrize("path", ["chain", "engine"])
def test_late_rank_is_an_error_not_a_hang(path):
    """a tensor-parallel peer that never delivers its all-reduce words must become LNB_ETIMEOUT
    on the waiting rank within seconds (bounded in-kernel waits), in the kernel chain and in the persistent engine."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29671", os.path.join(ROOT, "tools", "tp_timeout_check.py"), path]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=180, cwd=ROOT)
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert lines, out.stdout[-2000:] + out.stderr[-2000:]
    res = json.loads(lines[-1])
    assert res["ok"] and res["code"] == -6 and res["seconds"] < 4.0, res
