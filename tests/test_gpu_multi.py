"""Two-rank parity on a box with >= 2 GPUs (skipped on the single-GPU test box): tools/parity_multi.py under torchrun."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs (NCCL refuses two ranks on one device)")
@pytest.mark.parametrize("cfg,scale", [("C5", 1), ("C2", 4)])
def test_two_rank_partition_is_bit_exact(cfg, scale):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "parity_multi.py"), cfg, str(scale)]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert '"head_bit_exact": true' in r.stdout and '"gap_solve_bit_exact": true' in r.stdout


@pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs (NCCL refuses two ranks on one device)")
@pytest.mark.parametrize("deal", ["roundrobin", "last"])
def test_two_rank_amr_hierarchy_is_bit_exact(deal):
    """3-level hierarchy, refined-level boxes dealt out to the ranks: copy plans cross ranks (pack / ncclSend-Recv / unpack)"""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29534", os.path.join(ROOT, "tools", "parity_multi_amr.py"), "3", deal]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert '"head_bit_exact_per_level": [true, true, true]' in r.stdout
