"""The sharded table over real NCCL / P2P with one process per GPU (needs >= 2 GPUs; skipped otherwise).

Runs ``tests/dist_check.py --parity-only`` under torchrun for each exchange path: device-flag handshake
(default), host-ordered P2P inboxes, staged NCCL send/recv.  Every rank compares its gathered result with
the oracle.
"""

from __future__ import annotations

import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus() -> int:
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("mode", ["device_flags", "p2p_host_ordered", "staged_nccl"])
def test_sharded_over_nccl(mode):
    if _n_gpus() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ)
    env["OBP_B200_P2P"] = "0" if mode == "staged_nccl" else "1"
    env["OBP_B200_DEVICE_FLAGS"] = "1" if mode == "device_flags" else "0"
    port = {"device_flags": 29611, "p2p_host_ordered": 29612, "staged_nccl": 29613}[mode]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dist_check.py"), "--parity-only"]
    res = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "parity vs oracle OK on 2 ranks" in res.stdout
