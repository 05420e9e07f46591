"""Peer-window collectives (csrc/peer.cu: gradient all-reduce and halo exchange as kernels over NVLink peer memory) against
NCCL on two GPUs.  Needs >= 2 devices: skipped on the single-GPU test box, run with `gpurun --gpus 2`."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_peer_window_collectives_match_nccl():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(ROOT, "tests", "peer_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "peer_worker: ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
