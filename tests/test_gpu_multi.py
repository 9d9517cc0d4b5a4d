"""
Two (or more) GPUs, one process each under torchrun: the sharded evaluation's data path.  Skipped on a one-GPU box
(the gloo tests in test_api_cpu.py cover the partition and the collectives' ordering there).
"""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_sharded_evaluation_on_two_gpus():
    import torch
    count = torch.cuda.device_count()
    if count < 2:
        pytest.skip("needs two GPUs")
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "multi_gpu_worker.py")
    proc = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                           "--master-addr", "127.0.0.1", "--master-port", str(port), worker],
                          capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-4000:]
    assert "multi_gpu_worker: ok on 2 ranks" in proc.stdout
