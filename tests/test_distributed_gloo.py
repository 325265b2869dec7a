"""N > 1 path on CPU: world_size 2 over gloo (the GPU box runs the same code over NCCL)."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("n", [11, 2])
def test_spmd_row_sharding_and_allgather_world2(n):
    port = 29500 + (os.getpid() + n) % 2000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(HERE, "_gloo_worker.py"), str(n)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-4000:]
    assert f"GLOO_OK {n} 2" in proc.stdout
