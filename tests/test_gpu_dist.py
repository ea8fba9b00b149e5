"""GPU (>= 2 devices): the row-sharded multi-GPU path under torchrun, checked against dense SVD and the oracle."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("env", [{}, {"PROPACK_B200_SPMV_GROUPS": "2"}, {"PROPACK_B200_FUSED_COLLECTIVES": "0"}],
                         ids=["fused", "fused-chunked-spmv", "nccl-only"])
def test_row_sharded_drivers_two_gpus(env):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run under `gpurun --gpus 2`)")
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tests", "dist_check.py")]
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900, env=dict(os.environ, **env))
    sys.stdout.write(p.stdout[-4000:]); sys.stderr.write(p.stderr[-4000:])
    assert p.returncode == 0 and "DIST_CHECK_OK" in p.stdout
