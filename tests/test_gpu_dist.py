"""GPU (>= 2 devices): the row-sharded multi-GPU path under torchrun, checked against dense SVD and the oracle.

Runs on every visible GPU (up to 8); on a single-GPU lease the tests skip.  `__graft_entry__.smoke()` runs the same
check (small cases only) whenever it sees more than one device, and the logs of the 2 / 4 / 8-GPU runs are kept under
profiles/.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_dist_check(world, env, port=29517, timeout=1500):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dist_check.py")]
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout, env=dict(os.environ, **env))


@pytest.mark.gpu
@pytest.mark.parametrize("env", [{}, {"PROPACK_B200_SPMV_PHASES": "1", "DIST_CHECK_LARGE_ROWS": "0"},
                                 {"PROPACK_B200_FUSED_COLLECTIVES": "0", "DIST_CHECK_LARGE_ROWS": "0"},
                                 {"PROPACK_B200_SPMV": "csr", "DIST_CHECK_LARGE_ROWS": "0"},
                                 {"PROPACK_B200_PUSH": "ce", "DIST_CHECK_LARGE_ROWS": "0"}],
                         ids=["fused-phased-sell", "fused-one-phase", "nccl-only", "csr-kernel", "copy-engine-transport"])
def test_row_sharded_drivers_all_gpus(env):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run under `gpurun --gpus N`)")
    world = min(8, n)
    p = run_dist_check(world, env)
    sys.stdout.write(p.stdout[-6000:]); sys.stderr.write(p.stderr[-4000:])
    assert p.returncode == 0 and "DIST_CHECK_OK" in p.stdout
