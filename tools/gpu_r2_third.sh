#!/bin/bash
# round 2, third pass (N GPUs): full GPU suite, dist check, sharded sweep
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 1700 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r2_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_gpu.log; tail -12 gpurun_out/r2_pytest_gpu.log
if [ "$N" -ge 2 ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/sharded_sweep.py c5-small "PHASES=1,PUSH=32" "PHASES=2,PUSH=32" "PHASES=2,PUSH=8" > gpurun_out/r2_sweep${N}_c5small.log 2>&1; grep '^{' gpurun_out/r2_sweep${N}_c5small.log; tail -3 gpurun_out/r2_sweep${N}_c5small.log | grep -v '^{'
fi
