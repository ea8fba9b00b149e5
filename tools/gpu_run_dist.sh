#!/bin/bash
# multi-GPU parity check (run with gpurun --gpus N)
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/dist_gpus.txt
NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py > gpurun_out/dist_check_$N.log 2>&1; echo "rc=$?" >> gpurun_out/dist_check_$N.log
tail -30 gpurun_out/dist_check_$N.log
