#!/bin/bash
# 8-GPU: parity check (fused collectives) + C5 bench line
mkdir -p gpurun_out
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tests/dist_check.py > gpurun_out/dist_check_${N}_fused1.log 2>&1; echo "rc=$?" >> gpurun_out/dist_check_${N}_fused1.log
grep -E "dist_check\]|DIST_CHECK|rc=|rror" gpurun_out/dist_check_${N}_fused1.log | tail -12
bash tools/gpu_bench_sharded.sh $N ${2:-c5}
