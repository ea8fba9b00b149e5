#!/bin/bash
# multi-GPU parity check (run with gpurun --gpus N): fused peer-memory collectives with the push kernel (default), with the
# copy-engine transport, and the NCCL-only path
mkdir -p gpurun_out
N=${1:-2}
for MODE in default ce nccl; do
  ENVV=""
  [ "$MODE" = "ce" ] && ENVV="PROPACK_B200_PUSH=ce DIST_CHECK_LARGE_ROWS=0"
  [ "$MODE" = "nccl" ] && ENVV="PROPACK_B200_FUSED_COLLECTIVES=0 DIST_CHECK_LARGE_ROWS=0"
  env $ENVV timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py > gpurun_out/dist_check_${N}_$MODE.log 2>&1; echo "rc=$?" >> gpurun_out/dist_check_${N}_$MODE.log
  grep -E "dist_check|DIST_CHECK|rc=|rror" gpurun_out/dist_check_${N}_$MODE.log | tail -12
done
