#!/bin/bash
# multi-GPU parity check (run with gpurun --gpus N): fused peer-memory collectives (default) and the NCCL-only path
mkdir -p gpurun_out
N=${1:-2}
for FUSED in 1 2 0; do   # 1: fused collectives; 2: fused + column-grouped (chunk-pipelined) SpMV forced; 0: NCCL only
  GROUPS_ENV=""; [ "$FUSED" = "2" ] && GROUPS_ENV="PROPACK_B200_SPMV_GROUPS=$N"
  env $GROUPS_ENV PROPACK_B200_FUSED_COLLECTIVES=$([ "$FUSED" = "0" ] && echo 0 || echo 1) timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$FUSED tests/dist_check.py > gpurun_out/dist_check_${N}_fused$FUSED.log 2>&1; echo "rc=$?" >> gpurun_out/dist_check_${N}_fused$FUSED.log
  grep -E "dist_check|DIST_CHECK|rc=|rror" gpurun_out/dist_check_${N}_fused$FUSED.log | tail -12
done
