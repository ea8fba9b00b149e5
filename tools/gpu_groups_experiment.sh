#!/bin/bash
# 8-GPU experiment: number of column groups of the chunk-pipelined sharded SpMV on C5
mkdir -p gpurun_out
N=${1:-8}
for G in 1 2; do
  PROPACK_B200_SPMV_GROUPS=$G timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$G bench.py --gpus $N --steps 2 --warmup 1 --workload c5 --no-e2e > gpurun_out/bench_c5_n${N}_G$G.json 2> gpurun_out/bench_c5_n${N}_G$G.err; echo "rc=$?" >> gpurun_out/bench_c5_n${N}_G$G.err
  python - "$G" <<'PY'
import json,sys
G=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/bench_c5_n8_G{G}.json').read().strip().splitlines()[-1])
    print('G',G,'ms',d['ms_per_step'],'steps/s',d['value'],'conv',d['converged'],'phases',d['phases_ms_profiled_solve'])
except Exception as e:
    print('G',G,'ERR',e); print(open(f'gpurun_out/bench_c5_n8_G{G}.err').read()[-1500:])
PY
done
