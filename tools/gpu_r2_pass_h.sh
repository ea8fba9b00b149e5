#!/bin/bash
# round 2, pass H (N GPUs, final code): row-sharded parity with the default transport (incl. the 1M-row config-5 pattern) and with the
# copy-engine transport (small cases), then the sharded bench of config 5.
mkdir -p gpurun_out
O=gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py > $O/r02_dist_check_${N}_final.log 2>&1; echo "rc=$?" >> $O/r02_dist_check_${N}_final.log
grep -E "DIST_CHECK|rc=|rror|FAIL" $O/r02_dist_check_${N}_final.log | tail -6
PROPACK_B200_PUSH=ce DIST_CHECK_LARGE_ROWS=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tests/dist_check.py > $O/r02_dist_check_${N}_final_ce.log 2>&1; echo "rc=$?" >> $O/r02_dist_check_${N}_final_ce.log
grep -E "DIST_CHECK|rc=|rror|FAIL" $O/r02_dist_check_${N}_final_ce.log | tail -6
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 3 --warmup 2 > $O/r02_bench_c5_n${N}_final.json 2> $O/r02_bench_c5_n${N}_final.err; echo "bench c5 N=$N rc=$?"
python - <<P
import json
try:
    d=json.load(open('gpurun_out/r02_bench_c5_n${N}_final.json'))
    print({k:d.get(k) for k in ('value','ms_per_step','lanczos_steps_per_solve','converged','info','sigma_1','gpu_launches','host_syncs_per_solve','collectives_total')})
    print(d['e2e']); print(d['phases_ms_profiled_solve'], d['profiled_solve_ms']); print(d['roofline'])
except Exception as e: print('bench parse failed', e)
P
tail -3 $O/r02_bench_c5_n${N}_final.err
