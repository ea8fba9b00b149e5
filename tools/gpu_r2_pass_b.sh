#!/bin/bash
# round 2, pass B (N GPUs of one box): row-sharded parity (tests/dist_check.py incl. the 1M-row config-5 pattern and the dense
# sharded cases) with the log kept, then the sharded bench of config 5 (and config 3, dense) on all GPUs.
mkdir -p gpurun_out
O=gpurun_out
N=$(nvidia-smi -L | wc -l)
{ nvidia-smi --query-gpu=index,name,memory.total --format=csv; nproc; free -g | head -2; nvidia-smi topo -m | head -12; } > $O/r02_box_n$N.txt 2>&1
DIST_CHECK_LARGE_ROWS=${DIST_CHECK_LARGE_ROWS:-1000000} timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py > $O/r02_dist_check_$N.log 2>&1; echo "rc=$?" >> $O/r02_dist_check_$N.log
grep -E "dist_check|DIST_CHECK|rc=|rror" $O/r02_dist_check_$N.log | tail -16
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 2 > $O/r02_bench_c5_n$N.json 2> $O/r02_bench_c5_n$N.err; echo "bench c5 N=$N rc=$?"
python - <<P
import json
try:
    d=json.load(open('gpurun_out/r02_bench_c5_n$N.json'))
    print({k:d.get(k) for k in ('value','ms_per_step','lanczos_steps_per_solve','converged','info','sigma_1','gpu_launches','host_syncs_per_solve','collectives_total')})
    print(d['e2e']); print(d['phases_ms_profiled_solve'], d['profiled_solve_ms'])
except Exception as e: print('bench parse failed', e)
P
tail -3 $O/r02_bench_c5_n$N.err
if [ "${RUN_C3:-1}" = "1" ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --workload c3 --steps 1 --warmup 1 > $O/r02_bench_c3_n$N.json 2> $O/r02_bench_c3_n$N.err; echo "bench c3 N=$N rc=$?"
python - <<P
import json
try:
    d=json.load(open('gpurun_out/r02_bench_c3_n$N.json'))
    print({k:d.get(k) for k in ('value','ms_per_step','lanczos_steps_per_solve','converged','info','sigma_1','gpu_launches')})
    print(d['e2e']); print(d['phases_ms_profiled_solve'], d['profiled_solve_ms'])
except Exception as e: print('bench parse failed', e)
P
tail -3 $O/r02_bench_c3_n$N.err
fi
