#!/bin/bash
# round 2, second pass (2 GPUs): SELL variants, full GPU suite, dist check, 2-GPU sharded sweep, 1-GPU bench c5
mkdir -p gpurun_out
for V in 1 3 4; do
  PROPACK_B200_SELL_VARIANT=$V timeout 600 python tools/spmv_micro.py c5 c2 c4 > gpurun_out/r2_spmv_var$V.json 2> gpurun_out/r2_spmv_var$V.err; echo "V=$V $(python -c "
import json,sys; d=json.load(open('gpurun_out/r2_spmv_var$V.json')); print({k:round(v['us'],1) for k,v in d.items() if isinstance(v,dict) and 'noflush' not in k})")"
done
PROPACK_B200_SELL_VARIANT=1 timeout 600 python tools/spmv_micro.py c4 > gpurun_out/r2_spmv_c4_v1.json 2>&1; cat gpurun_out/r2_spmv_c4_v1.json
timeout 1700 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -x > gpurun_out/r2_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_gpu.log; tail -8 gpurun_out/r2_pytest_gpu.log
timeout 900 python bench.py --workload c5 --steps 2 --warmup 1 > gpurun_out/r2_bench_c5_n1.json 2> gpurun_out/r2_bench_c5_n1.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_c5_n1.json')); print({k:d[k] for k in ('value','ms_per_step','phases_ms','spmv','reorth','gpu_launches','host_syncs_per_solve')}); print(d['e2e']); print(d['roofline']); print(d.get('cpu_baseline'))"; tail -3 gpurun_out/r2_bench_c5_n1.err
