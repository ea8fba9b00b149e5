#!/bin/bash
# ncu evidence: launch list of the bench command + --set full capture of the hot kernels
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 1500 --csv --log-file gpurun_out/launches_r01c.csv python bench.py --workload c2 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
PROF_REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'spmv_kernel|gemv_t_kernel|gemv_n_kernel|gemm_real_kernel|gemm_tall' -c 24 -o gpurun_out/prof_r01c -f python tools/prof_target.py c2 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_bench.log; tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out | tail -5
