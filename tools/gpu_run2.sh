#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python tools/prof_target.py c2 > gpurun_out/micro.log 2>&1
timeout 900 python bench.py --workload c2 --steps 3 --warmup 2 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "rc=$?" >> gpurun_out/bench_c2.err
# per-launch device times of the bench command (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 1500 --csv --log-file gpurun_out/launches_r01.csv python bench.py --workload c2 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
# full capture of the hot kernels
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'spmv_kernel|gemv_t_kernel|gemv_n_kernel|gemm_tall_kernel' -c 14 -o gpurun_out/prof_r01 -f python tools/prof_target.py c2 > gpurun_out/ncu_full.log 2>&1
tail -25 gpurun_out/pytest_gpu.log; cat gpurun_out/micro.log; cat gpurun_out/bench_c2.json; tail -3 gpurun_out/bench_c2.err; tail -3 gpurun_out/ncu_bench.log; tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out
