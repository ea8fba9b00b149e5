#!/bin/bash
# GPU regression: full gpu test-suite, kernel micro-benchmarks, C2 bench line
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python tools/prof_target.py c2 > gpurun_out/micro.log 2>&1
timeout 900 python bench.py --workload c2 --steps 3 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "rc=$?" >> gpurun_out/bench_c2.err
tail -4 gpurun_out/pytest_gpu.log; cat gpurun_out/micro.log; python -c "
import json; d=json.load(open('gpurun_out/bench_c2.json')); print({k:d[k] for k in ('value','ms_per_step','phases_ms','spmv','isolated_kernels_gbs')}); print(d['e2e']); print(d['roofline'])"; tail -3 gpurun_out/bench_c2.err
