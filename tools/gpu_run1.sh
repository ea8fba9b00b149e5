#!/bin/bash
# first GPU contact: smoke, parity tests, small + full bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt; lscpu | head -20 >> gpurun_out/nproc.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --workload c2-small --steps 2 --warmup 1 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "rc=$?" >> gpurun_out/bench_small.err
timeout 900 python bench.py --workload c2 --steps 2 --warmup 1 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "rc=$?" >> gpurun_out/bench_c2.err
tail -5 gpurun_out/smoke.log; tail -30 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_small.json; tail -3 gpurun_out/bench_small.err; cat gpurun_out/bench_c2.json; tail -3 gpurun_out/bench_c2.err
