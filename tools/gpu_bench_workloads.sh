#!/bin/bash
# single-GPU bench lines for the restarted (IRL) workloads
mkdir -p gpurun_out
for WL in "$@"; do
  timeout 1500 python bench.py --workload $WL --steps 1 --warmup 1 > gpurun_out/bench_${WL}_n1.json 2> gpurun_out/bench_${WL}_n1.err; echo "rc=$?" >> gpurun_out/bench_${WL}_n1.err
  cat gpurun_out/bench_${WL}_n1.json; tail -4 gpurun_out/bench_${WL}_n1.err
done
