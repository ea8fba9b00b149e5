#!/bin/bash
# what the driver runs at round end: gpu tests, smoke, default bench (both arms)
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
( time timeout 900 python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "rc=$?" >> gpurun_out/bench_default.err
( time timeout 900 python bench.py --impl reference ) > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "rc=$?" >> gpurun_out/bench_reference.err
tail -12 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; cat gpurun_out/bench_default.json; tail -5 gpurun_out/bench_default.err; cat gpurun_out/bench_reference.json; tail -5 gpurun_out/bench_reference.err
