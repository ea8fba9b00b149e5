#!/bin/bash
# gemm_tall DMMA evidence (ncu --set full, 2 launches) + example driver test
mkdir -p gpurun_out
PROF_REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_tall_kernel' -c 2 -o gpurun_out/prof_r01_gemm -f python tools/prof_target.py c2 > gpurun_out/ncu_gemm.log 2>&1
tail -2 gpurun_out/ncu_gemm.log
timeout 600 python -m pytest tests/test_gpu_drivers.py -m gpu -q -k "example_driver or dense_synth" -p no:cacheprovider 2>&1 | tail -3
timeout 300 python examples/example.py tests/golden/illc1850.rra --k 10 --kmax 100 --compare tests/golden/Sigma_illc1850.ascii > gpurun_out/example_illc1850.log 2>&1; cat gpurun_out/example_illc1850.log
