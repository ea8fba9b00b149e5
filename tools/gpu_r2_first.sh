#!/bin/bash
# round 2, first GPU pass: kernel + driver parity, SELL vs CSR SpMV, 2-GPU dist check
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2_gpus.txt; nproc >> gpurun_out/r2_gpus.txt
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/r2_pytest_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_kernels.log
tail -15 gpurun_out/r2_pytest_kernels.log
timeout 600 python tools/spmv_micro.py c2 c5 c4 > gpurun_out/r2_spmv_sell.json 2> gpurun_out/r2_spmv_sell.err; cat gpurun_out/r2_spmv_sell.json; tail -3 gpurun_out/r2_spmv_sell.err
PROPACK_B200_SPMV=csr timeout 600 python tools/spmv_micro.py c2 c5 c4 > gpurun_out/r2_spmv_csr.json 2> gpurun_out/r2_spmv_csr.err; cat gpurun_out/r2_spmv_csr.json; tail -3 gpurun_out/r2_spmv_csr.err
timeout 1500 python -m pytest tests/test_gpu_drivers.py tests/test_gpu_at_size.py -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r2_pytest_drivers.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_drivers.log
tail -25 gpurun_out/r2_pytest_drivers.log
N=$(nvidia-smi -L | wc -l)
if [ "$N" -ge 2 ]; then
  DIST_CHECK_LARGE_ROWS=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py > gpurun_out/r2_dist_check_${N}.log 2>&1; echo "rc=$?" >> gpurun_out/r2_dist_check_${N}.log
  grep -E "dist_check|DIST_CHECK|rc=|rror" gpurun_out/r2_dist_check_${N}.log | tail -12
fi
