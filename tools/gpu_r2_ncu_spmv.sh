#!/bin/bash
# ncu --set full of the SELL SpMV on the config-5 operands (both directions) + one CSR-kernel capture for comparison
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:spmv_sell_kernel -s 8 -c 2 -o gpurun_out/r2_prof_spmv_c5 -f python tools/spmv_micro.py c5 > gpurun_out/r2_ncu_spmv_c5.log 2>&1
tail -3 gpurun_out/r2_ncu_spmv_c5.log
PROPACK_B200_SPMV=csr ncu --set full --clock-control none -k regex:spmv_kernel -s 8 -c 1 -o gpurun_out/r2_prof_spmv_c5_csr -f python tools/spmv_micro.py c5 > gpurun_out/r2_ncu_spmv_c5_csr.log 2>&1
tail -3 gpurun_out/r2_ncu_spmv_c5_csr.log
ls -la gpurun_out/*.ncu-rep
