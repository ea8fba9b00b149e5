#!/bin/bash
mkdir -p gpurun_out
timeout 600 tools/bin/spmv_lab 1000000 10 20 > gpurun_out/spmv_lab_1m.log 2>&1
timeout 600 tools/bin/spmv_lab 10000000 10 5 > gpurun_out/spmv_lab_10m.log 2>&1
timeout 600 tools/bin/spmv_lab 1000000 40 10 > gpurun_out/spmv_lab_1m_40.log 2>&1
cat gpurun_out/spmv_lab_1m.log gpurun_out/spmv_lab_10m.log gpurun_out/spmv_lab_1m_40.log
