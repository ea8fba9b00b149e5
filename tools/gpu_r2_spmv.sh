#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout 600 -p no:cacheprovider -k "sell or aprod or panels" > gpurun_out/r2_pytest_sell.log 2>&1; tail -3 gpurun_out/r2_pytest_sell.log
for V in 0 1 2; do for G in 1 2 4; do
  PROPACK_B200_SELL_VARIANT=$V PROPACK_B200_SPMV_COLBLOCKS=$G timeout 600 python tools/spmv_micro.py c5 c2 > gpurun_out/r2_spmv_jag_V${V}_G$G.json 2> gpurun_out/r2_spmv_jag_V${V}_G$G.err; echo "V=$V G=$G $(python -c "
import json,sys; d=json.load(open('gpurun_out/r2_spmv_jag_V${V}_G$G.json')); print({k:round(v['us'],1) for k,v in d.items() if isinstance(v,dict) and 'noflush' not in k})")"
done; done
