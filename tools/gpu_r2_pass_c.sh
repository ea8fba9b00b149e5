#!/bin/bash
# round 2, pass C (1 GPU): FP64 tensor-pipe ceiling, gemm_tall variants, ncu --set full of the l = 300 GEMV pair and the
# 101-column restart GEMM at config-5 size, SpMV on a banded matrix (gathers not the bound), the whole GPU suite, final bench.
mkdir -p gpurun_out
O=gpurun_out
tools/bin/dmma_peak > $O/r02_dmma_peak.txt 2>&1; cat $O/r02_dmma_peak.txt
timeout 300 python tools/fp64_peak.py > $O/r02_fp64_peak.json 2> $O/r02_fp64_peak.err; cat $O/r02_fp64_peak.json
PROF_NO_SPMV=1 PROF_LS=300 timeout 300 python tools/prof_target.py c5 > $O/r02_gemm_mt1.txt 2>&1; tail -1 $O/r02_gemm_mt1.txt
PROPACK_B200_GEMM_MT2=1 PROF_NO_SPMV=1 PROF_LS=300 timeout 300 python tools/prof_target.py c5 > $O/r02_gemm_mt2.txt 2>&1; tail -1 $O/r02_gemm_mt2.txt
PROPACK_B200_GEMM_MT2=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 600 -p no:cacheprovider -k "gemm or ritzvec" > $O/r02_pytest_gemm_mt2.log 2>&1; tail -2 $O/r02_pytest_gemm_mt2.log
PROF_NO_SPMV=1 PROF_REPS=1 PROF_LS=300 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemv_t_tma_kernel|gemv_n_kernel|gemm_tall' -c 16 -o $O/r02_prof_gemv_gemm -f python tools/prof_target.py c5 > $O/r02_ncu_gemv_gemm.log 2>&1
tail -2 $O/r02_ncu_gemv_gemm.log
timeout 600 python tools/spmv_micro.py banded:4000000 > $O/r02_spmv_banded.json 2> $O/r02_spmv_banded.err; cat $O/r02_spmv_banded.json
timeout 1700 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > $O/r02_pytest_gpu_full.log 2>&1; echo "rc=$?" >> $O/r02_pytest_gpu_full.log; tail -6 $O/r02_pytest_gpu_full.log
timeout 1500 python bench.py --steps 3 --warmup 3 > $O/r02_bench_c5_n1_final.json 2> $O/r02_bench_c5_n1_final.err; echo "bench c5 rc=$?"
python - <<'P'
import json
try:
    d=json.load(open('gpurun_out/r02_bench_c5_n1_final.json'))
    print({k:d.get(k) for k in ('value','ms_per_step','lanczos_steps_per_solve','phases_ms','gpu_launches','host_syncs_per_solve')})
    print(d['e2e']); print({k:d['roofline'][k] for k in ('kernel','achieved','frac','share_of_solve','traffic')})
except Exception as e: print('bench c5 parse failed', e)
P
timeout 600 python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline > $O/r02_bench_c2_n1_final.json 2> $O/r02_bench_c2_n1_final.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_c2_n1_final.json')); print(d['ms_per_step'], d['e2e'])"
