#!/bin/bash
# round 2, pass A (1 GPU): kernel parity, bench on configs 5 and 2, ncu launch list of the bench, ncu --set full of the hot kernels
# at working size (config 5 operands, l = 300 reorthogonalisation, 101-column restart GEMM), then the whole GPU suite.
mkdir -p gpurun_out
O=gpurun_out
{ nvidia-smi --query-gpu=index,name,memory.total --format=csv; nproc; free -g | head -2; } > $O/r02_box.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout 600 -p no:cacheprovider > $O/r02_pytest_kernels.log 2>&1; echo "rc=$?" >> $O/r02_pytest_kernels.log
tail -4 $O/r02_pytest_kernels.log
timeout 1500 python bench.py > $O/r02_bench_c5_n1.json 2> $O/r02_bench_c5_n1.err; echo "bench c5 rc=$?"
python - <<'P'
import json
try:
    d=json.load(open('gpurun_out/r02_bench_c5_n1.json'))
    print({k:d.get(k) for k in ('value','ms_per_step','lanczos_steps_per_solve','phases_ms','gpu_launches','host_syncs_per_solve','isolated_kernels_gbs')})
    print(d['e2e']); print({k:d['roofline'][k] for k in ('kernel','achieved','frac','share_of_solve')}); print(d['reorth']); print(d['spmv']); print(d.get('cpu_baseline',{}).get('value'))
except Exception as e: print('bench c5 parse failed', e)
P
tail -3 $O/r02_bench_c5_n1.err
timeout 900 python bench.py --workload c2 --steps 5 --warmup 3 > $O/r02_bench_c2_n1.json 2> $O/r02_bench_c2_n1.err; echo "bench c2 rc=$?"
python - <<'P'
import json
try:
    d=json.load(open('gpurun_out/r02_bench_c2_n1.json'))
    print({k:d.get(k) for k in ('value','ms_per_step','lanczos_steps_per_solve','phases_ms','gpu_launches','host_syncs_per_solve','isolated_kernels_gbs')})
    print(d['e2e'])
except Exception as e: print('bench c2 parse failed', e)
P
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 4000 --csv --log-file $O/r02_launches_bench_c5.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $O/r02_ncu_bench.log 2>&1
python tools/ncu_launches.py $O/r02_launches_bench_c5.csv > $O/r02_launches_bench_c5_summary.txt 2>&1; cat $O/r02_launches_bench_c5_summary.txt
PROF_REPS=1 PROF_LS=64,300 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'spmv_sell_kernel|gemv_t_kernel|gemv_t_tma|gemv_n_kernel|gemm_tall' -c 16 -o $O/r02_prof_c5 -f python tools/prof_target.py c5 > $O/r02_ncu_full.log 2>&1
tail -3 $O/r02_ncu_full.log; ls -la $O/*.ncu-rep
timeout 1700 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider --deselect tests/test_gpu_kernels.py > $O/r02_pytest_gpu.log 2>&1; echo "rc=$?" >> $O/r02_pytest_gpu.log; tail -8 $O/r02_pytest_gpu.log
