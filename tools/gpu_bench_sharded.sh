#!/bin/bash
# multi-GPU bench line (strong scaling of C2) ; run with gpurun --gpus N
mkdir -p gpurun_out
N=${1:-2}; WL=${2:-c2}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 2 --warmup 1 --workload $WL > gpurun_out/bench_${WL}_n$N.json 2> gpurun_out/bench_${WL}_n$N.err; echo "rc=$?" >> gpurun_out/bench_${WL}_n$N.err
cat gpurun_out/bench_${WL}_n$N.json; tail -5 gpurun_out/bench_${WL}_n$N.err
