#!/bin/bash
# round 2, pass F (N GPUs): SELL kernel shape x phases on the row-sharded config 5 (short panel rows are latency-bound)
# (record of what produced profiles/r02_sweep8_c5_*.jsonl: the CHAINS / GRAPH / DEPTH / FUSED keys belong to experiments that were
#  measured and then removed from the library -- tools/sharded_sweep.py ignores keys it no longer knows)
mkdir -p gpurun_out
O=gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/sharded_sweep.py ${SWEEP_WL:-c5} \
  "MODE=sm,PHASES=4,PUSH=64" "MODE=sm,PHASES=4,PUSH=64,VARIANT=2" "MODE=sm,PHASES=4,PUSH=64,VARIANT=0" "MODE=sm,PHASES=2,PUSH=64" \
  "MODE=sm,PHASES=2,PUSH=64,VARIANT=2" "MODE=sm,PHASES=4,PUSH=48" "MODE=ce,PHASES=4,CHAINS=1,GRAPH=0" "MODE=sm,PHASES=8,PUSH=64,VARIANT=2" > $O/r02_sweep${N}_c5_shape.log 2>&1
grep '^{' $O/r02_sweep${N}_c5_shape.log | python -c "
import sys, json
for ln in sys.stdin:
    d = json.loads(ln); print(d['setting'], [round(x, 1) for x in d['ms']], d['converged'], d['info'], d['phases_ms'], d.get('spmv_isolated'))"
grep -v '^{' $O/r02_sweep${N}_c5_shape.log | grep -E "RuntimeError" | head -3; true
