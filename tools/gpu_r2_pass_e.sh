#!/bin/bash
# round 2, pass E (N GPUs): push-kernel grid / depth sweep on config 5 + isolated local panel time (tools/sharded_sweep.py)
# (record of what produced profiles/r02_sweep8_c5_*.jsonl: the CHAINS / GRAPH / DEPTH / FUSED keys belong to experiments that were
#  measured and then removed from the library -- tools/sharded_sweep.py ignores keys it no longer knows)
mkdir -p gpurun_out
O=gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/sharded_sweep.py ${SWEEP_WL:-c5} \
  "MODE=sm,PHASES=4,PUSH=64" "MODE=sm,PHASES=4,PUSH=96" "MODE=sm,PHASES=4,PUSH=128" "MODE=sm,PHASES=4,PUSH=148" "MODE=sm,PHASES=4,PUSH=64,DEPTH=8" \
  "MODE=sm,PHASES=4,PUSH=128,DEPTH=8" "MODE=sm,PHASES=2,PUSH=128" "MODE=sm,PHASES=8,PUSH=128" "MODE=sm,PHASES=1,PUSH=128" > $O/r02_sweep${N}_c5_push.log 2>&1
grep '^{' $O/r02_sweep${N}_c5_push.log | python -c "
import sys, json
for ln in sys.stdin:
    d = json.loads(ln); print(d['setting'], [round(x, 1) for x in d['ms']], d['converged'], d['info'], d['phases_ms'], d.get('spmv_isolated'))"
tail -3 $O/r02_sweep${N}_c5_push.log | grep -v '^{'; true
