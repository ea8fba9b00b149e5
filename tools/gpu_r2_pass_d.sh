#!/bin/bash
# round 2, pass D (N GPUs): parity of the copy-engine all-gather transport (tests/dist_check.py), then one torchrun launch that
# solves config 5 row-sharded under several transport / phase settings (tools/sharded_sweep.py), then the sharded bench.
# (record of what produced profiles/r02_sweep8_c5_*.jsonl: the CHAINS / GRAPH / DEPTH / FUSED keys belong to experiments that were
#  measured and then removed from the library -- tools/sharded_sweep.py ignores keys it no longer knows)
mkdir -p gpurun_out
O=gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py > $O/r02_dist_check_${N}_ce.log 2>&1; echo "rc=$?" >> $O/r02_dist_check_${N}_ce.log
grep -E "dist_check|DIST_CHECK|rc=|rror" $O/r02_dist_check_${N}_ce.log | tail -16
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/sharded_sweep.py ${SWEEP_WL:-c5} \
  "MODE=sm,PHASES=4,PUSH=32" "MODE=ce,PHASES=4,CHAINS=2" "MODE=ce,PHASES=4,CHAINS=1" "MODE=ce,PHASES=4,CHAINS=3" "MODE=ce,PHASES=8,CHAINS=2" \
  "MODE=ce,PHASES=2,CHAINS=2" "MODE=ce,PHASES=1,CHAINS=7" "MODE=ce,PHASES=4,CHAINS=2,GRAPH=0" "MODE=sm,PHASES=4,PUSH=64" > $O/r02_sweep${N}_c5.log 2>&1
grep '^{' $O/r02_sweep${N}_c5.log | python -c "
import sys, json
for ln in sys.stdin:
    d = json.loads(ln); print(d['setting'], [round(x, 1) for x in d['ms']], d['converged'], d['info'], d['phases_ms'])"
tail -3 $O/r02_sweep${N}_c5.log | grep -v '^{'
