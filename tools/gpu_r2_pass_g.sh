#!/bin/bash
# round 2, pass G (N GPUs): parity of the single-launch phase-split SpMV (tests/dist_check.py), then fused vs per-phase launches x
# transport on config 5
# (record of what produced profiles/r02_sweep8_c5_*.jsonl: the CHAINS / GRAPH / DEPTH / FUSED keys belong to experiments that were
#  measured and then removed from the library -- tools/sharded_sweep.py ignores keys it no longer knows)
mkdir -p gpurun_out
O=gpurun_out
N=$(nvidia-smi -L | wc -l)
PROPACK_B200_PUSH=${DC_PUSH:-sm} timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py > $O/r02_dist_check_${N}_fused.log 2>&1; echo "rc=$?" >> $O/r02_dist_check_${N}_fused.log
grep -E "dist_check|DIST_CHECK|rc=|rror" $O/r02_dist_check_${N}_fused.log | tail -16
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/sharded_sweep.py ${SWEEP_WL:-c5} \
  "MODE=sm,PHASES=4,PUSH=64,FUSED=1" "MODE=sm,PHASES=4,PUSH=64,FUSED=0" "MODE=sm,PHASES=8,PUSH=64,FUSED=1" "MODE=sm,PHASES=2,PUSH=64,FUSED=1" \
  "MODE=ce,PHASES=4,CHAINS=1,GRAPH=0,FUSED=1" "MODE=ce,PHASES=8,CHAINS=1,GRAPH=0,FUSED=1" "MODE=ce,PHASES=8,CHAINS=2,FUSED=1" "MODE=sm,PHASES=8,PUSH=32,FUSED=1" > $O/r02_sweep${N}_c5_fused.log 2>&1
grep '^{' $O/r02_sweep${N}_c5_fused.log | python -c "
import sys, json
for ln in sys.stdin:
    d = json.loads(ln); print(d['setting'], [round(x, 1) for x in d['ms']], d['converged'], d['info'], d['phases_ms'], d.get('spmv_isolated'))"
grep -v '^{' $O/r02_sweep${N}_c5_fused.log | grep -E "RuntimeError" | head -3; true
