#!/bin/bash
# Scratch: N-GPU value-only runs of bench.py (development aid): overlapped vs plain all-reduce.
N=${1:-8}
run() { echo "== $1"; env $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus $N --steps 40 --warmup 5 --value-only 2>/dev/null | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print(round(d['value'],1), 'views/s', round(d['ms_per_step'],3), 'ms', {k: round(v,3) for k,v in d['stage_ms'].items()})
except Exception as e: print('failed', e)"; }
run "X=0"
run "GSR_PLAIN_ALLREDUCE=1"
run "X=0"
run "GSR_PLAIN_ALLREDUCE=1"
