#!/bin/bash
# Scratch: N-GPU value-only runs of bench.py under a few NCCL settings (development aid).
N=${1:-8}
run() { echo "== $1"; env $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus $N --steps 40 --warmup 5 --value-only 2>/dev/null | tail -1; }
run "X=0"
run "NCCL_ALGO=NVLS"
run "NCCL_ALGO=Ring"
run "NCCL_ALGO=Tree"
run "GSR_EXACT_BINNING=1"
