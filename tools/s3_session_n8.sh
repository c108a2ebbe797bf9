#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 tools/bench_train_only.py "$2" 2>&1 | grep -v "^W\|^\*\*\*\|Setting OMP" ; }
echo "== NCCL_MAX_CTAS=8"; NCCL_MAX_CTAS=8 run 29521 "max_ctas=8" | grep "train\|buckets"
echo "== NCCL_MAX_CTAS=4"; NCCL_MAX_CTAS=4 DD_DDP_TRACE=1 run 29522 "max_ctas=4" | grep "train\|buckets"
