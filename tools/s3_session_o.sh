#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
o() { python -c "
import sys, torch
sys.path.insert(0, '.')
import bench
d = bench.bench_optim(torch.device('cuda:0'))
print('optim', round(d['ms'], 3), 'ms', round(d['roofline']['frac'], 3))
" 2>&1 | tail -1; }
echo "== default"; o
echo "== warp per row"; DD_OPTIM_WARP=1 o
echo "== optim tests, warp per row"; DD_OPTIM_WARP=1 timeout 300 python -m pytest tests/test_gpu_zz_optim.py -x -q 2>&1 | tail -2
