#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== backward tests"; timeout 900 python -m pytest tests/test_gpu_backward.py tests/test_gpu_zz_optim.py -x -q 2>&1 | tail -3
echo "== bench (train)"; python bench.py --no-cpu-baseline --no-format --no-dae --no-gpu-eager --steps 20 > gpurun_out/r02_bench_s2o.json 2> gpurun_out/r02_bench_s2o.err; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_s2o.json'))
print(d['value'])
print('train', d['train_step']['value'], d['train_step']['ms_per_step'])
print('optim', d['optim_step']['ms'], d['optim_step']['train_step_with_optimizer']['value'])
"
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
