#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== full gpu tests"; (time timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputests_s2k.log 2>&1); tail -3 gpurun_out/r02_gputests_s2k.log
echo "== bench (sampler + train + optim)"; python bench.py --no-cpu-baseline --no-format --no-dae --no-gpu-eager --steps 30 > gpurun_out/r02_bench_s2k.json 2> gpurun_out/r02_bench_s2k.err; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_s2k.json'))
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['launches'], d['roofline']['avg_launch_us'])
print('train', d['train_step']['value'], d['train_step']['ms_per_step'])
print('optim', d['optim_step']['ms'], d['optim_step']['roofline']['frac'], d['optim_step']['train_step_with_optimizer']['value'])
"
