#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
Q="--no-cpu-baseline --no-format --no-dae --no-train --no-gpu-eager --steps 30"
b() { python bench.py $Q 2> gpurun_out/s2c_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['gpu_launches'])"; }
echo "== parity"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_backward.py tests/test_gpu_dae.py -x -q 2>&1 | tail -4
echo "== launch floor"; python tools/launch_floor.py 2>&1 | tail -6
echo "== bench"; b
echo "== bench prefetch 2"; DD_L2_PREFETCH_LEVEL=2 b
echo "== bench fuse 3"; DD_FUSE_CAT_LEVEL=3 b
echo "== bench fuse 0"; DD_FUSE_CAT_LEVEL=0 b
echo "== prefix"; python tools/prefix_times.py > gpurun_out/prefix_times.log 2>&1; tail -1 gpurun_out/prefix_times.log; cp gpurun_out/prefix_times.csv gpurun_out/prefix_times_c.csv
echo "== full bench"; python bench.py > gpurun_out/r02_bench_s2c.json 2> gpurun_out/r02_bench_s2c.err; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_s2c.json'))
print(d['value'], d['e2e']['value'], d['roofline']['frac'])
for k in ['train_step','dae_decode','ddec_forward']: print(k, d[k]['value'], d[k].get('roofline',{}).get('frac'))
print(d['optim_step']['train_step_with_optimizer']['value'])
"
