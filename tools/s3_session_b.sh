#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== parity"; timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_backward.py tests/test_gpu_zz_optim.py -x -q 2>&1 | tail -3
echo "== bench with train"; python bench.py --no-cpu-baseline --no-format --no-dae --no-gpu-eager --steps 30 2> gpurun_out/s3b_bench.err | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step']); t=d['train_step']; print({k:(round(v,2) if isinstance(v,float) else v) for k,v in t.items() if not isinstance(v,(dict,list))}); print(t.get('with_optimizer')); print(t.get('global_batch_32_on_one_gpu'))"
echo "== prefix times"; timeout 600 python tools/prefix_times.py > gpurun_out/s3b_prefix.log 2>&1; tail -3 gpurun_out/s3b_prefix.log; cp gpurun_out/prefix_times.csv gpurun_out/r02_prefix_times_s3.csv
