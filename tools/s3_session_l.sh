#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== parity"; timeout 1200 python -m pytest tests/test_gpu_backward.py tests/test_gpu_parity.py -x -q 2>&1 | tail -3
for i in 1 2; do
echo "== bench with train ($i)"; python bench.py --no-cpu-baseline --no-format --no-dae --no-gpu-eager --steps 30 2> gpurun_out/s3l_bench.err | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step']); t=d['train_step']; print(t['value'], t['ms_per_step']); print(t['with_optimizer']['value'], t['global_batch_32_on_one_gpu']['value'], t['global_batch_32_on_one_gpu']['ms_per_step']); print(d['optim_step']['ms'])"
done
