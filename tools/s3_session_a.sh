#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
Q="--no-cpu-baseline --no-format --no-dae --no-train --no-gpu-eager --steps 30"
b() { python bench.py $Q 2> gpurun_out/s3a_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['gpu_launches'])"; }
echo "== parity"; timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dae.py tests/test_gpu_backward.py tests/test_gpu_zz_b4_2.py tests/test_gpu_zzzz_full_size.py -x -q 2>&1 | tail -3
echo "== bench"; b
echo "== bench again"; b
echo "== legs"; python tools/bench_legs.py 2>&1 | tail -4
