#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
Q="--no-cpu-baseline --no-format --no-dae --no-train --no-gpu-eager --steps 30"
b() { python bench.py $Q 2> gpurun_out/s2j_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['gpu_launches'])"; }
echo "== parity (fuse all, default)"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_backward.py tests/test_gpu_zz_b4_2.py tests/test_gpu_dae.py tests/test_gpu_ddec.py -x -q 2>&1 | tail -3
echo "== parity fuse off"; DD_FUSE_CAT_LEVEL=99 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "unet or sampler" 2>&1 | tail -1
for fl in 99 0 2 1 99 0; do echo "== bench DD_FUSE_CAT_LEVEL=$fl"; DD_FUSE_CAT_LEVEL=$fl b; done
