#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
Q="--no-cpu-baseline --no-format --no-dae --no-train --no-gpu-eager --steps 30"
b() { python bench.py $Q 2> gpurun_out/s2b_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['gpu_launches'])"; }
echo "== launch floor, PDL on"; python tools/launch_floor.py 2>&1 | tail -6
echo "== launch floor, PDL off"; DD_DISABLE_PDL=1 python tools/launch_floor.py 2>&1 | tail -6
echo "== parity prefetch"; timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "unet_small or unet_default or sampler" 2>&1 | tail -2
for pl in 99 2 99 3 0; do echo "== bench DD_L2_PREFETCH_LEVEL=$pl"; DD_L2_PREFETCH_LEVEL=$pl b; done
echo "== bench no side streams"; DD_NO_SIDE=1 b
echo "== bench no side streams, no prefetch"; DD_NO_SIDE=1 DD_L2_PREFETCH_LEVEL=99 b
echo "== bench fuse 3"; DD_FUSE_CAT_LEVEL=3 b
echo "== bench no PDL"; DD_DISABLE_PDL=1 b
