#!/bin/bash
# round 2, session 2, GPU call A: fused decoder concat (strided conv outputs + skip slices) and the tap-stacked 3x3 kernel
# on images shorter than a tile (DD_DX_MIN_H)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
Q="--no-cpu-baseline --no-format --no-dae --no-train --no-gpu-eager --steps 30"
echo "== parity (new schedule)"; timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_b4_2.py -x -q -k "mpconv or scale_silu or unet or sampler or elementwise or b4_2" 2>&1 | tail -4
echo "== parity DD_DX_MIN_H=2"; DD_DX_MIN_H=2 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "mpconv_vs_oracle or unet_small or unet_default or concat_slices" 2>&1 | tail -4
echo "== parity fuse all"; DD_FUSE_CAT_LEVEL=0 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "unet_small or unet_default" 2>&1 | tail -2
b() { python bench.py $Q 2> gpurun_out/s2a_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['gpu_launches'])"; }
for fl in 99 3 2 0; do echo "== bench DD_FUSE_CAT_LEVEL=$fl"; DD_FUSE_CAT_LEVEL=$fl b; done
for mh in 4 2; do echo "== bench DD_DX_MIN_H=$mh"; DD_DX_MIN_H=$mh b; done
echo "== prefix times DD_DX_MIN_H=2"; DD_DX_MIN_H=2 python tools/prefix_times.py > gpurun_out/prefix_times.log 2>&1; tail -2 gpurun_out/prefix_times.log; mv gpurun_out/prefix_times.csv gpurun_out/prefix_times_mh2.csv
