#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
Q="--no-cpu-baseline --no-format --no-dae --no-train --no-gpu-eager --steps 30"
b() { python bench.py $Q 2> gpurun_out/s2d_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['gpu_launches'])"; }
echo "== launch floor"; python tools/launch_floor.py 2>&1 | tail -8
echo "== bench"; b
echo "== full gpu tests"; (time timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputests_s2d.log 2>&1); tail -4 gpurun_out/r02_gputests_s2d.log
