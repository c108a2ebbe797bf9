#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== parity (3-D weight box in the halo kernel)"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dae.py tests/test_gpu_ddec.py -x -q 2>&1 | tail -3
for dm in 0 128 256 512; do
  echo "== DD_DX_DENSE_MAX=$dm"
  DD_DX_DENSE_MAX=$dm timeout 600 python -m pytest tests/test_gpu_dae.py tests/test_gpu_ddec.py -x -q 2>&1 | tail -1
  DD_DX_DENSE_MAX=$dm python tools/bench_legs.py 2>&1 | tail -2
done
