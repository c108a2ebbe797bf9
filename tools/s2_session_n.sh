#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== attention tests"; timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_backward.py tests/test_gpu_zz_b4_2.py tests/test_gpu_ddec.py tests/test_gpu_zzz_sampler_options.py -x -q -k "attention or unet or sampler or ddec or b4_2" 2>&1 | tail -3
echo "== launch floor (attention)"; python tools/launch_floor.py 2>&1 | tail -8 | head -2
echo "== ncu spectral kernels"
for k in fgla_istft_kernel fgla_stft_update_kernel stft_mel_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/r02_ncu_$k python tools/bench_format.py 16 3 > gpurun_out/r02_ncu_spectral.log 2>&1
done
ls -la gpurun_out/r02_ncu_fgla* gpurun_out/r02_ncu_stft*
