#!/bin/bash
# round 2 ncu evidence: launch list of one UNet evaluation + full captures of the dominant kernels + role timelines
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/r02_unet_fwd_launches_v2.csv python tools/profile_unet.py > gpurun_out/r02_profile_unet.log 2>&1
tail -1 gpurun_out/r02_profile_unet.log
echo "== full capture: tap-stacked 3x3 kernel (level-0 res0 256->512 and res1 512->256)"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv3x3_dx -c 2 \
    -o gpurun_out/r02_ncu_dx_l0 python tools/profile_unet.py >> gpurun_out/r02_profile_unet.log 2>&1
echo "== full capture: per-tap kernel, 1x1 at level 0 and at level 4"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_igemm_kernel -s 1 -c 1 \
    -o gpurun_out/r02_ncu_igemm_l0 python tools/profile_unet.py >> gpurun_out/r02_profile_unet.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_igemm_kernel -s 40 -c 1 \
    -o gpurun_out/r02_ncu_igemm_l4 python tools/profile_unet.py >> gpurun_out/r02_profile_unet.log 2>&1
echo "== full capture: attention (344 tokens)"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attention_kernel -c 1 \
    -o gpurun_out/r02_ncu_attention python tools/profile_unet.py >> gpurun_out/r02_profile_unet.log 2>&1
ls -la gpurun_out/*.ncu-rep
echo "== role timelines (tap-stacked kernel)"
DD_CONV_TRACE=1 python tools/trace_halo.py 2 32 688 512 256 8 2 > gpurun_out/r02_trace_dx_s2_512to256.log 2>&1; tail -22 gpurun_out/r02_trace_dx_s2_512to256.log
DD_CONV_TRACE=1 python tools/trace_halo.py 2 32 688 256 512 8 1 > gpurun_out/r02_trace_dx_s2_256to512.log 2>&1; tail -22 gpurun_out/r02_trace_dx_s2_256to512.log
