#!/bin/bash
# First GPU call of round 2: everything that was written at the end of round 1 after the GPU budget ran out
# (optimizer-side sweep) gets its parity run, its bench numbers and its ncu evidence in ONE gpurun call:
#
#   gpurun --timeout 2400 -- 'bash tools/r02_first_gpu_session.sh'
#
# Outputs land in gpurun_out/ (copy what is kept into profiles/ as r02_*).
set -u
mkdir -p gpurun_out
echo "=== parity: optimizer sweep" | tee gpurun_out/r02_optim_tests.log
timeout 600 python -m pytest tests/test_gpu_zz_optim.py -x -q 2>&1 | tail -25 | tee -a gpurun_out/r02_optim_tests.log
echo "=== compute-sanitizer memcheck over the small optimizer tests" | tee -a gpurun_out/r02_optim_tests.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_zz_optim.py -x -q \
    -k "golden or seeded or without" 2>&1 | tail -15 | tee -a gpurun_out/r02_optim_tests.log
echo "=== parity: sampler options (seamless_loop / stereo_fix)" | tee gpurun_out/r02_sampler_options_tests.log
timeout 600 python -m pytest tests/test_gpu_zzz_sampler_options.py tests/test_gpu_zzz_dataset_encode.py -q 2>&1 | tail -25 | tee -a gpurun_out/r02_sampler_options_tests.log
echo "=== role timeline of the halo conv kernel (per-tile fixed cost, DESIGN.md section 4)"
DD_CONV_TRACE=1 timeout 300 python tools/trace_halo.py > gpurun_out/r02_trace_halo.log 2>&1
tail -40 gpurun_out/r02_trace_halo.log
echo "=== per-layer conv microbenchmarks: default, residual L2 prefetch, more halo stages / narrower weight panels"
timeout 400 python tools/bench_convs.py r02_base > gpurun_out/r02_bench_convs_base.log 2>&1
DD_EPI_PREFETCH=1 timeout 400 python tools/bench_convs.py r02_prefetch > gpurun_out/r02_bench_convs_prefetch.log 2>&1
DD_HALO_MIN_STAGES=4 timeout 400 python tools/bench_convs.py r02_stages4 > gpurun_out/r02_bench_convs_stages4.log 2>&1
DD_HALO_NTILE_MAX=64 timeout 400 python tools/bench_convs.py r02_ntile64 > gpurun_out/r02_bench_convs_ntile64.log 2>&1
tail -3 gpurun_out/r02_bench_convs_base.log gpurun_out/r02_bench_convs_prefetch.log gpurun_out/r02_bench_convs_stages4.log gpurun_out/r02_bench_convs_ntile64.log
echo "=== bench (all legs, optim_step last)"
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_n1_first.json 2> gpurun_out/r02_bench_n1_first.err
tail -c 1500 gpurun_out/r02_bench_n1_first.json
echo "=== optimizer sweep: descriptor search in global memory (default) vs shared memory"
timeout 300 python tools/profile_optim.py 2>&1 | tail -1 | tee gpurun_out/r02_optim_ab.log
DD_OPTIM_SMEM_SEARCH=1 timeout 300 python tools/profile_optim.py 2>&1 | tail -1 | tee -a gpurun_out/r02_optim_ab.log
DD_OPTIM_PERSISTENT=1 timeout 300 python tools/profile_optim.py 2>&1 | tail -1 | tee -a gpurun_out/r02_optim_ab.log
echo "--- parity of the two experimental variants" | tee -a gpurun_out/r02_optim_ab.log
DD_OPTIM_SMEM_SEARCH=1 timeout 300 python -m pytest tests/test_gpu_zz_optim.py -q -k "golden or seeded or without" 2>&1 | tail -2 | tee -a gpurun_out/r02_optim_ab.log
DD_OPTIM_PERSISTENT=1 timeout 300 python -m pytest tests/test_gpu_zz_optim.py -q -k "golden or seeded or without" 2>&1 | tail -2 | tee -a gpurun_out/r02_optim_ab.log
echo "=== ncu: launch list + full capture of dd_optim_step_batched"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"optim_step|grad_sqnorm|grad_norm_finish" \
    -c 12 --csv --log-file gpurun_out/r02_optim_launches.csv python tools/profile_optim.py > gpurun_out/r02_profile_optim.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:optim_step_batched -s 2 -c 1 \
    -o gpurun_out/r02_ncu_optim_step python tools/profile_optim.py >> gpurun_out/r02_profile_optim.log 2>&1
ncu -i gpurun_out/r02_ncu_optim_step.ncu-rep --page raw --csv 2>/dev/null | \
    grep -E "dram__bytes_(read|write).sum|gpu__time_duration.sum|dram__throughput.avg.pct|sm__warps_active.avg.pct|launch__grid_size|achieved_occupancy" \
    > gpurun_out/r02_ncu_optim_step_summary.csv
tail -5 gpurun_out/r02_profile_optim.log
