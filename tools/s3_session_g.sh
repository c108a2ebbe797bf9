#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== train-step launch list"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_train_step_launches_v3.csv python tools/profile_train.py > gpurun_out/s3g_profile_train.log 2>&1
tail -2 gpurun_out/s3g_profile_train.log
python tools/summarize_launches.py gpurun_out/r02_train_step_launches_v3.csv | head -40
