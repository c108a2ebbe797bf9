#!/bin/bash
# runs every developer GPU check in its own process under a timeout; logs to gpurun_out/dev_check.log
mkdir -p gpurun_out
for c in "$@"; do
  echo "=== $c" 
  timeout 180 python tools/dev_check_ops.py $c 2>&1 | tail -25
  echo "exit: $?"
done
