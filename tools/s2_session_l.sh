#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
DD_CONV_TRACE=1 python tools/trace_halo.py 2 32 688 512 256 8 2 > gpurun_out/r02_trace_dx_s2l_512to256.log 2>&1; grep -A3 "CTA 1\|CTA 2" gpurun_out/r02_trace_dx_s2l_512to256.log | grep "period\|entry\|set-up" 
DD_CONV_TRACE=1 python tools/trace_halo.py 2 4 86 2048 1024 8 2 > gpurun_out/r02_trace_dx_s2l_l3.log 2>&1; grep "period\|entry\|set-up\|first tiles" gpurun_out/r02_trace_dx_s2l_l3.log | head -12
