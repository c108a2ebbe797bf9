#!/bin/bash
# ncu durations of a few degenerate MPConv shapes: separates the fixed launch cost from the per-k-iteration / per-tile cost
for sh in "2 2 43 64 64 1 1" "2 2 43 1280 64 1 1" "2 2 43 2560 64 1 1" "2 2 43 64 1280 1 1" "2 2 43 1280 1280 1 1" "2 2 43 1280 2560 1 1" "2 2 43 2560 1280 3 8" "2 4 86 1024 1024 1 1" "2 32 688 256 256 1 1"; do
  ncu --metrics gpu__time_duration.sum --clock-control none --csv python tools/one_conv.py $sh 2>/dev/null | grep conv_igemm | tail -2 | awk -F'","' -v s="$sh" '{print s, "|", $5, "| grid", $9, "|", $NF}' | cut -c1-160
done
