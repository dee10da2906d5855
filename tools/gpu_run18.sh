#!/bin/bash
# where does the time of the weight-streamed big layers go?  autotune table (all variants) under SB_ABLATE masks
export PYTHONUNBUFFERED=1
O=gpurun_out; mkdir -p $O
for L in "512 256 3 64 64 8" "256 128 3 128 128 8" "512 512 3 32 32 8" "128 64 3 256 256 8"; do
  for A in 0 1 2 4 5; do
    echo "== layer $L ablate $A"
    SB_ABLATE=$A SB_DEBUG=1 timeout 120 python tools/prof_layer.py $L 1 2>&1 | tee -a $O/ablate_full.log | grep "sb_conv_tc\] op 2 " | cut -c1-900
  done
done
