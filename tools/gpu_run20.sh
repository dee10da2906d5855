#!/bin/bash
# ablation masks incl. 8 (no TMEM loads): is the MMA issue loop or the epilogue the limiter of the resident-weight layers?
export PYTHONUNBUFFERED=1
O=gpurun_out; mkdir -p $O
for L in "128 64 3 256 256 8" "64 64 3 256 256 8" "32 32 3 512 512 8" "256 128 3 128 128 8"; do
  for A in 0 8 12 13 14; do
    echo "== layer $L ablate $A"
    SB_ABLATE=$A SB_DEBUG=1 timeout 120 python tools/prof_layer.py $L 1 2>&1 | tee -a $O/ablate2_full.log | grep "sb_conv_tc\] op 2 " | sed 's/persist 9999[0-9.]*/persist -/' | cut -c1-700
  done
done
