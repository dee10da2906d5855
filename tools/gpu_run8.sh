#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
for a in 0 1 2 4 8 16 32 36 63; do
echo -n "ablate $a: "
SB_C01_ABLATE=$a SB_DEBUG=1 timeout 120 python tools/time_conv01.py 2>&1 | grep "first block" | tail -1
done
