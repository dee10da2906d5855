#!/bin/bash
export PYTHONUNBUFFERED=1
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader -lms 100 > $O/probe_clocks.csv &
SMI=$!
MMA_PROBE_OFFSETS_ONLY=1 timeout 120 tools/probes/mma_probe > $O/mma_probe_offsets.txt 2>&1; echo "probe rc=$?"; tail -9 $O/mma_probe_offsets.txt
kill $SMI
sort $O/probe_clocks.csv | uniq -c | sort -rn | head -8
