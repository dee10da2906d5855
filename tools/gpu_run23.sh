#!/bin/bash
export PYTHONUNBUFFERED=1
O=gpurun_out; mkdir -p $O
timeout 500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_head -f -o $O/head_kernel python bench.py --steps 1 --warmup 3 --ncu-step > $O/ncu_head.log 2>&1
echo "rc=$?"
ncu -i $O/head_kernel.ncu-rep --page details 2>/dev/null | grep -E "k_head|Duration|DRAM Throughput|Memory Throughput|L1/TEX Hit|L2 Hit|Achieved Occupancy|Theoretical Occupancy|Registers|Issue Slots Busy|No Eligible|Warp Cycles Per Issued|Stall|Executed Ipc|Block Limit|Eligible Warps|Mem Busy|Max Bandwidth|Dynamic Shared|Waves Per SM|One or More|Est. Speedup|stall" | head -80
