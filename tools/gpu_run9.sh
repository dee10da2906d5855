#!/bin/bash
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_model.py -m gpu -q -x -k conv01 2>&1 | tail -3
SB_C01_TIMING=1 SB_DEBUG=1 timeout 120 python tools/time_conv01.py 2>&1 | grep "k_conv01 timing" | tail -14
SB_DEBUG=1 timeout 120 python tools/time_conv01.py 2>&1 | grep "first block" | tail -1
