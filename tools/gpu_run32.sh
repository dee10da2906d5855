#!/bin/bash
export PYTHONUNBUFFERED=1
O=gpurun_out; mkdir -p $O
SAN_MODE=new bash tools/gpu_sanitize.sh
timeout 1200 python -m pytest tests -m gpu -q > $O/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02_pytest_gpu.log
