#!/bin/bash
export PYTHONUNBUFFERED=1
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q -x -k "stem or hourglass or (single_layers and (5- or 7-))" > $O/pytest_k57.log 2>&1; echo rc=$?; tail -15 $O/pytest_k57.log
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo rc=$?; tail -4 $O/pytest_gpu.log
timeout 300 python tools/bench_configs.py > $O/configs.jsonl 2> $O/configs.err; echo "configs rc=$?"; cat $O/configs.jsonl; tail -3 $O/configs.err
