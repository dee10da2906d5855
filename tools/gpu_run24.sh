#!/bin/bash
export PYTHONUNBUFFERED=1
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo rc=$?; tail -5 $O/pytest_gpu.log
SB_DEBUG=1 BENCH_VERBOSE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity --sustained-seconds 0 > $O/bench_quick.json 2> $O/bench_quick.err; echo "rc=$?"
python - <<PY
import json; d=json.load(open('gpurun_out/bench_quick.json')); r=d['roofline']; print(d['value'], d['ms_per_step'], d['e2e']['value'], r['frac'], r['kernel_ms_per_step'])
PY
grep "^\[op" $O/bench_quick.err | awk '{printf "%s ", $4} END {print ""}'
