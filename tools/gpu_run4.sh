#!/bin/bash
# L2 prefetch distance sweep for the activation boxes of the persistent conv kernels
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
echo "== pytest gpu (model tests only)"; date +%s
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
for d in 0 2 4 8 16; do
echo "== prefetch $d"; date +%s
SB_PREFETCH_DIST=$d BENCH_VERBOSE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity --sustained-seconds 0 > $O/bench_pf$d.json 2> $O/bench_pf$d.err; echo "rc=$?"
python - <<PY
import json; d=json.load(open('gpurun_out/bench_pf$d.json')); print('prefetch $d:', d['value'], d['ms_per_step'], d['e2e']['value'])
PY
grep "^\[op" $O/bench_pf$d.err | awk '{printf "%s ", $4} END {print ""}'
done
date +%s
