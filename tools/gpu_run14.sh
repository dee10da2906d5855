#!/bin/bash
export PYTHONUNBUFFERED=1
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q -x > $O/pytest_model.log 2>&1; echo rc=$?; tail -5 $O/pytest_model.log
SB_DEBUG=1 BENCH_VERBOSE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity --sustained-seconds 0 > $O/bench_quick.json 2> $O/bench_quick.err; echo "rc=$?"
python - <<'PY'
import json; d=json.load(open('gpurun_out/bench_quick.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])
PY
grep "mc2\|tconv" $O/bench_quick.err | tail -40 | cut -c1-400
grep "^\[op" $O/bench_quick.err | awk '{printf "%s ", $4} END {print ""}'
