#!/bin/bash
# cta_group::2 pair twins: parity first (bounded), then the whole model suite, then autotune timings + bench line
export PYTHONUNBUFFERED=1
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_model.py -m gpu -q -x -k "test_tc_single_layers and (256-128 or 128-256 or 256-512 or 128-128)" > $O/pytest_pair.log 2>&1; echo "pair rc=$?"; tail -8 $O/pytest_pair.log
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q -x > $O/pytest_model.log 2>&1; echo rc=$?; tail -5 $O/pytest_model.log
SB_DEBUG=1 BENCH_VERBOSE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity --sustained-seconds 0 > $O/bench_quick.json 2> $O/bench_quick.err; echo "rc=$?"
python - <<'PY'
import json; d=json.load(open('gpurun_out/bench_quick.json')); r=d['roofline']; print(d['value'], d['ms_per_step'], d['e2e']['value'], r['frac'], r['kernel_ms_per_step'], r['kernel_share_of_step'], r.get('per_op_sum_ms'))
PY
grep "2cta\|tconv" $O/bench_quick.err | tail -40 | cut -c1-600
grep "^\[op" $O/bench_quick.err | awk '{printf "%s ", $4} END {print ""}'
