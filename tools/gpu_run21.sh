#!/bin/bash
# head kernel: parity (model suite + split suite), then bench with the default / a larger shared-memory budget
export PYTHONUNBUFFERED=1
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_split_precision.py tests/test_gpu_reference_models.py -m gpu -q -x > $O/pytest_model.log 2>&1; echo rc=$?; tail -5 $O/pytest_model.log
for KB in 196 218; do
SB_SMEM_BUDGET_KB=$KB SB_DEBUG=1 BENCH_VERBOSE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity --sustained-seconds 0 > $O/bench_quick_$KB.json 2> $O/bench_quick_$KB.err; echo "rc=$?"
python - <<PY
import json; d=json.load(open('gpurun_out/bench_quick_$KB.json')); r=d['roofline']; print('budget $KB', d['value'], d['ms_per_step'], d['e2e']['value'], r['frac'], r['kernel_ms_per_step'])
PY
grep "^\[op" $O/bench_quick_$KB.err | awk '{printf "%s ", $4} END {print ""}'
done
