#!/bin/bash
# precision 2 (split fp16): parity tests, then C4 timing of the split path, then the fp16 path (register-count regression check)
export PYTHONUNBUFFERED=1
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_split_precision.py -m gpu -q -x > $O/pytest_split.log 2>&1; echo "split rc=$?"; tail -15 $O/pytest_split.log
SB_DEBUG=1 BENCH_VERBOSE=1 timeout 300 python bench.py --precision split --steps 10 --warmup 3 --no-cpu-baseline --no-parity --sustained-seconds 0 > $O/bench_split.json 2> $O/bench_split.err; echo "rc=$?"
tail -3 $O/bench_split.err | cut -c1-300
python - <<'PY'
import json; d=json.load(open('gpurun_out/bench_split.json')); r=d['roofline']; print('split', d['value'], d['ms_per_step'], d['e2e']['value'], r['frac'], r['kernel_ms_per_step'])
PY
grep "^\[op" $O/bench_split.err | awk '{printf "%s ", $4} END {print ""}'
BENCH_VERBOSE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity --sustained-seconds 0 > $O/bench_quick.json 2> $O/bench_quick.err; echo "rc=$?"
python - <<'PY'
import json; d=json.load(open('gpurun_out/bench_quick.json')); r=d['roofline']; print('fp16', d['value'], d['ms_per_step'], d['e2e']['value'], r['frac'], r['kernel_ms_per_step'])
PY
grep "^\[op" $O/bench_quick.err | awk '{printf "%s ", $4} END {print ""}'
