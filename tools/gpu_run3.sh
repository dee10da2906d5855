#!/bin/bash
# dead-store elimination + staged fp32 head stores + single-record D2H; multicast streaming kernel as an autotune experiment
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
echo "== pytest gpu"; date +%s
timeout 900 python -m pytest tests -m gpu -q -x -rxXs > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu.log
echo "== bench quick"; date +%s
SB_DEBUG=1 BENCH_VERBOSE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity --sustained-seconds 0 > $O/bench_quick.json 2> $O/bench_quick.err; echo "rc=$?"
python - <<'PY'
import json; d=json.load(open('gpurun_out/bench_quick.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])
PY
grep "^\[op" $O/bench_quick.err
for cs in 2 4; do
echo "== multicast $cs"; date +%s
SB_ENABLE_MULTICAST=$cs SB_DEBUG=1 BENCH_VERBOSE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity --sustained-seconds 0 > $O/bench_mc$cs.json 2> $O/bench_mc$cs.err; echo "rc=$?"
python - <<PY
import json; d=json.load(open('gpurun_out/bench_mc$cs.json')); print(d['value'], d['ms_per_step'])
PY
grep "^\[op" $O/bench_mc$cs.err | head -40
done
date +%s
