#!/bin/bash
# post-processing rewrite: parity tests + per-kernel device times (ncu launch list of 3 warm steps)
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
echo "== pytest gpu"; date +%s
timeout 900 python -m pytest tests -m gpu -q -x -rxXs > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu.log
echo "== tune save"; date +%s
SB_TUNE_SAVE=$O/tune.txt timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-parity --sustained-seconds 0 > $O/bench_quick.json 2> $O/bench_quick.err; echo "rc=$?"
python - <<'PY'
import json; d=json.load(open('gpurun_out/bench_quick.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'])
PY
echo "== ncu launch list"; date +%s
SB_TUNE_LOAD=$O/tune.txt timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none --profile-from-start off --csv --log-file $O/launches.csv python bench.py --steps 3 --warmup 3 --ncu-step > $O/ncu_list.log 2>&1
echo "list rc=$?"
python tools/ncu_summarize.py $O/launches.csv $O/launches_summary.md $O/tc_traffic.json 3 | head -60
date +%s
