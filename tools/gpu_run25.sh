#!/bin/bash
export PYTHONUNBUFFERED=1
O=gpurun_out; mkdir -p $O
MMA_PROBE_OFFSETS_ONLY=1 timeout 120 tools/probes/mma_probe > $O/mma_probe_offsets.txt 2>&1; echo "probe rc=$?"; cat $O/mma_probe_offsets.txt
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_zz_full_size.py tests/test_gpu_split_precision.py -m gpu -q -x -k "forward or bottomup or full or split or c4" > $O/pytest_head.log 2>&1; echo rc=$?; tail -3 $O/pytest_head.log
BENCH_VERBOSE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity --sustained-seconds 0 > $O/bench_quick.json 2> $O/bench_quick.err; echo "rc=$?"
python - <<PY
import json; d=json.load(open('gpurun_out/bench_quick.json')); r=d['roofline']; print(d['value'], d['ms_per_step'], d['e2e']['value'], r['frac'], r['kernel_ms_per_step'])
PY
grep "^\[op" $O/bench_quick.err | awk '{printf "%s ", $4} END {print ""}'
timeout 600 python tools/bench_configs.py c5 --steps 9 > $O/c5_b16.jsonl 2> $O/c5_b16.err; echo "c5 rc=$?"; cat $O/c5_b16.jsonl | cut -c1-900; tail -3 $O/c5_b16.err | cut -c1-300
