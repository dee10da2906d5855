#!/bin/bash
export PYTHONUNBUFFERED=1
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_split_precision.py -m gpu -q > $O/pytest_split.log 2>&1; echo "split rc=$?"; tail -15 $O/pytest_split.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sustained-seconds 0 > $O/bench_strict.json 2> $O/bench_strict.err; echo "rc=$?"
tail -3 $O/bench_strict.err | cut -c1-300
python - <<'PY'
import json; d=json.load(open('gpurun_out/bench_strict.json')); print(d['value']); print(json.dumps(d['parity'], indent=1)); print(json.dumps(d['strict_tensor_core'], indent=1))
PY
