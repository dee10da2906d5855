#!/bin/bash
# programmatic dependent launch: parity suite, then bench with / without it
export PYTHONUNBUFFERED=1
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_split_precision.py tests/test_gpu_zz_full_size.py tests/test_gpu_reference_models.py -m gpu -q -x > $O/pytest_pdl.log 2>&1; echo rc=$?; tail -4 $O/pytest_pdl.log
for D in 0 1; do
  if [ $D = 1 ]; then export SB_DISABLE_PDL=1; fi
  BENCH_VERBOSE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity --sustained-seconds 0 > $O/bench_pdl_$D.json 2> $O/bench_pdl_$D.err; echo "rc=$?"
  python - <<PY
import json; d=json.load(open('gpurun_out/bench_pdl_$D.json')); r=d['roofline']; print('disable_pdl=$D', d['value'], d['ms_per_step'], d['e2e']['value'], r['frac'], r['kernel_ms_per_step'])
PY
  grep "^\[op" $O/bench_pdl_$D.err | awk '{printf "%s ", $4} END {print ""}'
done
