#!/bin/bash
export PYTHONUNBUFFERED=1
O=gpurun_out; mkdir -p $O
run() {
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity --sustained-seconds 0 > $O/b.json 2> $O/b.err; echo "rc=$? [$*]"
  python - <<PY
import json; d=json.load(open('gpurun_out/b.json')); r=d['roofline']; print('   ', d['value'], d['ms_per_step'], d['e2e']['value'], r['frac'], r['kernel_ms_per_step'])
PY
}
run A=1
run SB_FORCE_FUSED_TCONV=1
run SB_FORCE_FUSED_TCONV=2
run SB_DISABLE_FORK=1
run A=2
