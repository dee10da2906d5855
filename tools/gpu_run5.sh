#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
echo "== pytest gpu"; date +%s
timeout 900 python -m pytest tests -m gpu -q -x -rxXs > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu.log
NG=$(nvidia-smi -L | wc -l)
echo "gpus: $NG"
if [ "$NG" -ge 2 ]; then
for ex in p2p nccl; do
echo "== bench N=2 exchange=$ex"; date +%s
SB_EXCHANGE=$ex timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-parity > $O/bench_n2_$ex.json 2> $O/bench_n2_$ex.err; echo "rc=$?"
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_n2_$ex.json')); print('$ex', d['value'], d['ms_per_step'], d['e2e']['value'], d['sustained']['value'] if d.get('sustained') else None, d['config']['exchange'])
except Exception as e: print('parse failed', e)
PY
tail -5 $O/bench_n2_$ex.err
done
echo "== bench N=1"; date +%s
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity > $O/bench_n1.json 2> $O/bench_n1.err; echo "rc=$?"
python - <<PY
import json; d=json.load(open('gpurun_out/bench_n1.json')); print('n1', d['value'], d['ms_per_step'], d['e2e']['value'], d['sustained']['value'])
PY
fi
date +%s
