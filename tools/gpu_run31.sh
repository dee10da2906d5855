#!/bin/bash
# 8 GPUs: the bench line with the peer-memory exchange (and PDL) at full width
export PYTHONUNBUFFERED=1
O=gpurun_out; mkdir -p $O
N=${N:-8}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-parity --sustained-seconds 0 > $O/bench_${N}gpu.json 2> $O/bench_${N}gpu.err; echo "rc=$?"
python - <<PY
import json; d=json.load(open('gpurun_out/bench_${N}gpu.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['exchange'][:90])
PY
tail -4 $O/bench_${N}gpu.err | cut -c1-300
