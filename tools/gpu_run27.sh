#!/bin/bash
# 2 GPUs: peer-memory record exchange tests + a 2-rank bench line
export PYTHONUNBUFFERED=1
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_gather.py -m gpu -q > $O/pytest_gather.log 2>&1; echo "gather rc=$?"; tail -3 $O/pytest_gather.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-parity --sustained-seconds 0 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; echo "rc=$?"
python - <<PY
import json; d=json.load(open('gpurun_out/bench_2gpu.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['exchange'][:80])
PY
tail -3 $O/bench_2gpu.err | cut -c1-300
