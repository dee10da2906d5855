mkdir -p gpurun_out
timeout 75 python -m pytest tests -q -m gpu -x > gpurun_out/t6.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/t6.log
BENCH_VERBOSE=1 timeout 60 python bench.py --no-cpu-baseline --steps 20 --warmup 3 > gpurun_out/bench6.json 2> gpurun_out/bench6.err; echo "bench rc=$?"
grep "^\[op" gpurun_out/bench6.err | head -8
python -c "
import json;d=json.load(open('gpurun_out/bench6.json'));print(d['value'],d['e2e']['value'],d['roofline']['frac'],d['ms_per_step'])"
