mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_reference_models.py -q -m gpu -x > gpurun_out/t_ref.log 2>&1; echo "ref rc=$?"; tail -15 gpurun_out/t_ref.log
timeout 400 python -m pytest tests/test_gpu_model.py -q -m gpu > gpurun_out/t_model.log 2>&1; echo "model rc=$?"; tail -15 gpurun_out/t_model.log
SB_DEBUG=1 BENCH_VERBOSE=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench2.json 2> gpurun_out/bench2.err; echo "bench rc=$?"
grep "op 1 \|^\[op  [0-5]\]" gpurun_out/bench2.err | tail -12 | cut -c1-400
python -c "
import json;d=json.load(open('gpurun_out/bench2.json'));print(d['value'],d['e2e']['value'],d['roofline']['frac'],d['ms_per_step'])"
