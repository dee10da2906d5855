#!/bin/bash
# Round-2 first GPU call: full GPU suite (xfail markers dropped), first executions of the opt-in kernels, sanitizer, bench.
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
echo "== pytest gpu"; date +%s
timeout 900 python -m pytest tests -m gpu -q -x -rxXs > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu.log
echo "== multicast experimental"; date +%s
SB_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_model.py -k multicast -q > $O/pytest_mc.log 2>&1; echo "mc rc=$?"; tail -15 $O/pytest_mc.log
echo "== scan4"; date +%s
SB_ENABLE_SCAN4=1 timeout 300 python -m pytest tests/test_gpu_peaks.py tests/test_gpu_paf.py -q > $O/pytest_scan4.log 2>&1; echo "scan4 rc=$?"; tail -5 $O/pytest_scan4.log
echo "== bench"; date +%s
SB_DEBUG=1 BENCH_VERBOSE=1 SB_TUNE_SAVE=$O/tune.txt timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_1gpu.json 2> $O/bench_1gpu.err
echo "bench rc=$?"; tail -c 3000 $O/bench_1gpu.json
echo "== sanitizer (memcheck) smoke"; date +%s
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > $O/sanitizer_memcheck_smoke.log 2>&1; echo "memcheck rc=$?"; tail -8 $O/sanitizer_memcheck_smoke.log
date +%s
