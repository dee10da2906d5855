#!/bin/bash
# One gpurun call that produces every artefact profiles/ is built from:
#   bench line (+ autotune picks), ncu launch list with DRAM bytes, ncu --set full of one warm step, gpu tests.
# usage (from the repo root on the GPU box): bash tools/gpu_evidence.sh [skip_tests]
set -u
O=gpurun_out
PFX=${PFX:-r02}          # round prefix of the ncu artefacts (profiles/${PFX}_*)
mkdir -p $O
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
./tools/probes/store_probe > $O/store_probe.txt 2>&1
timeout 60 python tools/bw_probe.py > $O/bw_probe.txt 2>&1
echo "== bench"; date +%s
SB_DEBUG=1 BENCH_VERBOSE=1 SB_TUNE_SAVE=$O/tune.txt timeout 420 python bench.py > $O/bench_1gpu.json 2> $O/bench_1gpu.err
echo "bench rc=$?"; tail -c 600 $O/bench_1gpu.json
echo "== reference arm"; date +%s
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
echo "ref rc=$?"
echo "== ncu launch list"; date +%s
SB_TUNE_LOAD=$O/tune.txt timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_list.log 2>&1
echo "list rc=$?"
echo "== ncu full (one warm step)"; date +%s
SB_TUNE_LOAD=$O/tune.txt timeout 500 ncu --set full --clock-control none --profile-from-start off \
  -f -o $O/${PFX}_step_full python bench.py --steps 1 --warmup 2 --no-cpu-baseline --ncu-step > $O/ncu_full.log 2>&1
echo "full rc=$?"
timeout 120 ncu -i $O/${PFX}_step_full.ncu-rep --page raw --csv > $O/${PFX}_step_full_raw.csv 2>/dev/null
echo "== ncu source-level capture of the top kernel (one launch)"; date +%s
SB_TUNE_LOAD=$O/tune.txt timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:k_conv_tc_halo -c 2 -f -o $O/${PFX}_top_kernel python bench.py --steps 1 --warmup 2 --no-cpu-baseline --ncu-step > $O/ncu_top.log 2>&1
echo "top rc=$?"
timeout 120 ncu -i $O/${PFX}_top_kernel.ncu-rep --page source --csv > $O/${PFX}_top_kernel_source.csv 2>/dev/null
timeout 120 ncu -i $O/${PFX}_top_kernel.ncu-rep --page details > $O/${PFX}_top_kernel_details.txt 2>/dev/null
sz=$(stat -c %s $O/${PFX}_step_full.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -gt 35000000 ]; then echo "ncu-rep too big ($sz), keeping CSV only"; rm -f $O/${PFX}_step_full.ncu-rep; fi
ls -la $O
echo "== other configs"; date +%s
timeout 300 python tools/bench_configs.py > $O/configs.jsonl 2> $O/configs.err
echo "configs rc=$?"; cat $O/configs.jsonl
if [ "${1:-}" != "skip_tests" ]; then
  echo "== pytest gpu"; date +%s
  timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1
  echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
fi
date +%s
