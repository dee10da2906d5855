#!/bin/bash
# One gpurun call that produces every artefact profiles/ is built from:
#   bench line (+ autotune picks, parity, sustained), reference arm, ncu launch list with DRAM bytes of exactly K warm steps,
#   ncu --set full of one warm step, source-level capture of the top kernels, secondary configs, gpu tests.
#   (compute-sanitizer: tools/gpu_sanitize.sh, a separate gpurun call)
# usage (from the repo root on the GPU box): bash tools/gpu_evidence.sh [skip_tests]
set -u
O=gpurun_out
PFX=${PFX:-r02}          # round prefix of the ncu artefacts (profiles/${PFX}_*)
mkdir -p $O
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
echo "== smoke"; date +%s
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${PFX}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/${PFX}_smoke.log
echo "== bench"; date +%s
SB_DEBUG=1 BENCH_VERBOSE=1 SB_TUNE_SAVE=$O/tune.txt timeout 600 python bench.py --steps 20 --warmup 5 > $O/${PFX}_bench_1gpu.json 2> $O/${PFX}_bench_1gpu.err
echo "bench rc=$?"; tail -c 400 $O/${PFX}_bench_1gpu.json; echo
grep "^\[op\|first block\|first layer\|tconv" $O/${PFX}_bench_1gpu.err > $O/${PFX}_autotune_and_per_op.txt
grep "sb_conv_tc\] op" $O/${PFX}_bench_1gpu.err >> $O/${PFX}_autotune_and_per_op.txt
cp $O/tune.txt $O/${PFX}_autotune_picks.txt
echo "== reference arm"; date +%s
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $O/${PFX}_bench_reference.json 2> $O/bench_reference.err
echo "ref rc=$?"
echo "== ncu launch list (3 warm steps in the profiler window, the benched picks)"; date +%s
SB_TUNE_LOAD=$O/tune.txt timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none --profile-from-start off --csv --log-file $O/${PFX}_launches.csv python bench.py --steps 3 --warmup 3 --ncu-step > $O/ncu_list.log 2>&1
echo "list rc=$?"
python tools/ncu_summarize.py $O/${PFX}_launches.csv $O/${PFX}_launches_summary.md $O/${PFX}_tc_traffic.json 3 > /dev/null
echo "== ncu full (one warm step)"; date +%s
SB_TUNE_LOAD=$O/tune.txt timeout 600 ncu --set full --clock-control none --profile-from-start off \
  -f -o $O/${PFX}_step_full python bench.py --steps 1 --warmup 3 --ncu-step > $O/ncu_full.log 2>&1
echo "full rc=$?"
timeout 120 ncu -i $O/${PFX}_step_full.ncu-rep --page raw --csv > $O/${PFX}_step_full_raw.csv 2>/dev/null
python tools/ncu_full_summary.py $O/${PFX}_step_full_raw.csv $O/${PFX}_step_full_summary.md > /dev/null
echo "== ncu source-level capture of the fused first block + the heaviest conv"; date +%s
SB_TUNE_LOAD=$O/tune.txt timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:"k_conv01|k_conv_tc_prog" -c 3 -f -o $O/${PFX}_top_kernel python bench.py --steps 1 --warmup 3 --ncu-step > $O/ncu_top.log 2>&1
echo "top rc=$?"
timeout 120 ncu -i $O/${PFX}_top_kernel.ncu-rep --page details > $O/${PFX}_top_kernel_details.txt 2>/dev/null
sz=$(stat -c %s $O/${PFX}_step_full.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -gt 30000000 ]; then echo "ncu-rep too big ($sz), keeping CSV only"; rm -f $O/${PFX}_step_full.ncu-rep; fi
echo "== other configs"; date +%s
timeout 400 python tools/bench_configs.py > $O/${PFX}_other_configs.jsonl 2> $O/configs.err
echo "configs rc=$?"; cat $O/${PFX}_other_configs.jsonl
if [ "${1:-}" != "skip_tests" ]; then
  echo "== pytest gpu"; date +%s
  timeout 900 python -m pytest tests -m gpu -q > $O/${PFX}_pytest_gpu.log 2>&1
  echo "pytest rc=$?"; tail -3 $O/${PFX}_pytest_gpu.log
fi
ls -la $O | head -50
date +%s
