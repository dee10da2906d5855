#!/bin/bash
# the ncu part of tools/gpu_evidence.sh alone (bench line first: it records the autotune picks the ncu runs replay)
set -u
O=gpurun_out
PFX=${PFX:-r02}
mkdir -p $O
export PYTHONUNBUFFERED=1
SB_DEBUG=1 BENCH_VERBOSE=1 SB_TUNE_SAVE=$O/tune.txt timeout 600 python bench.py --steps 20 --warmup 5 > $O/${PFX}_bench_1gpu.json 2> $O/${PFX}_bench_1gpu.err
echo "bench rc=$?"; python -c "import json; d=json.load(open('$O/${PFX}_bench_1gpu.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['strict_tensor_core']['value'])"
grep "^\[op\|first block\|first layer\|tconv" $O/${PFX}_bench_1gpu.err > $O/${PFX}_autotune_and_per_op.txt
grep "sb_conv_tc\] op" $O/${PFX}_bench_1gpu.err >> $O/${PFX}_autotune_and_per_op.txt
cp $O/tune.txt $O/${PFX}_autotune_picks.txt; cat $O/tune.txt | tr '\n' ';'; echo
SB_TUNE_LOAD=$O/tune.txt timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none --profile-from-start off --csv --log-file $O/${PFX}_launches.csv python bench.py --steps 3 --warmup 3 --ncu-step > $O/ncu_list.log 2>&1
echo "list rc=$?"
python tools/ncu_summarize.py $O/${PFX}_launches.csv $O/${PFX}_launches_summary.md $O/${PFX}_tc_traffic.json 3 > /dev/null
SB_TUNE_LOAD=$O/tune.txt timeout 600 ncu --set full --clock-control none --profile-from-start off \
  -f -o $O/${PFX}_step_full python bench.py --steps 1 --warmup 3 --ncu-step > $O/ncu_full.log 2>&1
echo "full rc=$?"
timeout 120 ncu -i $O/${PFX}_step_full.ncu-rep --page raw --csv > $O/${PFX}_step_full_raw.csv 2>/dev/null
python tools/ncu_full_summary.py $O/${PFX}_step_full_raw.csv $O/${PFX}_step_full_summary.md > /dev/null
SB_TUNE_LOAD=$O/tune.txt timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:"k_conv01|k_conv_tc_prog" -c 3 -f -o $O/${PFX}_top_kernel python bench.py --steps 1 --warmup 3 --ncu-step > $O/ncu_top.log 2>&1
echo "top rc=$?"
timeout 120 ncu -i $O/${PFX}_top_kernel.ncu-rep --page details > $O/${PFX}_top_kernel_details.txt 2>/dev/null
rm -f $O/${PFX}_step_full.ncu-rep
timeout 600 python -m pytest tests/test_gpu_zz_full_size.py tests/test_gpu_model.py tests/test_gpu_split_precision.py -m gpu -q -k "full or c4 or forward or bottomup or split" > $O/pytest_full.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_full.log
