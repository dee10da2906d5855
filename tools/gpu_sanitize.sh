#!/bin/bash
# compute-sanitizer over the smoke path and one reduced C4 step per kernel variant / precision (tools/sanitize_step.py)
set -u
O=gpurun_out
PFX=${PFX:-r02}
mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > $O/${PFX}_sanitizer_memcheck_smoke.log 2>&1; echo "memcheck smoke rc=$?"; tail -3 $O/${PFX}_sanitizer_memcheck_smoke.log
timeout 900 compute-sanitizer --target-processes all --tool memcheck --print-limit 20 python tools/sanitize_step.py ${SAN_MODE:-} > $O/${PFX}_sanitizer_memcheck_c4_variants.log 2>&1; echo "memcheck variants rc=$?"; tail -4 $O/${PFX}_sanitizer_memcheck_c4_variants.log
timeout 900 compute-sanitizer --target-processes all --tool racecheck --print-limit 20 python tools/sanitize_step.py quick > $O/${PFX}_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 $O/${PFX}_sanitizer_racecheck.log
