#!/bin/bash
# compute-sanitizer on the smoke case and on small 2-rank (single-process rank-thread world) cases: one tool per GPU call
#   bash tools/run_sanitizer.sh memcheck | racecheck
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TOOL=${1:-memcheck}
O=gpurun_out
SEL='test_general_tiles_next_to_ghost_columns or (test_flat_kernel_irregular_rows and float32-int32 and 2) or (test_compact_row_walk and 48 and 2) or (test_plan_import_route and int32 and 2)'
export HPCLA_SANITIZER_RUN=1
timeout 1200 compute-sanitizer --tool $TOOL --error-exitcode 7 --log-file $O/r2_sanitizer_${TOOL}_smoke.log python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_sanitizer_${TOOL}_smoke.out 2>&1
echo "smoke rc=$?" > $O/r2_sanitizer_${TOOL}_rc.log
timeout 2400 compute-sanitizer --tool $TOOL --error-exitcode 7 --log-file $O/r2_sanitizer_${TOOL}_tests.log python -m pytest tests/test_gpu_round2.py -m gpu -q -x -k "$SEL" > $O/r2_sanitizer_${TOOL}_tests.out 2>&1
echo "tests rc=$?" >> $O/r2_sanitizer_${TOOL}_rc.log
cat $O/r2_sanitizer_${TOOL}_rc.log
tail -n 5 $O/r2_sanitizer_${TOOL}_smoke.log $O/r2_sanitizer_${TOOL}_tests.log $O/r2_sanitizer_${TOOL}_tests.out
