#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 125 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file $O/r2_sanitizer_memcheck_smoke.log python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_sanitizer_memcheck_smoke.out 2>&1
echo "memcheck smoke rc=$?"
tail -n 6 $O/r2_sanitizer_memcheck_smoke.log; tail -n 3 $O/r2_sanitizer_memcheck_smoke.out
