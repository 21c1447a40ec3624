#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -x -q --timeout 300 -k "direct_halo or compact or graph" > $O/r2e_pytest.log 2>&1; echo "tests rc=$?" > $O/r2e_env.log
timeout 600 python bench.py --steps 20 --warmup 3 --workload cg-512 --no-cpu-baseline > $O/r2e_bench_cg_n1.json 2> $O/r2e_bench_cg_n1.err
HPCLA_CG_UNFUSED=1 timeout 600 python bench.py --steps 20 --warmup 3 --workload cg-512 --no-cpu-baseline > $O/r2e_bench_cg_n1_unfused.json 2> $O/r2e_bench_cg_n1_unfused.err
tail -n 5 $O/r2e_pytest.log; cat $O/r2e_env.log
for f in $O/r2e_bench_*.json; do echo "== $f"; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['check'][-120:])"; done
