#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 200 python -m pytest tests/test_gpu_round2.py -m gpu -x -q --timeout 60 -k "direct_halo" > $O/r2f_pytest.log 2>&1; echo "tests rc=$?" > $O/r2f_env.log
timeout 300 python bench.py --steps 20 --warmup 3 --workload cg-512 --no-cpu-baseline > $O/r2f_bench_cg_n1.json 2> $O/r2f_bench_cg_n1.err
tail -n 5 $O/r2f_pytest.log; cat $O/r2f_env.log
python -c "
import json
d=json.loads(open('$O/r2f_bench_cg_n1.json').read().strip().splitlines()[-1]); print(d['ms_per_step'])"
