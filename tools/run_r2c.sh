#!/bin/bash
# round 2, GPU call C (1 GPU): nnz-split kernel v3 (values in registers, ballot scan, L2 window), compact walk without the slow barrier
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_round2.py -m gpu -x -q --timeout 900 > $O/r2c_pytest_round2.log 2>&1; echo "round2 tests rc=$?" > $O/r2c_env.log
for w in powerlaw-20m poisson256 poisson256-i64; do
  timeout 600 python bench.py --steps 30 --warmup 5 --workload $w --no-cpu-baseline > $O/r2c_bench_$w.json 2> $O/r2c_bench_$w.err
done
HPCLA_FLAT_L2_WINDOW=0 timeout 600 python bench.py --steps 30 --warmup 5 --workload powerlaw-20m --no-cpu-baseline > $O/r2c_bench_powerlaw-20m_nowindow.json 2> $O/r2c_bench_powerlaw-20m_nowindow.err
HPCLA_FLAT_KEEP_X=1 HPCLA_FLAT_L2_WINDOW=0 timeout 600 python bench.py --steps 30 --warmup 5 --workload powerlaw-20m --no-cpu-baseline > $O/r2c_bench_powerlaw-20m_keepx.json 2> $O/r2c_bench_powerlaw-20m_keepx.err
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload powerlaw-20m > $O/r2c_plain_powerlaw.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmv_flat -s 4 -c 1 -o $O/r2c_prof_flat_powerlaw -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload powerlaw-20m > $O/r2c_ncu_powerlaw.log 2>&1
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r2c_plain_poisson.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmv_cwalk -s 4 -c 1 -o $O/r2c_prof_cwalk_poisson -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r2c_ncu_poisson.log 2>&1
tail -n 3 $O/r2c_pytest_round2.log
cat $O/r2c_env.log
for f in $O/r2c_bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["ms_per_step"], d.get("median_ms_per_step"), d["roofline"]["frac"], d["roofline"]["kernel"], (d.get("e2e") or {}).get("ms_per_step"), (d.get("e2e") or {}).get("copies_only_ms_per_step"))
except Exception as e:
    print("no line:", e)
PY
done
