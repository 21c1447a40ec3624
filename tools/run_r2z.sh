#!/bin/bash
# round 2, final 1-GPU call on HEAD: whole GPU suite, the driver's bench command + its ncu launch list, ring variants, smoke
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 > $O/r2z_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?" > $O/r2z_env.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2z_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r2z_env.log
timeout 300 python bench.py --steps 20 --warmup 5 > $O/r2z_bench_n1.json 2> $O/r2z_bench_n1.err && \
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file $O/r2z_launches_poisson.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2z_ncu_launches.log 2>&1
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/r2z_bench_reference.json 2> $O/r2z_bench_reference.err
HPCLA_RING=1 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > $O/r2z_bench_poisson256_ring.json 2> $O/r2z_bench_poisson256_ring.err
HPCLA_RING=2 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload poisson256-spmm8 > $O/r2z_bench_spmm8_ring.json 2> $O/r2z_bench_spmm8_ring.err
HPCLA_RING=2 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload poisson256-spmm4 > $O/r2z_bench_spmm4_ring.json 2> $O/r2z_bench_spmm4_ring.err
HPCLA_COMPACT=1 HPCLA_RING=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --workload stencil27-192 > $O/r2z_bench_stencil27_ring.json 2> $O/r2z_bench_stencil27_ring.err
tail -n 4 $O/r2z_pytest_gpu.log; cat $O/r2z_env.log; tail -n 2 $O/r2z_smoke.log
for f in $O/r2z_bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["ms_per_step"], d.get("median_ms_per_step"), (d.get("roofline") or {}).get("frac"), (d.get("roofline") or {}).get("kernel"), (d.get("cpu_baseline") or {}).get("value"))
except Exception as e:
    print("no line:", e)
PY
done
