#!/bin/bash
# round 2, GPU call B (1 GPU): round-2 tests again (graph fix, persistent nnz-split kernel), ncu on the two new kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_round2.py -m gpu -x -q --timeout 900 > $O/r2b_pytest_round2.log 2>&1; echo "round2 tests rc=$?" > $O/r2b_env.log
for w in powerlaw-20m poisson256 stencil27-192 laplace2d-1000; do
  timeout 600 python bench.py --steps 30 --warmup 5 --workload $w --no-cpu-baseline > $O/r2b_bench_$w.json 2> $O/r2b_bench_$w.err
done
HPCLA_COMPACT=1 timeout 600 python bench.py --steps 30 --warmup 5 --workload stencil27-192 --no-cpu-baseline > $O/r2b_bench_stencil27-192_compact.json 2> $O/r2b_bench_stencil27-192_compact.err
HPCLA_FLAT_KEEP_X=1 timeout 600 python bench.py --steps 30 --warmup 5 --workload powerlaw-20m --no-cpu-baseline > $O/r2b_bench_powerlaw-20m_keepx.json 2> $O/r2b_bench_powerlaw-20m_keepx.err
timeout 600 python bench.py --steps 30 --warmup 5 --graph --no-cpu-baseline > $O/r2b_bench_poisson256_graph.json 2> $O/r2b_bench_poisson256_graph.err
# ncu: the compact row walk (poisson256) and the nnz-split kernel (powerlaw-2m keeps the capture short; same kernel)
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r2b_plain_poisson.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmv_cwalk -s 4 -c 1 -o $O/r2b_prof_cwalk_poisson -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r2b_ncu_poisson.log 2>&1
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload powerlaw-20m > $O/r2b_plain_powerlaw.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmv_flat -s 4 -c 1 -o $O/r2b_prof_flat_powerlaw -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload powerlaw-20m > $O/r2b_ncu_powerlaw.log 2>&1
tail -n 3 $O/r2b_pytest_round2.log
cat $O/r2b_env.log
for f in $O/r2b_bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["ms_per_step"], d.get("median_ms_per_step"), d["roofline"]["frac"], d["roofline"]["kernel"], (d.get("e2e") or {}).get("ms_per_step"), (d.get("e2e") or {}).get("copies_only_ms_per_step"))
except Exception as e:
    print("no line:", e)
PY
done
