#!/bin/bash
# round 2, GPU call D (1 GPU): nnz-split kernel v4 (interleaved scans) at 14 / 10 / 7 warps, sparse x dense on compact tiles
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_round2.py -m gpu -x -q --timeout 900 > $O/r2d_pytest_round2.log 2>&1; echo "round2 tests rc=$?" > $O/r2d_env.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 900 > $O/r2d_pytest_parity.log 2>&1; echo "parity tests rc=$?" >> $O/r2d_env.log
timeout 600 python bench.py --steps 30 --warmup 5 --workload powerlaw-20m --no-cpu-baseline > $O/r2d_bench_powerlaw-20m.json 2> $O/r2d_bench_powerlaw-20m.err
for v in w7 w10; do
  HPCLA_LIB=$PWD/linearalgebrampi.jl_b200/libhpcla_b200_$v.so timeout 600 python bench.py --steps 30 --warmup 5 --workload powerlaw-20m --no-cpu-baseline > $O/r2d_bench_powerlaw-20m_$v.json 2> $O/r2d_bench_powerlaw-20m_$v.err
done
HPCLA_FLAT_KEEP_X=1 timeout 600 python bench.py --steps 30 --warmup 5 --workload powerlaw-20m --no-cpu-baseline > $O/r2d_bench_powerlaw-20m_keepx.json 2> $O/r2d_bench_powerlaw-20m_keepx.err
for k in 4 8 16; do
  timeout 600 python bench.py --steps 20 --warmup 3 --workload poisson256-spmm$k --no-cpu-baseline > $O/r2d_bench_spmm$k.json 2> $O/r2d_bench_spmm$k.err
done
HPCLA_SPMM_COMPACT=0 timeout 600 python bench.py --steps 20 --warmup 3 --workload poisson256-spmm8 --no-cpu-baseline > $O/r2d_bench_spmm8_plain.json 2> $O/r2d_bench_spmm8_plain.err
timeout 600 python bench.py --steps 20 --warmup 3 --workload cg-512 --no-cpu-baseline > $O/r2d_bench_cg_n1.json 2> $O/r2d_bench_cg_n1.err
timeout 600 python bench.py --steps 20 --warmup 3 --workload cg-512 --no-cpu-baseline --graph > $O/r2d_bench_cg_n1_graph.json 2> $O/r2d_bench_cg_n1_graph.err
timeout 600 python bench.py --steps 20 --warmup 3 --workload stencil27-192-T --no-cpu-baseline > $O/r2d_bench_stencil27-T.json 2> $O/r2d_bench_stencil27-T.err
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload powerlaw-20m > $O/r2d_plain_powerlaw.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmv_flat -s 4 -c 1 -o $O/r2d_prof_flat_powerlaw -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload powerlaw-20m > $O/r2d_ncu_powerlaw.log 2>&1
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload poisson256-spmm8 > $O/r2d_plain_spmm8.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_cwalk -s 2 -c 1 -o $O/r2d_prof_spmm8 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload poisson256-spmm8 > $O/r2d_ncu_spmm8.log 2>&1
tail -n 3 $O/r2d_pytest_round2.log $O/r2d_pytest_parity.log
cat $O/r2d_env.log
for f in $O/r2d_bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["ms_per_step"], d.get("median_ms_per_step"), d["roofline"]["frac"], d["roofline"]["kernel"], d["check"][:100])
except Exception as e:
    print("no line:", e)
PY
done
