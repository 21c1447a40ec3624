timeout 600 python tools/bench_next_rows.py 2>&1 | grep -v Warn | tee gpurun_out/r21_next_rows_n1.jsonl
echo "== EB4=4"
HPCLA_LIB=$PWD/linearalgebrampi.jl_b200/libhpcla_b200_eb4.so timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload poisson256-spmm8 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])"
echo "== EB4=2 (default)"
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload poisson256-spmm8 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])"
