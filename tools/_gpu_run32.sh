timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 200 python tools/tune_spmv.py --workload poisson256 2>&1 | grep -v Warn | tail -1
timeout 200 python bench.py --steps 50 --warmup 3 --workload cg-512 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cg fused  ', d['ms_per_step'], d['value'], d['gpu_launches'])"
HPCLA_CG_UNFUSED=1 timeout 200 python bench.py --steps 50 --warmup 3 --workload cg-512 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cg unfused', d['ms_per_step'], d['value'], d['gpu_launches'])"
