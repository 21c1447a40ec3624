#!/bin/bash
# round 2, GPU call A (1 GPU): new-kernel tests first, then the rest of the suite, probes and the headline benches
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $O/r2a_env.log 2>&1
nproc >> $O/r2a_env.log; free -g | head -2 >> $O/r2a_env.log
timeout 1500 python -m pytest tests/test_gpu_round2.py -m gpu -x -q --timeout 600 > $O/r2a_pytest_round2.log 2>&1; echo "round2 tests rc=$?" >> $O/r2a_env.log
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_nccl.py -m gpu -q --timeout 600 > $O/r2a_pytest_parity.log 2>&1; echo "parity tests rc=$?" >> $O/r2a_env.log
timeout 120 tools/gather_probe > $O/r2a_gather_probe.txt 2>&1
timeout 300 python tools/pcie_probe.py > $O/r2a_pcie_probe_n1.jsonl 2> $O/r2a_pcie_probe_n1.err
timeout 600 python bench.py --steps 50 --warmup 5 > $O/r2a_bench_poisson256.json 2> $O/r2a_bench_poisson256.err
HPCLA_COMPACT=0 timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $O/r2a_bench_poisson256_plainwalk.json 2> $O/r2a_bench_poisson256_plainwalk.err
for w in powerlaw-20m stencil27-192 poisson256-i64 laplace2d-1000; do
  timeout 600 python bench.py --steps 30 --warmup 5 --workload $w --no-cpu-baseline > $O/r2a_bench_$w.json 2> $O/r2a_bench_$w.err
done
HPCLA_FLAT_KEEP_X=1 timeout 600 python bench.py --steps 30 --warmup 5 --workload powerlaw-20m --no-cpu-baseline > $O/r2a_bench_powerlaw-20m_keepx.json 2> $O/r2a_bench_powerlaw-20m_keepx.err
HPCLA_SPMV_KIND=general timeout 600 python bench.py --steps 30 --warmup 5 --workload powerlaw-20m --no-cpu-baseline > $O/r2a_bench_powerlaw-20m_oldgeneral.json 2> $O/r2a_bench_powerlaw-20m_oldgeneral.err
tail -3 $O/r2a_pytest_round2.log $O/r2a_pytest_parity.log
cat $O/r2a_env.log
for f in $O/r2a_bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["ms_per_step"], d.get("median_ms_per_step"), d["roofline"]["frac"], d["roofline"]["kernel"], (d.get("e2e") or {}).get("ms_per_step"), (d.get("e2e") or {}).get("copies_only_ms_per_step"))
except Exception as e:
    print("no line:", e)
PY
done
