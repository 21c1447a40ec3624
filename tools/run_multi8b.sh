#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TAG=r2n8
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
PORT=29900
run_bench () {
  local NAME=$1; shift
  PORT=$((PORT+1))
  timeout 300 $TR --nproc-per-node 8 --master-port $PORT bench.py --gpus 8 --steps 100 --warmup 10 --no-cpu-baseline --no-e2e "$@" > $O/${TAG}_bench_${NAME}_n8.json 2> $O/${TAG}_bench_${NAME}_n8.err
  echo "bench $NAME rc=$?" >> $O/${TAG}_env.log
}
run_bench weak_nccl
run_bench weak_direct --halo direct
run_bench strong256_nccl --workload poisson256-strong
run_bench strong256_direct --workload poisson256-strong --halo direct
run_bench cg_nccl --workload cg-512
run_bench cg_direct --workload cg-512 --halo direct
cat $O/${TAG}_env.log
for f in $O/${TAG}_bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["n_gpus"], d["config"]["workload"], "ms", round(d["ms_per_step"],5), "median", d.get("median_ms_per_step"))
except Exception as e:
    print("no line:", e)
PY
done
