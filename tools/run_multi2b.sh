#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TAG=r2n2
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
PORT=29800
PYTHONPATH=$PWD:$PWD/tests timeout 600 $TR --nproc-per-node 2 --master-port $PORT tests/_nccl_worker.py > $O/${TAG}_nccl_worker_n2.log 2>&1
echo "nccl worker N=2 rc=$? ok=$(grep -c NCCL_OK $O/${TAG}_nccl_worker_n2.log)" > $O/${TAG}_env.log
run_bench () {
  local NAME=$1; shift
  PORT=$((PORT+1))
  timeout 400 $TR --nproc-per-node 2 --master-port $PORT bench.py --gpus 2 --steps 100 --warmup 10 --no-cpu-baseline "$@" > $O/${TAG}_bench_${NAME}_n2.json 2> $O/${TAG}_bench_${NAME}_n2.err
  echo "bench $NAME rc=$?" >> $O/${TAG}_env.log
}
run_bench weak_nccl
run_bench weak_nccl_graph --graph
run_bench weak_direct --halo direct
run_bench weak_direct_tl --halo direct --timeline
HPCLA_HALO_FLAG_KERNEL=1 run_bench weak_direct_flagkernel --halo direct
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
    tl=(d.get("detail") or {}).get("timeline")
    if tl: print("   timeline", [[round(v,4) for v in r.values()] for r in tl["per_rank_ms_from_x_ready"]])
except Exception as e:
    print("no line:", e)
PY
done
