#!/bin/bash
# round 2, 8-GPU call: NCCL parity at 4 and 8 ranks, weak / strong / CG lines (CUDA graphs are bench.py's default), the per-rank
# timeline, the direct halo beside NCCL, the 1-GPU runs of the 512^3 problems (efficiency "vs the 1-GPU run"), PCIe with all GPUs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TAG=${1:-r2m8}
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv > $O/${TAG}_env.log 2>&1
nvidia-smi topo -m >> $O/${TAG}_env.log 2>&1
nproc >> $O/${TAG}_env.log; ls /sys/devices/system/node | grep node >> $O/${TAG}_env.log; free -g | head -2 >> $O/${TAG}_env.log
PORT=29700
# the 512^3 problems on ONE GPU, in the background on GPUs 6 and 7 while the 4-rank runs use GPUs 0-3
CUDA_VISIBLE_DEVICES=7 timeout 900 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --workload poisson512-strong > $O/${TAG}_bench_strong512_n1.json 2> $O/${TAG}_bench_strong512_n1.err &
CUDA_VISIBLE_DEVICES=6 HPCLA_BENCH_CG_GRID=512 timeout 900 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --workload cg-512 > $O/${TAG}_bench_cg512_n1.json 2> $O/${TAG}_bench_cg512_n1.err &
run_bench () {  # N name extra-args...
  local N=$1 NAME=$2; shift 2
  PORT=$((PORT+1))
  timeout 600 $TR --nproc-per-node $N --master-port $PORT bench.py --gpus $N --steps 100 --warmup 10 --no-cpu-baseline "$@" > $O/${TAG}_bench_${NAME}_n$N.json 2> $O/${TAG}_bench_${NAME}_n$N.err
  echo "bench $NAME N=$N rc=$?" >> $O/${TAG}_env.log
}
PORT=$((PORT+1))
PYTHONPATH=$PWD:$PWD/tests timeout 600 $TR --nproc-per-node 4 --master-port $PORT tests/_nccl_worker.py > $O/${TAG}_nccl_worker_n4.log 2>&1
echo "nccl worker N=4 rc=$? ok=$(grep -c NCCL_OK $O/${TAG}_nccl_worker_n4.log)" >> $O/${TAG}_env.log
run_bench 4 weak
run_bench 4 strong256 --workload poisson256-strong
run_bench 4 cg --workload cg-512
wait
PORT=$((PORT+1))
PYTHONPATH=$PWD:$PWD/tests timeout 600 $TR --nproc-per-node 8 --master-port $PORT tests/_nccl_worker.py > $O/${TAG}_nccl_worker_n8.log 2>&1
echo "nccl worker N=8 rc=$? ok=$(grep -c NCCL_OK $O/${TAG}_nccl_worker_n8.log)" >> $O/${TAG}_env.log
run_bench 8 weak
run_bench 8 weak_nograph_timeline --timeline
run_bench 8 weak_direct --halo direct --timeline
run_bench 8 strong256 --workload poisson256-strong
run_bench 8 strong256_nograph_timeline --workload poisson256-strong --timeline
run_bench 8 strong256_direct --workload poisson256-strong --halo direct
run_bench 8 strong512 --workload poisson512-strong
run_bench 8 cg --workload cg-512
run_bench 8 cg_nograph --workload cg-512 --no-graph
PORT=$((PORT+1))
timeout 300 $TR --nproc-per-node 8 --master-port $PORT tools/pcie_probe.py > $O/${TAG}_pcie_probe_n8.jsonl 2> $O/${TAG}_pcie_probe_n8.err
# per-GPU spread: the same single-GPU multiply on every GPU of the box at the same time
for g in 0 1 2 3 4 5 6 7; do
  CUDA_VISIBLE_DEVICES=$g timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > $O/${TAG}_spread_gpu$g.json 2> $O/${TAG}_spread_gpu$g.err &
done
wait
tail -n 30 $O/${TAG}_env.log
for f in $O/${TAG}_bench_*.json $O/${TAG}_spread_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["n_gpus"], d["config"]["workload"], "ms", round(d["ms_per_step"],5), "median", d.get("median_ms_per_step"), "e2e", (d.get("e2e") or {}).get("ms_per_step"), "copies", (d.get("e2e") or {}).get("copies_only_ms_per_step"))
    tl=(d.get("detail") or {}).get("timeline")
    if tl: print("   timeline", [[round(v,4) for v in r.values()] for r in tl["per_rank_ms_from_x_ready"]])
except Exception as e:
    print("no line:", e)
PY
done
