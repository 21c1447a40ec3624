#!/bin/bash
# round 2, multi-GPU call: NCCL parity at every rank count the box allows, weak / strong / CG lines with the per-rank
# timeline and with CUDA graphs, PCIe probe with all GPUs at once, and the per-GPU spread of the single-GPU multiply.
#   bash tools/run_multi.sh <max gpus> <tag>
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
NMAX=${1:-2}
TAG=${2:-r2m}
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv > $O/${TAG}_env.log 2>&1
nvidia-smi topo -m >> $O/${TAG}_env.log 2>&1
nproc >> $O/${TAG}_env.log; ls /sys/devices/system/node | grep node >> $O/${TAG}_env.log
PORT=29600
for N in 2 4 8; do
  [ $N -le $NMAX ] || continue
  PORT=$((PORT+1))
  PYTHONPATH=$PWD:$PWD/tests timeout 900 $TR --nproc-per-node $N --master-port $PORT tests/_nccl_worker.py > $O/${TAG}_nccl_worker_n$N.log 2>&1
  echo "nccl worker N=$N rc=$? ok=$(grep -c NCCL_OK $O/${TAG}_nccl_worker_n$N.log)" >> $O/${TAG}_env.log
done
run_bench () {  # N name extra-args...
  local N=$1 NAME=$2; shift 2
  PORT=$((PORT+1))
  timeout 900 $TR --nproc-per-node $N --master-port $PORT bench.py --gpus $N --steps 100 --warmup 10 --no-cpu-baseline "$@" > $O/${TAG}_bench_${NAME}_n$N.json 2> $O/${TAG}_bench_${NAME}_n$N.err
  echo "bench $NAME N=$N rc=$?" >> $O/${TAG}_env.log
}
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $O/${TAG}_bench_weak_n1.json 2> $O/${TAG}_bench_weak_n1.err
for N in 2 4 8; do
  [ $N -le $NMAX ] || continue
  run_bench $N weak --timeline
  run_bench $N weak_graph --graph
  HPCLA_NCCL_MAX_CTAS=2 run_bench $N weak_ctas2
  run_bench $N weak_direct --halo direct --timeline
  run_bench $N strong256 --workload poisson256-strong --timeline
  run_bench $N strong256_direct --workload poisson256-strong --halo direct --timeline
  run_bench $N strong256_graph --workload poisson256-strong --graph
  run_bench $N cg --workload cg-512
  run_bench $N cg_graph --workload cg-512 --graph
  run_bench $N cg_direct --workload cg-512 --halo direct
done
if [ $NMAX -ge 8 ]; then
  run_bench 8 strong512 --workload poisson512-strong
  timeout 900 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --workload poisson512-strong > $O/${TAG}_bench_strong512_n1.json 2> $O/${TAG}_bench_strong512_n1.err
  # C5 "vs the 1-GPU run": CG on 512^3 on ONE GPU (cg-512 at --gpus 1 is the 256^3 weak member, so the grid is forced)
  HPCLA_BENCH_CG_GRID=512 timeout 900 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --workload cg-512 > $O/${TAG}_bench_cg512_n1.json 2> $O/${TAG}_bench_cg512_n1.err
fi
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --workload cg-512 > $O/${TAG}_bench_cg_n1.json 2> $O/${TAG}_bench_cg_n1.err
# per-GPU spread: the same single-GPU multiply on every GPU of the box, all at the same time
for ((g=0; g<NMAX; g++)); do
  CUDA_VISIBLE_DEVICES=$g timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > $O/${TAG}_spread_gpu$g.json 2> $O/${TAG}_spread_gpu$g.err &
done
wait
PORT=$((PORT+1))
timeout 600 $TR --nproc-per-node $NMAX --master-port $PORT tools/pcie_probe.py > $O/${TAG}_pcie_probe_n$NMAX.jsonl 2> $O/${TAG}_pcie_probe_n$NMAX.err
cat $O/${TAG}_env.log | tail -n 40
for f in $O/${TAG}_bench_*.json $O/${TAG}_spread_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["n_gpus"], d["config"]["workload"], "ms", round(d["ms_per_step"],5), "median", d.get("median_ms_per_step"), "e2e", (d.get("e2e") or {}).get("ms_per_step"), "copies", (d.get("e2e") or {}).get("copies_only_ms_per_step"))
    tl=(d.get("detail") or {}).get("timeline")
    if tl: print("   timeline", tl)
except Exception as e:
    print("no line:", e)
PY
done
