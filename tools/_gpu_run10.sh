nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 400 python -m pytest tests/test_gpu_nccl.py -x -q > gpurun_out/r10_pytest_nccl.log 2>&1; echo "pytest nccl rc=$?"
tail -3 gpurun_out/r10_pytest_nccl.log
for n in 1 2 4 8; do
  if [ $n = 1 ]; then L="timeout 200 python"; else L="timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n"; fi
  $L bench.py --gpus $n --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r10_bench_n$n.json 2> gpurun_out/r10_bench_n$n.err; echo "bench n=$n rc=$?"
done
for n in 1 8; do
  if [ $n = 1 ]; then L="timeout 200 python"; else L="timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n"; fi
  $L bench.py --gpus $n --steps 50 --warmup 3 --workload cg-512 --no-cpu-baseline > gpurun_out/r10_bench_cg_n$n.json 2> gpurun_out/r10_bench_cg_n$n.err; echo "cg n=$n rc=$?"
done
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 100 --warmup 5 --workload poisson512-strong > gpurun_out/r10_bench_strong512_n8.json 2> gpurun_out/r10_bench_strong512_n8.err; echo "strong512 rc=$?"
for f in gpurun_out/r10_bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","n_gpus","ms_per_step","achieved_gbs","gpu_launches")}, d["roofline"]["frac"], d["e2e"] and (d["e2e"]["value"], d["e2e"]["ms_per_step"]), d["config"]["workload"], d["clocks"])
except Exception as e: print("ERR", e)
PY
done
