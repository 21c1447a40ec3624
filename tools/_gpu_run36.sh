for n in 8 4; do
  L="timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n"
  $L bench.py --gpus $n --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r36_bench_n$n.json 2> gpurun_out/r36_bench_n$n.err; echo "bench n=$n rc=$?"
done
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus 8 --steps 50 --warmup 3 --workload cg-512 --no-cpu-baseline > gpurun_out/r36_bench_cg_n8.json 2> gpurun_out/r36_bench_cg_n8.err; echo "cg n=8 rc=$?"
timeout 120 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r36_bench_n1.json 2> gpurun_out/r36_bench_n1.err; echo "bench n=1 rc=$?"
for f in gpurun_out/r36_bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","n_gpus","ms_per_step","achieved_gbs","gpu_launches")}, d["roofline"]["frac"], d["e2e"] and (d["e2e"]["value"], d["e2e"]["ms_per_step"]), d["config"]["workload"], d["clocks"]["reasons"])
except Exception as e: print("ERR", e)
PY
done
