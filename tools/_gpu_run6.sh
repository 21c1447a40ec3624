python -m pytest tests/test_gpu_nccl.py -x -q > gpurun_out/r6_pytest_nccl.log 2>&1; echo "pytest nccl rc=$?"
tail -5 gpurun_out/r6_pytest_nccl.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r6_bench_n1.json 2> gpurun_out/r6_bench_n1.err; echo "bench1 rc=$?"
$TR bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/r6_bench_n2.json 2> gpurun_out/r6_bench_n2.err; echo "bench2 rc=$?"; tail -5 gpurun_out/r6_bench_n2.err
$TR bench.py --gpus 2 --steps 100 --warmup 5 --workload poisson256-strong > gpurun_out/r6_bench_n2_strong.json 2> gpurun_out/r6_bench_n2_strong.err; echo "bench2s rc=$?"
$TR bench.py --gpus 2 --steps 50 --warmup 3 --workload cg-512 > gpurun_out/r6_bench_n2_cg.json 2> gpurun_out/r6_bench_n2_cg.err; echo "bench2cg rc=$?"; tail -5 gpurun_out/r6_bench_n2_cg.err
python bench.py --steps 50 --warmup 3 --workload cg-512 --no-cpu-baseline > gpurun_out/r6_bench_n1_cg.json 2> gpurun_out/r6_bench_n1_cg.err; echo "bench1cg rc=$?"
B="python bench.py --steps 5 --warmup 3 --workload cg-512 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file gpurun_out/r6_launches_cg.csv $B > gpurun_out/ncu6.log 2>&1; echo "ncu rc=$?"
for f in gpurun_out/r6_bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","n_gpus","ms_per_step","achieved_gbs","gpu_launches")}, d["roofline"]["frac"], d["e2e"] and (d["e2e"]["value"], d["e2e"]["ms_per_step"]), d["config"]["workload"], d["clocks"])
except Exception as e: print("ERR", e)
PY
done
