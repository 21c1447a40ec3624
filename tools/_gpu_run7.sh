timeout 300 python -m pytest tests/test_gpu_nccl.py -x -q > gpurun_out/r7_pytest_nccl.log 2>&1; echo "pytest nccl rc=$?"
tail -3 gpurun_out/r7_pytest_nccl.log
TR="timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r7_bench_n1.json 2> gpurun_out/r7_bench_n1.err; echo "bench1 rc=$?"
$TR bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/r7_bench_n2.json 2> gpurun_out/r7_bench_n2.err; echo "bench2 rc=$?"; tail -5 gpurun_out/r7_bench_n2.err
$TR bench.py --gpus 2 --steps 100 --warmup 5 --workload poisson256-strong > gpurun_out/r7_bench_n2_strong.json 2> gpurun_out/r7_bench_n2_strong.err; echo "bench2s rc=$?"
$TR bench.py --gpus 2 --steps 50 --warmup 3 --workload cg-512 > gpurun_out/r7_bench_n2_cg.json 2> gpurun_out/r7_bench_n2_cg.err; echo "bench2cg rc=$?"; tail -5 gpurun_out/r7_bench_n2_cg.err
$TR bench.py --gpus 2 --steps 50 --warmup 5 --workload stencil27-192 > gpurun_out/r7_bench_n2_s27.json 2> gpurun_out/r7_bench_n2_s27.err; echo "bench2 s27 rc=$?"
for f in gpurun_out/r7_bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","n_gpus","ms_per_step","achieved_gbs","gpu_launches")}, d["roofline"]["frac"], d["e2e"] and (d["e2e"]["value"], d["e2e"]["ms_per_step"]), d["config"]["workload"], d["clocks"])
except Exception as e: print("ERR", e)
PY
done
